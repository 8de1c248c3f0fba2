"""Summarise an .ncu-rep (ncu --page raw --csv) into one CSV row per launch: duration, DRAM traffic (bytes), tensor / fma
pipe activity, issue activity, registers.  Usage: ncu_summary.py REP OUT.csv"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
def get(r, k, scale=False):
    if k not in ix: return ""
    v = r[ix[k]].replace(",", "")
    try: v = float(v)
    except ValueError: return r[ix[k]]
    if scale: v *= SCALE.get(units[ix[k]], 1)
    return v
cols = [("kernel", "Kernel Name", False), ("grid", "Grid Size", False), ("block", "Block Size", False),
        ("duration_ms", "gpu__time_duration.sum", True), ("dram_read_bytes", "dram__bytes_read.sum", True),
        ("dram_write_bytes", "dram__bytes_write.sum", True),
        ("tensor_pipe_pct_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
        ("fma_pipe_pct_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", False),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", False),
        ("regs_per_thread", "launch__registers_per_thread", False),
        ("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
        ("warp_instructions", "smsp__inst_executed.sum", False)]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([c[0] for c in cols])
    for r in rows[2:]:
        if len(r) != len(hdr): continue
        w.writerow([get(r, k, s) for _, k, s in cols])
print(open(out).read())
