"""Per-source-line hot spots of one kernel from an .ncu-rep (ncu --page source --csv): samples, executed instructions,
shared-memory excess wavefronts and the dominant stall reasons.  Usage: ncu_lines.py REP KERNEL_REGEX [TOP]"""
import csv, subprocess, sys, io, os
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
lines = []
seen_files = set()
cur_file = None
stop = False
for r in rows:
    if r and r[0] in ("File Path", "File Name"):
        if r[1] in seen_files:
            break  # second launch of the kernel: keep the first only
        seen_files.add(r[1])
        cur_file = r[1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or r[0] == "":
        continue
    r = list(r)
    r[1] = f"[{cur_file.split('/')[-1]}] " + r[1].strip()
    lines.append(r)
ix = {h: i for i, h in enumerate(hdr)}
# header has duplicate "Source"; first is CUDA source
def val(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_s = sum(val(r, "# Samples") for r in lines) or 1
tot_i = sum(val(r, "Instructions Executed") for r in lines) or 1
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"kernel {kern}: {int(tot_s)} samples, {int(tot_i)} warp instructions")
agg = {}
for s in stalls:
    agg[s] = sum(val(r, s) for r in lines)
print("stalls:", ", ".join(f"{k[6:]} {100*v/tot_s:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
lines.sort(key=lambda r: -val(r, "Instructions Executed" if os.environ.get("BY_INST") else "# Samples"))
for r in lines[:top]:
    st = sorted(((val(r, s), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{r[0]:>4} smp {100*val(r,'# Samples')/tot_s:5.1f}% inst {100*val(r,'Instructions Executed')/tot_i:5.1f}% "
          f"shx {int(val(r,'L1 Wavefronts Shared Excessive')):>9} {st[0][1]}/{st[1][1]} | {r[1].strip()[:100]}")
