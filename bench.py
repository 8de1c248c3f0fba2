#!/usr/bin/env python
"""Benchmark of the RAT-SPN hot path on B200 (driver contract: one JSON line on stdout from rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): RatSpn MNIST-shaped training step, F=784, D=5, R=40, S=I=40, C=1, batch 4096 per
GPU (weak scaling), fp32: forward log-likelihood + backward to all parameters (+ one flat NCCL gradient all-reduce
when N > 1).  Metric: samples/s over all GPUs.

  value     device-timed (CUDA events, max over ranks), inputs resident in HBM
  e2e       same step through the public module API with the batch coming from pinned host memory every step
            (H2D copy inside the timed region) and the loss read back to the host (D2H)
  roofline  the dominant kernel (the slowest of the three tcgen05 calls of the widest Sum layer S_1), timed live with
            CUDA events, L2 flushed between launches; ``achieved`` counts ALGORITHMIC flops (forward, dX and dW
            are one 2*N*d*c^2*S*R contraction each -- the second GEMM the dX formulation executes is not counted);
            ``tc_kernels`` lists forward / dX / dW, the S_1 forward+backward aggregate and the whole-step fraction
  extra     (N = 1 only) the other BASELINE.json configs -- cfg 1, cfg 3 (eager + CUDA graph), cfg 4, cfg 5-narrow,
            cfg 5-wide -- each next to the CPU oracle port timed in this same run on a bounded sample
  cpu_baseline  the CPU oracle port of the reference algorithm on the host cores, on a bounded micro-batch sample

``--impl reference`` times only that CPU path (the reference is pure Python/PyTorch and is not present on the GPU box;
the oracle restates its eager algorithm op for op, see oracle/spn_oracle.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(F=784, D=5, R=40, S=40, I=40, C=1)
BATCH_PER_GPU = 4096
METRIC = "ratspn_fwd_bwd_samples_per_sec"
UNIT = "samples/s"
# arithmetic of the path: fp32 inputs / outputs / accumulators everywhere; the shared-weight Sum layers multiply bf16
# operands on the tensor cores (tcgen05 kind::f16) -- inside the north star's 1e-4 relative log-likelihood tolerance
DTYPE = "f32 io/accumulate; bf16 mma operands (Sum layers), tf32 x 2 split operands (leaf)"


def workload_name(rows_per_gpu=BATCH_PER_GPU):
    return (f"RatSpn MNIST-shaped training step F={CFG['F']} D={CFG['D']} R={CFG['R']} S=I={CFG['S']} C=1, "
            f"batch {rows_per_gpu}/GPU, fwd+bwd fp32")


def contraction_flops_fwd(batch):
    """Algorithmic flops of the Sum-layer contractions, forward: 2*N*d*c^2*S*R per inner layer + root."""
    F, D, R, S, I = CFG["F"], CFG["D"], CFG["R"], CFG["S"], CFG["I"]
    total, per_layer = 0, []
    for j in range(1, D):
        d = 2 ** (D - j)
        ic = I * I if j == 1 else S * S
        fl = 2 * batch * d * ic * S * R
        per_layer.append(fl)
        total += fl
    root = 2 * batch * R * S * S * CFG["C"]
    return total + root, per_layer


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's eager algorithm)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_baseline(micro=8, iters=1, seed=0):
    from oracle import spn_oracle as O
    torch.manual_seed(seed)
    arch = O.Arch(**CFG)
    g = torch.Generator().manual_seed(seed)
    sd = {"permutation": torch.stack([torch.randperm(CFG["F"], generator=g) for _ in range(CFG["R"])], dim=-1),
          "_leaf.base_leaf.mean_param": torch.randn(1, CFG["F"], CFG["I"], CFG["R"], generator=g).requires_grad_(True),
          "_leaf.base_leaf.std_param": torch.randn(1, CFG["F"], CFG["I"], CFG["R"], generator=g).requires_grad_(True)}
    for j in range(1, CFG["D"]):
        d, ic = arch.sum_dims(j)
        sd[f"_inner_layers.{2 * j - 1}.weight_param"] = torch.randn(1, d, ic, CFG["S"], CFG["R"], generator=g).requires_grad_(True)
    sd["root.weight_param"] = torch.randn(1, 1, CFG["R"] * arch.ic_top, CFG["C"], 1, generator=g).requires_grad_(True)
    x = torch.rand(micro, 1, CFG["F"], 1, 1, generator=g)
    times = []
    for it in range(iters + 1):  # first iteration is a warm-up
        t0 = time.perf_counter()
        P = O.params_from_ratspn_state(sd, arch)
        loss = -O.forward(P, arch, x).mean()
        loss.backward()
        times.append(time.perf_counter() - t0)
        for v in sd.values():
            if v.requires_grad:
                v.grad = None
    t = float(np.median(times[1:])) if iters >= 1 else times[0]
    return {"value": micro / t, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"fwd+bwd of {micro} rows at full width (the eager broadcast+logsumexp algorithm cannot hold "
                      f"more: 671 GB temporary at 4096), median of {iters} after 1 warm-up, {t:.2f} s per micro-batch"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    steps = max(1, min(args.steps, 3))
    cb = cpu_baseline(micro=8, iters=steps)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": 1e3 * 8 / cb["value"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(), "note": "CPU host cores; each step = one 8-row micro-batch"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# The other BASELINE.json configs (cfg 1, 3, 4, 5): this implementation on the GPU (tools/bench_configs.py) and, beside each,
# the CPU oracle port of the reference's algorithm timed in the same run on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------------------------
def _median_time(fn, iters=2, warmup=1):
    ts = []
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        fn()
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def _oracle_ratspn_state(F, D, R, S, I, C, g, grad=True):
    from oracle import spn_oracle as O
    arch = O.Arch(F=F, D=D, R=R, S=S, I=I, C=C)
    mk = (lambda *sh: torch.randn(*sh, generator=g).requires_grad_(grad))
    sd = {"permutation": torch.stack([torch.randperm(F, generator=g) for _ in range(R)], dim=-1),
          "_leaf.base_leaf.mean_param": mk(1, F, I, R), "_leaf.base_leaf.std_param": mk(1, F, I, R)}
    for j in range(1, D):
        d, ic = arch.sum_dims(j)
        sd[f"_inner_layers.{2 * j - 1}.weight_param"] = mk(1, d, ic, S, R)
    sd["root.weight_param"] = mk(1, 1, R * arch.ic_top, C, 1)
    return arch, sd


def cpu_port_configs():
    """CPU legs of the secondary configs (oracle port, all host threads).  Every entry states its bounded sample."""
    from oracle import spn_oracle as O
    cores = torch.get_num_threads()
    out = {}
    g = torch.Generator().manual_seed(0)

    # cfg 1: the reference's own CPU-runnable case, full size (256 rows), fwd + bwd
    arch, sd = _oracle_ratspn_state(784, 3, 5, 10, 10, 1, g)
    x = torch.randn(256, 1, 784, 1, 1, generator=g)

    def cfg1():
        for v in sd.values():
            v.grad = None
        (-O.forward(O.params_from_ratspn_state(sd, arch), arch, x).mean()).backward()
    t = _median_time(cfg1, iters=3)
    out["cfg1"] = {"ms": 1e3 * t, "samples_per_s": 256 / t, "cores": cores, "kind": "port", "sample": "full workload (256 rows), median of 3"}

    # cfg 3: one SAC policy update of the Humanoid-shaped CSPN at full batch 256 (sb3.py:146-184, :336-344)
    import ratspn_b200 as rb
    torch.manual_seed(0)
    cc = rb.CspnConfig()
    cc.F_cond = (376,)
    cc.F, cc.R, cc.D, cc.I, cc.S, cc.C = 17, 3, 4, 3, 3, 1
    cc.dropout, cc.leaf_base_class, cc.tanh_squash = 0.0, rb.RatNormal, True
    cc.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
    cc.feat_layers, cc.sum_param_layers, cc.dist_param_layers = [64], [64], [64]
    cc.cond_layers_inner_act = torch.nn.LeakyReLU
    sd3 = {k: v.detach().clone() for k, v in rb.CSPN(cc).state_dict().items()}   # host-side parameter container only
    for k in list(sd3):
        if sd3[k].dtype.is_floating_point and (k.endswith(".weight") or k.endswith(".bias")):
            sd3[k].requires_grad_(True)
    arch3 = O.Arch(F=17, D=4, R=3, S=3, I=3, C=1, tanh_squash=True)
    Bc = 256
    obs = torch.randn(Bc, 376, generator=g)
    failed = torch.rand(Bc, 17, generator=g) < 0.05
    ev = torch.where(failed, torch.atanh(torch.rand(Bc, 17, generator=g) * 1.8 - 0.9), torch.full((Bc, 17), float("nan")))
    ev = ev.view(1, Bc, 17, 1, 1)
    xs = torch.randn(1, Bc, 17, 1, 1, generator=g).clamp(-2, 2)

    def cfg3():
        for v in sd3.values():
            v.grad = None
        ns = O.NoiseSource(generator=g)
        P = O.cspn_params(sd3, arch3, obs, min_sigma=0.001, max_sigma=1.0)   # the reference runs the gating net twice
        smp = O.sample_onehot(P, arch3, ns, n=1, evidence=ev)
        P = O.cspn_params(sd3, arch3, obs, min_sigma=0.001, max_sigma=1.0)
        H = O.recursive_entropy(P, arch3, ns, sample_size=5, marginal_mask=failed, differentiable=True)
        (-(0.1 * H + smp.sample.reshape(Bc, 17).sum(-1)).mean()).backward()
        P = O.cspn_params(sd3, arch3, obs, min_sigma=0.001, max_sigma=1.0)
        (-O.forward(P, arch3, xs).mean()).backward()
    t = _median_time(cfg3, iters=2)
    out["cfg3"] = {"us_per_update": 1e6 * t, "cores": cores, "kind": "port", "sample": "full workload (batch 256), median of 2"}

    # cfg 4: ancestral sampling at cfg-2 width; the eager gather holds ~64 samples (268 GB at 65536)
    arch4, sd4 = _oracle_ratspn_state(784, 5, 40, 40, 40, 1, g, grad=False)
    P4 = O.params_from_ratspn_state(sd4, arch4)
    n4 = 64

    def cfg4():
        with torch.no_grad():
            O.sample_index(P4, arch4, O.NoiseSource(generator=g), n=(n4,))
    t = _median_time(cfg4, iters=2)
    out["cfg4"] = {"samples_per_s": n4 / t, "cores": cores, "kind": "port",
                   "sample": f"{n4} of the 65536 samples per call (the eager weight gather needs 268 GB at 65536), median of 2"}
    del P4, sd4

    # cfg 5-narrow: per-conditional weights (R=8, S=I=16, D=6, F=1024), LL forward and forward + backward, 4 conditionals
    arch5 = O.Arch(F=1024, D=6, R=8, S=16, I=16, C=1)
    w5 = 4
    P5 = {"perm": torch.stack([torch.randperm(1024, generator=g) for _ in range(8)], dim=-1),
          "mu": torch.randn(w5, 1024, 16, 8, generator=g).requires_grad_(True),
          "sigma": (torch.rand(w5, 1024, 16, 8, generator=g) * 0.9 + 0.05).requires_grad_(True), "lw": []}
    raws = []
    for j in range(1, 6):
        d, ic = arch5.sum_dims(j)
        raws.append(torch.randn(w5, d, ic, 16, 8, generator=g).requires_grad_(True))
    raws.append(torch.randn(w5, 1, 8 * 256, 1, 1, generator=g).requires_grad_(True))
    x5 = torch.randn(1, w5, 1024, 1, 1, generator=g)

    def cfg5(grad):
        P5["lw"] = [torch.log_softmax(t_, dim=2) for t_ in raws[:-1]]
        P5["lw_root"] = torch.log_softmax(raws[-1], dim=2)
        if grad:
            for v in raws + [P5["mu"], P5["sigma"]]:
                v.grad = None
            (-O.forward(P5, arch5, x5).mean()).backward()
        else:
            with torch.no_grad():
                O.forward(P5, arch5, x5)
    t_f = _median_time(lambda: cfg5(False), iters=2)
    t_fb = _median_time(lambda: cfg5(True), iters=2)
    out["cfg5_narrow"] = {"conditionals_per_s_forward": w5 / t_f, "conditionals_per_s_fwd_bwd": w5 / t_fb, "cores": cores,
                          "kind": "port", "sample": f"{w5} conditionals (of 2048 / 512), one row each, median of 2"}
    del P5, raws

    # cfg 5-wide: R=64, S=I=64, D=6, F=1024, one weight set of 4.2 GB: LL forward of ONE row of the 2048 per set
    try:
        import psutil
        free_gb = psutil.virtual_memory().available / 2 ** 30
    except Exception:
        free_gb = 0.0
    if free_gb > 40:
        archw = O.Arch(F=1024, D=6, R=64, S=64, I=64, C=1)
        Pw = {"perm": torch.stack([torch.randperm(1024, generator=g) for _ in range(64)], dim=-1),
              "mu": torch.randn(1, 1024, 64, 64, generator=g), "sigma": torch.rand(1, 1024, 64, 64, generator=g) * 0.9 + 0.05,
              "lw": []}
        for j in range(1, 6):
            d, ic = archw.sum_dims(j)
            Pw["lw"].append(torch.log_softmax(torch.randn(1, d, ic, 64, 64, generator=g), dim=2))
        Pw["lw_root"] = torch.log_softmax(torch.randn(1, 1, 64 * 4096, 1, 1, generator=g), dim=2)
        xw = torch.randn(1, 1, 1024, 1, 1, generator=g)

        def cfgw():
            with torch.no_grad():
                O.forward(Pw, archw, xw)
        t = _median_time(cfgw, iters=1, warmup=1)
        out["cfg5_wide"] = {"rows_per_s_forward": 1 / t, "cores": cores, "kind": "port",
                            "sample": "1 row of the 2048 rows of one weight set (2.1 GB broadcast temporary per row), 1 timed pass"}
    else:
        out["cfg5_wide"] = {"skipped": f"needs > 40 GB of free host memory for one 4.2 GB weight set + temporaries ({free_gb:.0f} GB free)"}
    return out


def extra_configs(with_cpu=True):
    """GPU numbers of the secondary configs (functions of tools/bench_configs.py) with the CPU port beside each."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("spn_bench_configs", os.path.join(ROOT, "tools", "bench_configs.py"))
    bc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bc)
    res = {}
    for name, fn in bc.CONFIGS:
        t0 = time.time()
        try:
            res[name] = fn()
        except Exception as e:  # every config is independent
            res[name] = {"error": f"{type(e).__name__}: {str(e).splitlines()[0] if str(e) else ''}"}
        res[name]["wall_s"] = time.time() - t0
        torch.cuda.empty_cache()
    if with_cpu:
        t0 = time.time()
        try:
            cpu = cpu_port_configs()
        except Exception as e:
            cpu = {"error": f"{type(e).__name__}: {e}"}
        for k, v in cpu.items():
            tgt = k if k in res else next((n for n in res if n.startswith(k)), None)
            if isinstance(v, dict) and tgt is not None:
                res[tgt]["cpu_baseline"] = v
            else:
                res.setdefault("cpu_errors", {})[k] = v
        res["cpu_wall_s"] = time.time() - t0
    return res


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampling
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop, self.thread = index, [], threading.Event(), None
        self.nvml = None
        try:  # NVML directly: a sample every few ms, so that even a 45 ms timed region gets several
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
            bits = [getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                    getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                    getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4)]
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            reasons(h)
            self.nvml = (pynvml, h, reasons, bits, mx)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        pynvml, h, reasons, bits, mx = self.nvml
        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        mask = int(reasons(h))
        self.rows.append([str(sm), str(mx)] + ["Active" if mask & b else "Not Active" for b in bits])

    def _loop(self):
        if self.nvml is not None:
            while not self.stop.is_set():
                try:
                    self._sample_nvml()
                except Exception:
                    pass
                self.stop.wait(0.004)
            return
        while not self.stop.is_set():  # fall-back: nvidia-smi (slow: ~100 ms per sample)
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)
        if not self.rows:  # the region was shorter than one sampling period: sample right at its end
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active")
                                                         for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def build_model(device):
    import ratspn_b200 as rb
    torch.manual_seed(0)
    np.random.seed(0)
    cfg = rb.RatSpnConfig()
    cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = (CFG[k] for k in "FDRSIC")
    cfg.dropout, cfg.leaf_base_class, cfg.leaf_base_kwargs, cfg.tanh_squash = 0.0, rb.RatNormal, {}, False
    return rb.RatSpn(cfg).to(device)


_RESULT_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configs (cfg 1, 3, 4, 5) and their CPU legs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch rows per GPU (default); strong: --batch rows in total, split over the GPUs")
    ap.add_argument("--graph", action="store_true",
                    help="replay the step from a CUDA graph (1 GPU only; uses torch's capturable Adam: 8.83 ms vs 8.65 ms for the "
                         "default eager launches with the flat Adam -- the step is device-bound)")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON record): everything else written to fd 1 by libraries (the NCCL version
    # banner, warnings) is sent to stderr
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    import ratspn_b200 as rb
    from ratspn_b200 import _abi
    from ratspn_b200.parallel import FlatAdam, FlatGradAllReducer, broadcast_parameters

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path for --impl ours")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    if args.scaling == "strong":
        from ratspn_b200.parallel import shard_bounds
        lo, hi = shard_bounds(args.batch, world, rank)
        B = hi - lo   # contiguous shard of the global batch
    global_rows = args.batch if args.scaling == "strong" else world * B
    model = build_model(dev)
    broadcast_parameters(model)
    overlap = os.environ.get("SPN_DP_OVERLAP", "1") != "0"
    reducer = FlatGradAllReducer(model.parameters(), overlap=overlap)
    nbuf = 4
    gen = torch.Generator().manual_seed(1234 + rank)
    host = [torch.rand(B, 1, CFG["F"], 1, 1, generator=gen).pin_memory() for _ in range(nbuf)]
    resident = [h.to(dev) for h in host]

    sum_layers = [m for m in model.modules() if hasattr(m, "invalidate_weight_cache")]

    # Adam as in mnist_gen_train.py:599, as one launch over flat buffers (spn_adam_step); SPN_TORCH_ADAM=1 -> torch's fused Adam
    if os.environ.get("SPN_TORCH_ADAM") or (args.graph and world == 1):  # FlatAdam keeps its step count on the host: not capturable
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True, capturable=True)
    else:
        opt = FlatAdam(reducer, lr=1e-3, module=model)

    def step(x):
        """One training step: LL forward, backward, gradient all-reduce, Adam update (mnist_gen_train.py:173-205)."""
        # the step changes the weights, so the bf16 operand images are rebuilt every step (no cached weights)
        for m in sum_layers:
            m.invalidate_weight_cache()
        reducer.zero_()
        loss = -model(x).mean()
        loss.backward()
        reducer.all_reduce()
        opt.step()
        return loss.detach()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(resident[i % nbuf])
    sync_all()

    # Optionally the whole step is captured once into a CUDA graph and replayed.  Off by default: the step is
    # device-bound, and the asynchronous per-parameter all-reduce hooks are not captured at N > 1.
    use_graph = args.graph and world == 1
    xin = resident[0].clone()
    l0 = _abi.LAUNCHES
    graph_err = None
    run = None
    if use_graph:
        try:
            graphed = rb.GraphedCall(lambda: step(xin), warmup=1)

            def run(x, non_blocking=False):
                xin.copy_(x, non_blocking=non_blocking)
                return graphed()
        except Exception as e:  # e.g. a collective that cannot be captured: fall back to eager launches
            graph_err = f"{type(e).__name__}: {str(e).splitlines()[0]}"
            run = None
    if run is None:
        def run(x, non_blocking=False):
            xin.copy_(x, non_blocking=non_blocking)
            return step(xin)
    # kernels of libspn_b200.so launched by one step, counted by CUPTI (torch.profiler) on one eager step: everything
    # that is not an ATen / optimizer / NCCL / memset kernel.  (_abi.LAUNCHES counts ABI calls, several kernels each.)
    sync_all()
    l0 = _abi.LAUNCHES
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(resident[0])
        torch.cuda.synchronize()
    abi_calls_per_step = _abi.LAUNCHES - l0
    foreign = ("at::", "nccl", "multi_tensor", "cub::", "Memset", "Memcpy", "cutlass", "cublas", "elementwise")
    launches_per_step = sum(e.count for e in prof.key_averages()
                            if e.device_type.name == "CUDA" and not any(t in e.key for t in foreign))
    for i in range(2):
        run(resident[i % nbuf])
    sync_all()

    # ---- device-timed region: K steps, inputs resident in HBM
    with ClockSampler(local) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for i in range(args.steps):
            run(resident[i % nbuf])
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
    launches = launches_per_step * args.steps
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = global_rows * args.steps / (ms / 1e3)

    # ---- end-to-end: pinned host batch -> H2D -> step -> loss D2H, every step
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.steps):
        loss = run(host[i % nbuf], non_blocking=True)
        loss_host = loss.item()
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = global_rows * args.steps / float(t.item())

    # ---- roofline: the tensor-core kernels of the widest Sum layer (S_1), each timed alone with CUDA events
    roof, kernels = None, None
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        ncu = {}
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                ncu = json.load(f)
        except Exception:
            pass
        from ratspn_b200 import ops

        def time_alone(fn, reps=10):
            flush = torch.empty(64 * 1024 * 1024, device=dev, dtype=torch.float32)  # 256 MB > 126 MB L2
            for _ in range(2):
                fn()
            ts = []
            for _ in range(reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            return float(np.mean(ts))

        with torch.no_grad():
            h0 = model._leaf(resident[0], perm32=model._perm32()[0], layout="drbc")
            w1 = model._inner_layers[1].weight_param.detach()
            wimg, logz = ops.tc_prepare(w1, True)
            fhand = ops.tc_fhand_buffer(h0, w1)
            out1 = ops.tc_sum_forward(h0, wimg, w1, logz, fhand)
            g1 = torch.randn_like(out1) / B
            t_fwd = time_alone(lambda: ops.tc_sum_forward(h0, wimg, w1, logz, fhand))
            hand = ops.tc_hand_buffer(h0, w1)
            t_bx = time_alone(lambda: ops.tc_sum_backward_x(h0, out1, g1, wimg, w1, fhand, hand))
            t_bw = time_alone(lambda: ops.tc_sum_backward_w(fhand, hand, h0, out1, g1, w1, logz))
            t_prep = time_alone(lambda: ops.tc_prepare(w1, True))
        # leaf layer alone (HBM bound: the [d0][R][B][I] activation is written once by the forward, its gradient read
        # once by the backward): tensor-core kernels (leaf_tc.cu) and, beside them, the exact fp32 kernels
        leaf = {}
        try:
            mu_l, sg_l = (t_.detach().clone().requires_grad_(True) for t_ in model._leaf.base_leaf._params_for(True))
            perm_l, card_l = model._perm32()[0], model._leaf.cardinality
            act_bytes = h0.numel() * 4
            gl = torch.randn_like(h0) / B
            hbm = peaks.get("hbm_gbs", 0.0) or 0.0
            for tc_on in (True, False):
                ops.LEAF_TENSOR_CORES = tc_on
                try:
                    t_lf = time_alone(lambda: ops.leaf_forward(resident[0], mu_l.detach(), sg_l.detach(), perm_l, card_l, False, "drbc"))
                    o_l = ops.leaf_forward(resident[0], mu_l, sg_l, perm_l, card_l, False, "drbc")
                    def bwd_only():
                        mu_l.grad = sg_l.grad = None
                        o_l.backward(gl, retain_graph=True)
                    t_lb = time_alone(bwd_only)
                finally:
                    ops.LEAF_TENSOR_CORES = True
                leaf["tensor_cores" if tc_on else "exact_fp32"] = {
                    "fwd_ms": t_lf, "bwd_ms": t_lb, "activation_bytes": act_bytes,
                    "fwd_write_GBps": act_bytes / t_lf / 1e6, "bwd_read_GBps": act_bytes / t_lb / 1e6,
                    "fwd_frac_of_hbm_peak": (act_bytes / t_lf / 1e6 / hbm) if hbm else None,
                    "bwd_frac_of_hbm_peak": (act_bytes / t_lb / 1e6 / hbm) if hbm else None}
            leaf["note"] = ("fwd = x transpose + parameter image + leaf kernel (spn_leaf_tc_fwd / spn_leaf_tiled_fwd); bwd = parameter "
                            "gradients; algorithmic bytes = one pass over the [d0][R][B][I] fp32 activation (839 MB)")
        except Exception as e:   # the roofline line must not die on the side measurement
            leaf = {"error": repr(e)}
        total_fwd, per_layer = contraction_flops_fwd(B)
        fl = per_layer[0]
        peak = peaks.get("bf16_tflops", 1590.0)
        peak_sus = peaks.get("bf16_tflops_sustained", peak)
        src = "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)" if peaks else "fallback 1590 (B200_PROFILING.md)"
        traffic_src = ncu.get("_source", None)
        kernels = {}
        # algorithmic flops per call: forward, dX, dW = one [B x 1600] x [1600 x 40] contraction per (scope,
        # repetition) each.  Executed = algorithmic for all three: forward and dX run ONE small-K GEMM per row tile on the
        # tensor cores (M = 128, K = 40 padded to 48, N = 1600) plus the contraction over the left channels on the fp32
        # pipe out of tensor memory.
        for key, name, t, executed in (("fwd", "tsum_fwd_kernel (spn_sum_tc_fwd)", t_fwd, 1),
                                       ("dx", "tsum_dx_kernel (spn_sum_tc_bwd_x, incl. the G pre-pass)", t_bx, 1),
                                       ("dw", "sum_tc_bwd_w_kernel (spn_sum_tc_bwd_w, incl. finish)", t_bw, 1)):
            ach = fl / (t / 1e3) / 1e12
            kernels[key] = {"call": name, "layer": "S_1 (d=16, K=1600, N=40, R=40)", "ms": t, "algorithmic_flops": fl,
                            "achieved_tflops": ach, "frac": ach / peak, "executed_flops": executed * fl,
                            "executed_frac": executed * ach / peak, "traffic": ncu.get(name.split(" ")[0]),
                            "traffic_source": traffic_src}
        t_s1 = t_fwd + t_bx + t_bw
        kernels["s1_fwd_bwd"] = {"ms": t_s1, "algorithmic_flops": 3 * fl, "achieved_tflops": 3 * fl / (t_s1 / 1e3) / 1e12,
                                 "frac": 3 * fl / (t_s1 / 1e3) / 1e12 / peak}
        step_ms = ms / args.steps
        kernels["whole_step"] = {"ms": step_ms, "algorithmic_flops": 3 * total_fwd,
                                 "achieved_tflops": 3 * total_fwd / (step_ms / 1e3) / 1e12,
                                 "frac_of_sustained_peak": 3 * total_fwd / (step_ms / 1e3) / 1e12 / peak_sus,
                                 "note": "Sum-layer contractions only (fwd + bwd = 3x forward, SURVEY 8d) over the whole training "
                                         "step incl. leaf, root, operand images, all-reduce and Adam"}
        kernels["tc_prepare"] = {"call": "spn_sum_tc_prepare (logz + operand image)", "ms": t_prep}
        kernels["leaf"] = leaf
        dom = max(("fwd", "dx", "dw"), key=lambda k: kernels[k]["ms"])
        roof = {"bound": "tensor", "kernel": f"{kernels[dom]['call']}, layer S_1 (d=16, K=1600, N=40, R=40), "
                                             f"{fl:.3e} algorithmic flop per call (2*N*d*c^2*S*R)",
                "achieved": kernels[dom]["achieved_tflops"], "peak": peak, "unit": "TFLOP/s",
                "frac": kernels[dom]["frac"], "executed_frac": kernels[dom]["executed_frac"],
                "traffic": kernels[dom]["traffic"], "traffic_source": traffic_src, "kernel_ms": kernels[dom]["ms"],
                "peak_source": src}

    if rank == 0:
        cb = None
        if not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            cb = cpu_baseline(micro=8, iters=5)
        extra = None
        if world == 1 and not args.no_extra:
            torch.cuda.empty_cache()
            extra = extra_configs(with_cpu=not args.no_cpu_baseline)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
                "config": {"workload": workload_name(B), "global_batch": global_rows, "rows_per_gpu": B,
                           "parallelism": f"dp{world}",
                           "step": "LL forward + backward + gradient all-reduce + Adam update (one flat kernel)",
                           "cuda_graph": bool(use_graph and graph_err is None),
                           "l2": "per-step working set (leaf activations 839 MB + 317 MB weights) exceeds the 126 MB "
                                 "L2; 4 rotating input batches"},
                "clocks": clk.summary(),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * CFG["F"] * 4,
                        "d2h_bytes_per_step": 4},
                "gpu_launches": launches, "abi_calls_per_step": abi_calls_per_step, "roofline": roof, "tc_kernels": kernels, "cpu_baseline": cb, "loss": loss_host,
                "extra": extra}
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
