"""Secondary measurements for the BASELINE.json configs other than the headline one (cfg 1, 3, 4, 5).

bench.py runs ``CONFIGS`` at N = 1 and reports them under ``extra`` next to the CPU oracle port timed in the same run;
run stand-alone, this script writes the GPU numbers as JSON (default: gpurun_out/configs.json).  CUDA-event timing, 3 warm-ups.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ratspn_b200 as rb  # noqa: E402
from ratspn_b200 import _abi  # noqa: E402

DEV = torch.device("cuda", 0)


def timed(fn, reps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    l0 = _abi.LAUNCHES
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, (_abi.LAUNCHES - l0) / reps


def ratspn(F, D, R, S, I, C=1, tanh=False, leaf_kwargs=None):
    cfg = rb.RatSpnConfig()
    cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = F, D, R, S, I, C
    cfg.dropout, cfg.leaf_base_class, cfg.leaf_base_kwargs, cfg.tanh_squash = 0.0, rb.RatNormal, leaf_kwargs or {}, tanh
    return rb.RatSpn(cfg).to(DEV)


def cfg1():
    torch.manual_seed(0); np.random.seed(0)
    m = ratspn(784, 3, 5, 10, 10)
    x = torch.randn(256, 1, 784, 1, 1, device=DEV)

    def step():
        for p in m.parameters():
            p.grad = None
        (-m(x).mean()).backward()
    ms, launches = timed(step, reps=20)
    return {"config": "cfg1 RatSpn F=784 D=3 R=5 S=I=10 B=256 fwd+bwd", "ms": ms, "samples_per_s": 256 / ms * 1e3,
            "abi_calls_per_step": launches}


def cfg3():
    """CSPN SAC policy update pattern (sb3.py:89-109, :146-184, :336-344)."""
    torch.manual_seed(0); np.random.seed(0)
    cfg = rb.CspnConfig()
    cfg.F_cond = (376,)
    cfg.F, cfg.R, cfg.D, cfg.I, cfg.S, cfg.C = 17, 3, 4, 3, 3, 1
    cfg.dropout = 0.0
    cfg.leaf_base_class = rb.RatNormal
    cfg.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
    cfg.tanh_squash = True
    cfg.feat_layers, cfg.sum_param_layers, cfg.dist_param_layers = [64], [64], [64]
    cfg.cond_layers_inner_act = torch.nn.LeakyReLU
    m = rb.CSPN(cfg).to(DEV)
    B = 256
    obs = torch.randn(B, 376, device=DEV)
    failed = torch.rand(B, 17, device=DEV) < 0.05
    ev = torch.where(failed, torch.atanh(torch.rand(B, 17, device=DEV) * 1.8 - 0.9), torch.full((B, 17), float("nan"), device=DEV))
    ev = ev.view(1, B, 17, 1, 1)
    xs = torch.randn(1, B, 17, 1, 1, device=DEV).clamp(-2, 2)

    def update():
        m.zero_grad(set_to_none=True)
        with m.conditioned(obs):  # one gating-network pass for sample + entropy (one backward)
            ctx = m.sample(mode="onehot", condition=obs, evidence=ev)
            H, _ = m.recursive_entropy_approx(condition=obs, sample_size=5, marginal_mask=failed)
        act = ctx.sample.reshape(B, 17)
        (-(0.1 * H + act.sum(-1)).mean()).backward()
        m.zero_grad(set_to_none=True)
        (-m(xs, condition=obs).mean()).backward()

    def rollout():
        with torch.no_grad():
            m.sample(mode="index", condition=obs[:1], evidence=ev[:, :1])

    ms, launches = timed(update, reps=5)
    ms_r, launches_r = timed(rollout, reps=20)
    # the same update captured into one CUDA graph (graphs.GraphedCall): grads pre-allocated and zeroed in place
    params = [p for p in m.parameters() if p.requires_grad]
    for p in params:
        p.grad = torch.zeros_like(p)

    def update_static():
        for p in params:
            p.grad.zero_()
        with m.conditioned(obs):  # one gating-network pass for sample + entropy (one backward)
            ctx = m.sample(mode="onehot", condition=obs, evidence=ev)
            H, _ = m.recursive_entropy_approx(condition=obs, sample_size=5, marginal_mask=failed)
        act = ctx.sample.reshape(B, 17)
        (-(0.1 * H + act.sum(-1)).mean()).backward()
        (-m(xs, condition=obs).mean()).backward()
        return act
    g = rb.GraphedCall(update_static)
    ms_g, _ = timed(g, reps=20)
    obs1, ev1 = obs[:1].clone(), ev[:, :1].clone()

    def rollout_static():
        with torch.no_grad():
            return m.sample(mode="index", condition=obs1, evidence=ev1).sample
    gr = rb.GraphedCall(rollout_static)
    ms_gr, _ = timed(gr, reps=50)
    return {"config": "cfg3 CSPN SAC update obs 376 / act 17, R=I=S=3 D=4, B=256: onehot sample + evidence, recursive "
                      "entropy (n=5, mask) fwd+bwd, LL fwd+bwd", "us_per_update": ms * 1e3, "us_per_update_cuda_graph": ms_g * 1e3,
            "rollout_index_sample_with_evidence_B1_us_cuda_graph": ms_gr * 1e3, "abi_calls_per_update": launches,
            "rollout_index_sample_with_evidence_B1_us": ms_r * 1e3, "abi_calls_per_rollout": launches_r,
            "note": "latency bound: 1467 circuit parameters per conditional; roofline fractions are meaningless here"}


def cfg4():
    torch.manual_seed(0); np.random.seed(0)
    m = ratspn(784, 5, 40, 40, 40)
    n = 65536
    out = {}
    with torch.no_grad():
        for tag, kw in (("index", {}), ("mpe", {"is_mpe": True})):
            ms, launches = timed(lambda: m.sample(mode="index", n=n, **kw), reps=5)
            out[tag] = {"ms": ms, "samples_per_s": n / ms * 1e3, "output_GB_per_s": n * 784 * 4 / ms / 1e6,
                        "abi_calls": launches}
    out["config"] = "cfg4 RatSpn F=784 D=5 R=40 S=I=40 ancestral sampling, 65536 samples"
    return out


def cfg5_narrow(w=2048):
    """Per-sample-weight (CSPN) regime: R=8, S=I=16, D=6, F=1024; one weight set per conditional, injected directly."""
    torch.manual_seed(0); np.random.seed(0)
    cfg = rb.CspnConfig()
    cfg.F_cond = (4,)
    cfg.F, cfg.R, cfg.D, cfg.I, cfg.S, cfg.C = 1024, 8, 6, 16, 16, 1
    cfg.dropout = 0.0
    cfg.leaf_base_class = rb.RatNormal
    cfg.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
    cfg.tanh_squash = False
    cfg.feat_layers, cfg.sum_param_layers, cfg.dist_param_layers = [4], [4], [4]
    # the gating heads of this width would be huge (Linear(4, 2e6)); build the structure only and inject weights
    m = rb.RatSpn.__new__(rb.CSPN)
    rb.RatSpn.__init__(m, cfg)
    m.replace_layer_params()
    m = m.to(DEV)
    nbytes = 0
    for layer in list(m._inner_layers) + [m.root]:
        if isinstance(layer, rb.Sum):
            shape = (w, layer.in_features, layer.in_channels, layer.out_channels, layer.num_repetitions)
            layer.weight_param = torch.log_softmax(torch.randn(shape, device=DEV), dim=2)
            nbytes += layer.weight_param.numel() * 4
        else:
            layer.num_conditionals = w
    base = m._leaf.base_leaf
    base.mean_param = torch.randn(w, 1024, 16, 8, device=DEV)
    base.std_param = torch.rand(w, 1024, 16, 8, device=DEV) * 0.9 + 0.05
    nbytes += 2 * base.mean_param.numel() * 4
    x = torch.randn(1, w, 1024, 1, 1, device=DEV)
    with torch.no_grad():
        ms, launches = timed(lambda: rb.RatSpn.forward(m, x), reps=5)
    return {"config": f"cfg5-narrow per-sample weights R=8 S=I=16 D=6 F=1024, w={w} conditionals, n=1: LL forward",
            "ms": ms, "conditionals_per_s": w / ms * 1e3, "parameter_bytes_read": nbytes,
            "achieved_GB_per_s": nbytes / ms / 1e6, "hbm_peak_GB_per_s": 6548.2, "frac": nbytes / ms / 1e6 / 6548.2,
            "abi_calls": launches}


def cfg5_narrow_train(w=512):
    """Same per-sample-weight regime, LL forward + backward w.r.t. every per-conditional parameter (the CSPN training step
    minus the gating MLP): weights read in the forward, read again and their gradient written in the backward."""
    torch.manual_seed(0); np.random.seed(0)
    cfg = rb.CspnConfig()
    cfg.F_cond = (4,)
    cfg.F, cfg.R, cfg.D, cfg.I, cfg.S, cfg.C = 1024, 8, 6, 16, 16, 1
    cfg.dropout = 0.0
    cfg.leaf_base_class = rb.RatNormal
    cfg.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
    cfg.tanh_squash = False
    cfg.feat_layers, cfg.sum_param_layers, cfg.dist_param_layers = [4], [4], [4]
    m = rb.RatSpn.__new__(rb.CSPN)
    rb.RatSpn.__init__(m, cfg)
    m.replace_layer_params()
    m = m.to(DEV)
    params, nbytes = [], 0
    for layer in list(m._inner_layers) + [m.root]:
        if isinstance(layer, rb.Sum):
            shape = (w, layer.in_features, layer.in_channels, layer.out_channels, layer.num_repetitions)
            layer.weight_param = torch.log_softmax(torch.randn(shape, device=DEV), dim=2).requires_grad_(True)
            params.append(layer.weight_param)
            nbytes += layer.weight_param.numel() * 4
        else:
            layer.num_conditionals = w
    base = m._leaf.base_leaf
    base.mean_param = torch.randn(w, 1024, 16, 8, device=DEV).requires_grad_(True)
    base.std_param = (torch.rand(w, 1024, 16, 8, device=DEV) * 0.9 + 0.05).requires_grad_(True)
    params += [base.mean_param, base.std_param]
    nbytes += 2 * base.mean_param.numel() * 4
    x = torch.randn(1, w, 1024, 1, 1, device=DEV)

    def step():
        for p in params:
            p.grad = None
        (-rb.RatSpn.forward(m, x).mean()).backward()
    ms, launches = timed(step, reps=5)
    return {"config": f"cfg5-narrow per-sample weights R=8 S=I=16 D=6 F=1024, w={w} conditionals, n=1: LL forward + backward",
            "ms": ms, "conditionals_per_s": w / ms * 1e3, "parameter_bytes": nbytes,
            "algorithmic_bytes (3 x parameters: read, read, write gradient)": 3 * nbytes,
            "achieved_GB_per_s": 3 * nbytes / ms / 1e6, "hbm_peak_GB_per_s": 6548.2, "frac": 3 * nbytes / ms / 1e6 / 6548.2,
            "abi_calls": launches}


def cfg5_narrow_train_raw(w=512):
    """As cfg5_narrow_train, but the trainable tensors are the UNNORMALISED weights (what a gating network emits):
    (a) reference semantics, F.log_softmax by ATen + normalised-weight kernels; (b) CspnConfig.fuse_weight_norm semantics,
    raw weights straight into spn_sum_ps_fwd_raw / _bwd_raw."""
    res = {}
    for fused in (False, True):
        torch.manual_seed(0); np.random.seed(0)
        cfg = rb.CspnConfig()
        cfg.F_cond = (4,)
        cfg.F, cfg.R, cfg.D, cfg.I, cfg.S, cfg.C = 1024, 8, 6, 16, 16, 1
        cfg.dropout = 0.0
        cfg.leaf_base_class = rb.RatNormal
        cfg.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
        cfg.tanh_squash = False
        cfg.feat_layers, cfg.sum_param_layers, cfg.dist_param_layers = [4], [4], [4]
        m = rb.RatSpn.__new__(rb.CSPN)
        rb.RatSpn.__init__(m, cfg)
        m.replace_layer_params()
        m = m.to(DEV)
        raws, inner = [], []
        for layer in list(m._inner_layers) + [m.root]:
            if isinstance(layer, rb.Sum):
                shape = (w, layer.in_features, layer.in_channels, layer.out_channels, layer.num_repetitions)
                raws.append(torch.randn(shape, device=DEV).requires_grad_(True))
                inner.append(layer)
            else:
                layer.num_conditionals = w
        base = m._leaf.base_leaf
        base.mean_param = torch.randn(w, 1024, 16, 8, device=DEV).requires_grad_(True)
        base.std_param = (torch.rand(w, 1024, 16, 8, device=DEV) * 0.9 + 0.05).requires_grad_(True)
        params = raws + [base.mean_param, base.std_param]
        x = torch.randn(1, w, 1024, 1, 1, device=DEV)

        def step():
            for p in params:
                p.grad = None
            for layer, raw in zip(inner, raws):
                is_root = layer is m.root
                layer._raw_weights = fused and not is_root
                layer.weight_param = raw if layer._raw_weights else torch.log_softmax(raw, dim=2)
            (-rb.RatSpn.forward(m, x).mean()).backward()
        ms, launches = timed(step, reps=5)
        res["fused" if fused else "log_softmax_by_aten"] = {"ms": ms, "abi_calls": launches}
        del m, raws, params
        torch.cuda.empty_cache()
    res["config"] = f"cfg5-narrow training step on UNNORMALISED per-conditional weights (w={w}): F.log_softmax + kernels vs fused"
    return res


def cfg5_wide(rows=2048, w=1):
    """BASELINE config 5 as named: wide conditional circuit R=64, S=I=64, D=6, F=1024; batch 16384 = 8 conditionals x 2048
    rows sharded over 8 GPUs, i.e. ONE conditional with 2048 rows per GPU: every inner Sum layer is a dense
    per-conditional contraction (4.3e12 flop forward) and runs on the tcgen05 kernels (P = 64, streamed weight images).
    The weights are injected directly (a gating head of this width would be Linear(h, 5e8))."""
    torch.manual_seed(0); np.random.seed(0)
    cfg = rb.CspnConfig()
    cfg.F_cond = (4,)
    cfg.F, cfg.R, cfg.D, cfg.I, cfg.S, cfg.C = 1024, 64, 6, 64, 64, 1
    cfg.dropout = 0.0
    cfg.leaf_base_class = rb.RatNormal
    cfg.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
    cfg.tanh_squash = False
    cfg.feat_layers, cfg.sum_param_layers, cfg.dist_param_layers = [4], [4], [4]
    m = rb.RatSpn.__new__(rb.CSPN)
    rb.RatSpn.__init__(m, cfg)
    m.replace_layer_params()
    m = m.to(DEV)
    params, flops = [], 0
    for layer in list(m._inner_layers) + [m.root]:
        if isinstance(layer, rb.Sum):
            shape = (w, layer.in_features, layer.in_channels, layer.out_channels, layer.num_repetitions)
            layer.weight_param = torch.log_softmax(torch.randn(shape, device=DEV), dim=2).requires_grad_(True)
            params.append(layer.weight_param)
            if layer is not m.root:
                flops += 2 * rows * w * layer.weight_param[0].numel()
        else:
            layer.num_conditionals = w
    base = m._leaf.base_leaf
    base.mean_param = torch.randn(w, 1024, 64, 64, device=DEV).requires_grad_(True)
    base.std_param = (torch.rand(w, 1024, 64, 64, device=DEV) * 0.9 + 0.05).requires_grad_(True)
    params += [base.mean_param, base.std_param]
    x = torch.randn(rows, w, 1024, 1, 1, device=DEV)
    assert m._tensor_core_sets_path_ok(x, False, m.max_layer_index)
    with torch.no_grad():
        ms_f, l_f = timed(lambda: rb.RatSpn.forward(m, x), reps=5)

    def step():
        for p in params:
            p.grad = None
        (-rb.RatSpn.forward(m, x).mean()).backward()
    ms_t, l_t = timed(step, reps=3)
    m.use_tensor_cores = False
    with torch.no_grad():
        ms_e, _ = timed(lambda: rb.RatSpn.forward(m, x), reps=1, warmup=1)
    peak = 1672.2
    return {"config": f"cfg5-wide CSPN R=64 S=I=64 D=6 F=1024, {w} conditional(s) x {rows} rows per GPU (tcgen05 path, bf16 operands)",
            "forward": {"ms": ms_f, "rows_per_s": rows * w / ms_f * 1e3, "sum_layer_flops": flops,
                        "achieved_tflops": flops / ms_f / 1e9, "frac_of_bf16_peak": flops / ms_f / 1e9 / peak, "abi_calls": l_f},
            "forward_backward": {"ms": ms_t, "rows_per_s": rows * w / ms_t * 1e3, "sum_layer_flops": 3 * flops,
                                 "achieved_tflops": 3 * flops / ms_t / 1e9, "frac_of_bf16_peak": 3 * flops / ms_t / 1e9 / peak,
                                 "abi_calls": l_t},
            "forward_exact_fp32_kernels_ms": ms_e}


CONFIGS = (("cfg1", cfg1), ("cfg3", cfg3), ("cfg4", cfg4), ("cfg5_narrow", cfg5_narrow),
           ("cfg5_narrow_train", cfg5_narrow_train), ("cfg5_narrow_train_raw", cfg5_narrow_train_raw), ("cfg5_wide", cfg5_wide))


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "configs.json")
    res = {"device": torch.cuda.get_device_name(0), "lib": _abi.version()}
    only = set(sys.argv[2:])   # optional: names of the configs to run
    for name, fn in CONFIGS:
        if only and name not in only:
            continue
        t0 = time.time()
        try:
            res[name] = fn()
        except Exception as e:  # keep going: every config is independent
            res[name] = {"error": f"{type(e).__name__}: {e}"}
        res[name]["wall_s"] = time.time() - t0
        print(name, json.dumps(res[name]), flush=True)
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
