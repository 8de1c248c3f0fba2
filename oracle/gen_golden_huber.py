"""Golden vectors for ``huber_entropy_lb`` (rat_spn.py:825-951) from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE.  Usage from the repo root:  python oracle/gen_golden_huber.py  -> tests/golden/huber.npz
Stores, for one RatSpn (F=16, D=3, R=3, S=3, I=4, C=2, leaf bounds as in the SAC actor) and one CSPN (F=8, D=2):
the state_dict / condition, and the bound for several (layer_index, flags, marginal_mask) combinations plus the
autograd gradients of the default call.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference  # noqa: E402


def main():
    rat_spn, cspn, _, dists, _ = load_reference()
    out = {}
    torch.manual_seed(11)
    np.random.seed(11)
    cfg = rat_spn.RatSpnConfig()
    cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = 16, 3, 3, 3, 4, 2
    cfg.dropout, cfg.leaf_base_class, cfg.tanh_squash = 0.0, dists.RatNormal, False
    cfg.leaf_base_kwargs = {"min_sigma": 0.05, "max_sigma": 1.5}
    m = rat_spn.RatSpn(cfg)
    with torch.no_grad():
        for p in m.parameters():
            if p.dtype.is_floating_point and p.requires_grad:
                p.add_(0.3 * torch.randn_like(p))
    for k, v in m.state_dict().items():
        out["rat__sd__" + k] = v.detach().numpy().copy()
    mask = torch.zeros(1, 16, dtype=torch.bool)
    mask[0, [2, 9, 10]] = True
    out["rat__mask"] = mask.numpy()
    for li in (2, 4, 6):
        lb, _ = m.huber_entropy_lb(layer_index=li, verbose=False)
        out[f"rat__lb_layer{li}"] = lb.detach().numpy()
    lb, log = m.huber_entropy_lb(verbose=True)
    out["rat__lb_default"] = lb.detach().numpy()
    out["rat__log_keys"] = np.array(sorted(log.keys()))
    out["rat__log_vals"] = np.array([log[k] for k in sorted(log.keys())], dtype=np.float64)
    lb, _ = m.huber_entropy_lb(verbose=False, marginal_mask=mask)
    out["rat__lb_mask"] = lb.detach().numpy()
    lb, _ = m.huber_entropy_lb(verbose=False, detach_weights=True, add_sub_weight_ent=True)
    out["rat__lb_flags"] = lb.detach().numpy()
    for p in m.parameters():
        p.grad = None
    lb, _ = m.huber_entropy_lb(verbose=False, add_sub_weight_ent=True)
    (lb * torch.tensor([1.0, -0.5])).sum().backward()
    for k, p in m.named_parameters():
        if p.grad is not None:
            out["rat__grad__" + k] = p.grad.numpy().copy()

    torch.manual_seed(12)
    np.random.seed(12)
    c = cspn.CspnConfig()
    c.F_cond = (5,)
    c.F, c.R, c.D, c.I, c.S, c.C = 8, 2, 2, 3, 3, 1
    c.dropout, c.leaf_base_class, c.tanh_squash = 0.0, dists.RatNormal, True
    c.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
    c.feat_layers, c.sum_param_layers, c.dist_param_layers = [8], [8], [8]
    c.cond_layers_inner_act = torch.nn.LeakyReLU
    cm = cspn.CSPN(c)
    for k, v in cm.state_dict().items():
        out["cspn__sd__" + k] = v.detach().numpy().copy()
    cond = torch.randn(4, 5)
    out["cspn__cond"] = cond.numpy()
    lb, _ = cm.huber_entropy_lb(condition=cond, verbose=False)
    out["cspn__lb"] = lb.detach().numpy()
    (-lb.mean()).backward()
    for k, p in cm.named_parameters():
        if p.grad is not None:
            out["cspn__grad__" + k] = p.grad.numpy().copy()
    # oracle vs reference on the stored quantities (appended to PIN_REPORT.txt)
    from oracle import spn_oracle as O
    arch = O.Arch(F=16, D=3, R=3, S=3, I=4, C=2)
    sd = {k[len("rat__sd__"):]: torch.from_numpy(v) for k, v in out.items() if k.startswith("rat__sd__")}
    P = O.params_from_ratspn_state(sd, arch, min_sigma=0.05, max_sigma=1.5, dtype=torch.float64)
    lines = []
    checks = [(f"huber lb layer {li}", O.huber_entropy_lb(P, arch, layer_index=li), out[f"rat__lb_layer{li}"]) for li in (2, 4, 6)]
    checks.append(("huber lb marginal_mask", O.huber_entropy_lb(P, arch, marginal_mask=mask), out["rat__lb_mask"]))
    checks.append(("huber lb detach + add_sub", O.huber_entropy_lb(P, arch, detach_weights=True, add_sub_weight_ent=True),
                   out["rat__lb_flags"]))
    for what, got, ref in checks:
        ref = torch.from_numpy(np.asarray(ref)).double()
        d = float((got - ref).abs().max())
        lines.append(f"{'ratspn_huber':18s} {what:34s} max|d|={d:.3e} rel={d / float(ref.abs().max()):.3e} shape={tuple(ref.shape)}")
    for ln in lines:
        print(ln)
    report = os.path.join(ROOT, "tests", "golden", "PIN_REPORT.txt")
    old = [l for l in open(report).read().splitlines() if not l.startswith("ratspn_huber")] if os.path.exists(report) else []
    with open(report, "w") as f:
        f.write("\n".join(old + lines) + "\n")
    path = os.path.join(ROOT, "tests", "golden", "huber.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if "lb" in k})


if __name__ == "__main__":
    main()
