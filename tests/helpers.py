"""Shared helpers of the test-suite: golden fixture loading and model construction from a fixture."""
import os

import numpy as np
import torch

from oracle import spn_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Case:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLD, name + ".npz"))
        m = [int(v) for v in self.z["meta"]]
        self.F, self.D, self.R, self.S, self.I, self.C, tanh = m[:7]
        self.tanh = bool(tanh)
        self.F_cond = m[7] if len(m) > 7 else None
        self.hidden = m[8] if len(m) > 8 else 8   # width of the CSPN's feat / sum / dist hidden layers
        self.arch = O.Arch(F=self.F, D=self.D, R=self.R, S=self.S, I=self.I, C=self.C, tanh_squash=self.tanh)

    def t(self, key):
        return torch.from_numpy(self.z[key])

    def has(self, key):
        return key in self.z.files

    def state_dict(self):
        return {k[4:]: torch.from_numpy(self.z[k]) for k in self.z.files if k.startswith("sd__")}

    def noise(self, tag):
        keys = sorted(k for k in self.z.files if k.startswith(tag + "__noise"))
        return [torch.from_numpy(self.z[k]) for k in keys]

    def grads(self, prefix="grad__"):
        return {k[len(prefix):]: torch.from_numpy(self.z[k]) for k in self.z.files if k.startswith(prefix)}


LEAF_KWARGS = {"ratspn_small": {}, "ratspn_cfg1": {}, "ratspn_pad_tanh": {"min_sigma": 0.001, "max_sigma": 1.0},
               "cspn_small": {"min_sigma": 0.001, "max_sigma": 1.0}, "cspn_cfg3": {"min_sigma": 0.001, "max_sigma": 1.0}}


def oracle_params(case: Case, requires_grad=False, dtype=torch.float32):
    sd = case.state_dict()
    raw = {}
    if requires_grad:
        for k in list(sd):
            if k.split(".")[-1] in ("weight_param", "mean_param", "std_param") and not k.startswith("_sampling_root"):
                sd[k] = sd[k].clone().requires_grad_(True)
                raw[k] = sd[k]
    lk = LEAF_KWARGS[case.name]
    P = O.params_from_ratspn_state(sd, case.arch, min_sigma=lk.get("min_sigma", 0.0), max_sigma=lk.get("max_sigma", 1.0),
                                   dtype=dtype)
    return P, raw


def build_ratspn(case: Case, device):
    import ratspn_b200 as rb
    cfg = rb.RatSpnConfig()
    cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = case.F, case.D, case.R, case.S, case.I, case.C
    cfg.dropout = 0.0
    cfg.leaf_base_class = rb.RatNormal
    cfg.leaf_base_kwargs = dict(LEAF_KWARGS[case.name])
    cfg.tanh_squash = case.tanh
    model = rb.RatSpn(cfg)
    missing = model.load_state_dict(case.state_dict(), strict=True)
    return model.to(device)


def build_cspn(case: Case, device):
    import ratspn_b200 as rb
    cfg = rb.CspnConfig()
    cfg.F_cond = (case.F_cond,)
    cfg.C = case.C
    cfg.feat_layers = [case.hidden]
    cfg.sum_param_layers = [case.hidden]
    cfg.dist_param_layers = [case.hidden]
    cfg.F, cfg.R, cfg.D, cfg.I, cfg.S = case.F, case.R, case.D, case.I, case.S
    cfg.dropout = 0.0
    cfg.leaf_base_class = rb.RatNormal
    cfg.tanh_squash = case.tanh
    cfg.cond_layers_inner_act = torch.nn.LeakyReLU
    cfg.leaf_base_kwargs = dict(LEAF_KWARGS[case.name])
    model = rb.CSPN(cfg)
    model.load_state_dict(case.state_dict(), strict=True)
    return model.to(device)


def assert_close(got, ref, rtol, atol, what=""):
    got = got.detach().cpu().double()
    ref = ref.detach().cpu().double()
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    both_inf = torch.isinf(got) & torch.isinf(ref) & (got == ref)
    err = torch.where(both_inf, torch.zeros_like(got), (got - ref).abs())
    tol = atol + rtol * ref.abs()
    tol = torch.where(torch.isfinite(tol), tol, torch.full_like(tol, atol))
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())} of {bad.numel()} elements differ; max abs err "
                           f"{float(err.max()):.3e}, max |ref| {float(ref.abs().max()):.3e}")
