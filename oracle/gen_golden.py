"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE.  Usage (from the repo root, in the container that has /root/reference):

    python oracle/gen_golden.py

For every case it stores the model ``state_dict``, the seeded inputs, the noise tensors the reference drew
(in draw order) and the reference's outputs / autograd gradients.  ``tests/test_oracle_golden.py`` replays them
through ``oracle/spn_oracle.py``; the GPU tests replay them through the CUDA path.  The script also prints the
oracle-vs-reference deviation of every stored quantity and writes it to tests/golden/PIN_REPORT.txt.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import spn_oracle as O  # noqa: E402
from oracle.ref_shim import load_reference, recorded_noise  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REPORT = []


def note(case, what, ref, got):
    ref = torch.as_tensor(ref)
    got = torch.as_tensor(got)
    if ref.dtype.is_floating_point:
        d = (ref.double() - got.double()).abs().max().item() if ref.numel() else 0.0
        rel = d / max(ref.double().abs().max().item(), 1e-30) if ref.numel() else 0.0
        line = f"{case:18s} {what:34s} max|d|={d:.3e} rel={rel:.3e} shape={tuple(ref.shape)}"
    else:
        mism = int((ref != got).sum())
        line = f"{case:18s} {what:34s} mismatches={mism} of {ref.numel()} shape={tuple(ref.shape)}"
    print(line)
    REPORT.append(line)


def sd_arrays(model):
    return {"sd__" + k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}


def grads_of(model):
    return {"grad__" + k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters()
            if p.requires_grad and p.grad is not None}


def zero_grads(model):
    for p in model.parameters():
        p.grad = None


def noise_arrays(prefix, rec):
    return {f"{prefix}__noise{i:02d}": t.numpy() for i, t in enumerate(rec)}


def save(name, arrays):
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def oracle_params_ratspn(model, arch, leaf_kwargs, dtype=torch.float32, requires_grad=False):
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    raw = {}
    if requires_grad:
        for k in list(sd):
            if k.endswith("weight_param") or k.endswith("mean_param") or k.endswith("std_param"):
                if not k.startswith("_sampling_root"):
                    sd[k] = sd[k].clone().requires_grad_(True)
                    raw[k] = sd[k]
    P = O.params_from_ratspn_state(sd, arch, min_sigma=leaf_kwargs.get("min_sigma", 0.0),
                                   max_sigma=leaf_kwargs.get("max_sigma", 1.0),
                                   min_mean=leaf_kwargs.get("min_mean"), max_mean=leaf_kwargs.get("max_mean"),
                                   dtype=dtype)
    return P, raw


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref_rat_spn, ref_cspn, ref_layers, ref_dist, ref_utils = load_reference()
    RatSpn, RatSpnConfig = ref_rat_spn.RatSpn, ref_rat_spn.RatSpnConfig
    RatNormal = ref_dist.RatNormal

    def build_ratspn(seed, F, D, R, S, I, C, tanh, leaf_kwargs):
        torch.manual_seed(seed)
        np.random.seed(seed)
        c = RatSpnConfig()
        c.F, c.D, c.R, c.S, c.I, c.C = F, D, R, S, I, C
        c.dropout = 0.0
        c.leaf_base_class = RatNormal
        c.leaf_base_kwargs = dict(leaf_kwargs)
        c.tanh_squash = tanh
        return RatSpn(c)

    # ------------------------------------------------------------------ case 1: small RatSpn, C=2, no squash
    case = "ratspn_small"
    lk = {}
    m = build_ratspn(1, 16, 2, 3, 4, 5, 2, False, lk)
    arch = O.Arch(F=16, D=2, R=3, S=4, I=5, C=2, tanh_squash=False)
    A = dict(sd_arrays(m))
    A["meta"] = np.array([16, 2, 3, 4, 5, 2, 0])
    x = torch.randn(7, 1, 16, 1, 1)
    x_nan = x.clone()
    x_nan[0, 0, 3] = float("nan")
    x_nan[2, 0, :] = float("nan")
    x_nan[5, 0, 10:] = float("nan")
    A["x"], A["x_nan"] = x.numpy(), x_nan.numpy()
    P, _ = oracle_params_ratspn(m, arch, lk)
    for li in range(0, 5):
        out = m(x, layer_index=li).detach()
        A[f"ll_layer{li}"] = out.numpy()
        note(case, f"forward layer_index={li}", out, O.forward(P, arch, x, layer_index=li))
    out = m(x_nan).detach()
    A["ll_nan"] = out.numpy()
    note(case, "forward with NaN", out, O.forward(P, arch, x_nan))
    zero_grads(m)
    (-m(x_nan).mean()).backward()
    A.update(grads_of(m))
    Pg, raw = oracle_params_ratspn(m, arch, lk, requires_grad=True)
    (-O.forward(Pg, arch, x_nan).mean()).backward()
    for k, t in raw.items():
        note(case, f"grad {k}", A["grad__" + k], t.grad)
    # x that already carries I and R dims, no permutation, partial forward (entropy call pattern rat_spn.py:689)
    x_ir = torch.randn(3, 1, 16, 5, 3)
    A["x_ir"] = x_ir.numpy()
    out = m(x_ir, layer_index=2, x_needs_permutation=False).detach()
    A["ll_ir_layer2"] = out.numpy()
    note(case, "forward x[I,R] layer 2", out, O.forward(P, arch, x_ir, layer_index=2, x_needs_permutation=False))
    # sampling
    for tag, kw in [("samp", dict(n=5)), ("mpe", dict(n=2, is_mpe=True)),
                    ("cls", dict(n=3, class_index=torch.tensor([1])))]:
        rec = []
        torch.manual_seed(11)
        with recorded_noise(ref_layers, ref_dist, rec):
            ctx = m.sample(mode="index", **kw)
        A.update(noise_arrays(tag, rec))
        A[f"{tag}__sample"] = ctx.sample.detach().numpy()
        A[f"{tag}__rep"] = ctx.repetition_indices.numpy()
        A[f"{tag}__parent"] = ctx.parent_indices.numpy()
        o = O.sample_index(P, arch, O.NoiseSource(replay=rec), n=kw["n"], class_index=kw.get("class_index"),
                           is_mpe=kw.get("is_mpe", False))
        note(case, f"{tag} sample", ctx.sample.detach(), o.sample)
        note(case, f"{tag} repetition idx", ctx.repetition_indices, o.repetition_indices)
        note(case, f"{tag} leaf channel idx", ctx.parent_indices, o.parent_indices)
    save(case, A)

    # ------------------------------------------------------------------ case 2: padded, tanh-squashed RatSpn
    case = "ratspn_pad_tanh"
    lk = {"min_sigma": 0.001, "max_sigma": 1.0}
    m = build_ratspn(2, 17, 4, 3, 3, 3, 1, True, lk)
    arch = O.Arch(F=17, D=4, R=3, S=3, I=3, C=1, tanh_squash=True)
    A = dict(sd_arrays(m))
    A["meta"] = np.array([17, 4, 3, 3, 3, 1, 1])
    P, _ = oracle_params_ratspn(m, arch, lk)
    x = torch.atanh(torch.rand(6, 1, 17, 1, 1) * 1.8 - 0.9)
    x[1, 0, 2:5] = float("nan")
    A["x"] = x.numpy()
    for li in (0, 1, 2, 5, 8):
        out = m(x, layer_index=li).detach()
        A[f"ll_layer{li}"] = out.numpy()
        note(case, f"forward layer_index={li}", out, O.forward(P, arch, x, layer_index=li))
    zero_grads(m)
    (-m(x).mean()).backward()
    A.update(grads_of(m))
    Pg, raw = oracle_params_ratspn(m, arch, lk, requires_grad=True)
    (-O.forward(Pg, arch, x).mean()).backward()
    for k, t in raw.items():
        note(case, f"grad {k}", A["grad__" + k], t.grad)
    ev = torch.atanh(torch.rand(2, 1, 17, 1, 1) * 1.8 - 0.9)
    ev[0, 0, ::2] = float("nan")
    ev[1, 0, 5:] = float("nan")
    A["evidence"] = ev.numpy()
    for tag, kw in [("ev", dict(n=3, evidence=ev)), ("evmpe", dict(n=1, evidence=ev, is_mpe=True)),
                    ("l3", dict(n=2, layer_index=3)), ("l2", dict(n=2, layer_index=2)),
                    ("l7raw", dict(n=2, layer_index=7, do_sample_postprocessing=False)),
                    ("l0raw", dict(n=3, layer_index=0, do_sample_postprocessing=False))]:
        rec = []
        torch.manual_seed(12)
        with recorded_noise(ref_layers, ref_dist, rec):
            ctx = m.sample(mode="index", **kw)
        A.update(noise_arrays(tag, rec))
        A[f"{tag}__sample"] = ctx.sample.detach().numpy()
        o = O.sample_index(P, arch, O.NoiseSource(replay=rec), n=kw["n"], evidence=kw.get("evidence"),
                           is_mpe=kw.get("is_mpe", False), layer_index=kw.get("layer_index"),
                           postprocess=kw.get("do_sample_postprocessing", True))
        note(case, f"{tag} sample", ctx.sample.detach(), o.sample)
        if ctx.repetition_indices is not None:
            A[f"{tag}__rep"] = ctx.repetition_indices.numpy()
            note(case, f"{tag} repetition idx", ctx.repetition_indices, o.repetition_indices)
    # onehot == index under MPE (reference test invariant tests/test_RatSpn.py:21-24)
    oh = m.sample(mode="onehot", is_mpe=True).sample.detach()
    ix = m.sample(mode="index", is_mpe=True).sample.detach()
    note(case, "reference onehot-vs-index MPE", ix, oh)
    # recursive entropy: no-grad / index mode, with and without marginal mask
    mask = torch.zeros(1, 17, dtype=torch.bool)
    mask[0, [1, 8, 16]] = True
    A["marginal_mask"] = mask.numpy()
    for tag, mm in [("ent", None), ("entmask", mask)]:
        rec = []
        torch.manual_seed(13)
        with torch.no_grad(), recorded_noise(ref_layers, ref_dist, rec):
            H, _ = m.recursive_entropy_approx(sample_size=4, marginal_mask=mm)
        A.update(noise_arrays(tag, rec))
        A[f"{tag}__H"] = H.numpy()
        o = O.recursive_entropy(P, arch, O.NoiseSource(replay=rec), sample_size=4, marginal_mask=mm)
        note(case, f"{tag} entropy (no grad)", H, o)
    # recursive entropy with gradients (onehot sampling), both aux settings
    for tag, awg in [("entgrad", False), ("entgradaux", True)]:
        rec = []
        torch.manual_seed(14)
        zero_grads(m)
        with recorded_noise(ref_layers, ref_dist, rec):
            H, _ = m.recursive_entropy_approx(sample_size=4, aux_with_grad=awg, marginal_mask=mask)
        (-H.sum()).backward()
        A.update(noise_arrays(tag, rec))
        A[f"{tag}__H"] = H.detach().numpy()
        g = grads_of(m)
        A.update({f"{tag}__{k}": v for k, v in g.items()})
        Pg, raw = oracle_params_ratspn(m, arch, lk, requires_grad=True)
        o = O.recursive_entropy(Pg, arch, O.NoiseSource(replay=rec), sample_size=4, aux_with_grad=awg,
                                marginal_mask=mask, differentiable=True)
        (-o.sum()).backward()
        note(case, f"{tag} entropy", H.detach(), o.detach())
        for k, t in raw.items():
            note(case, f"{tag} grad {k}", g["grad__" + k], t.grad)
    # differentiable sample with evidence
    rec = []
    torch.manual_seed(15)
    zero_grads(m)
    with recorded_noise(ref_layers, ref_dist, rec):
        ctx = m.sample(mode="onehot", n=2, evidence=ev)
    cw = torch.linspace(-1.0, 1.0, ctx.sample.numel()).view_as(ctx.sample)
    (ctx.sample * cw).sum().backward()
    A.update(noise_arrays("ohev", rec))
    A["ohev__sample"] = ctx.sample.detach().numpy()
    g = grads_of(m)
    A.update({f"ohev__{k}": v for k, v in g.items()})
    Pg, raw = oracle_params_ratspn(m, arch, lk, requires_grad=True)
    o = O.sample_onehot(Pg, arch, O.NoiseSource(replay=rec), n=2, evidence=ev)
    (o.sample * cw).sum().backward()
    note(case, "onehot+evidence sample", ctx.sample.detach(), o.sample.detach())
    for k, t in raw.items():
        note(case, f"ohev grad {k}", g["grad__" + k], t.grad if t.grad is not None else torch.zeros_like(t))
    save(case, A)

    # ------------------------------------------------------------------ case 3: small CSPN (SAC-policy shaped)
    case = "cspn_small"
    CSPN, CspnConfig = ref_cspn.CSPN, ref_cspn.CspnConfig
    torch.manual_seed(3)
    np.random.seed(3)
    c = CspnConfig()
    c.F_cond = (6,)
    c.C = 1
    c.feat_layers = [8]
    c.sum_param_layers = [8]
    c.dist_param_layers = [8]
    c.F, c.R, c.D, c.I, c.S = 9, 2, 3, 3, 2
    c.dropout = 0.0
    c.leaf_base_class = RatNormal
    c.tanh_squash = True
    c.cond_layers_inner_act = torch.nn.LeakyReLU
    c.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
    m = CSPN(c)
    arch = O.Arch(F=9, D=3, R=2, S=2, I=3, C=1, tanh_squash=True)
    A = dict(sd_arrays(m))
    A["meta"] = np.array([9, 3, 2, 2, 3, 1, 1, 6])
    B = 4
    cond = torch.randn(B, 6)
    x = torch.atanh(torch.rand(3, B, 9, 1, 1) * 1.8 - 0.9)
    A["cond"], A["x"] = cond.numpy(), x.numpy()

    def cspn_oracle_params(requires_grad=False):
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        raw = {}
        if requires_grad:
            for k, p in m.named_parameters():
                if p.requires_grad:
                    sd[k] = sd[k].requires_grad_(True)
                    raw[k] = sd[k]
        return O.cspn_params(sd, arch, cond, min_sigma=0.001, max_sigma=1.0), raw

    zero_grads(m)
    ll = m(x, condition=cond)
    (-ll.mean()).backward()
    A["ll"] = ll.detach().numpy()
    g = grads_of(m)
    A.update({f"ll__{k}": v for k, v in g.items()})
    Pc, raw = cspn_oracle_params(True)
    o = O.forward(Pc, arch, x)
    (-o.mean()).backward()
    note(case, "forward", ll.detach(), o.detach())
    for k, t in raw.items():
        note(case, f"ll grad {k}", g["grad__" + k], t.grad)
    for k in ("mu", "sigma", "lw_root"):
        A["setparams__" + k] = Pc[k].detach().numpy()
    ref_w = m.root.weight_param.detach()
    note(case, "set_params root weights", ref_w, Pc["lw_root"].detach())
    note(case, "set_params means", m._leaf.base_leaf.mean_param.detach(), Pc["mu"].detach())
    note(case, "set_params stds", m._leaf.base_leaf.std_param.detach(), Pc["sigma"].detach())
    # SAC call pattern: sample with evidence + recursive entropy with marginal mask (sb3.py:125-162)
    failed = torch.zeros(B, 9, dtype=torch.bool)
    failed[0, 2] = failed[1, 0] = failed[1, 7] = failed[3, 4] = True
    ev = torch.atanh(torch.rand(1, B, 9, 1, 1) * 1.8 - 0.9)
    ev[~failed.view(1, B, 9, 1, 1)] = float("nan")
    A["failed"], A["evidence"] = failed.numpy(), ev.numpy()
    rec = []
    torch.manual_seed(16)
    with torch.no_grad(), recorded_noise(ref_layers, ref_dist, rec):
        ctx = m.sample(mode="index", condition=cond, evidence=ev)
    A.update(noise_arrays("ixev", rec))
    A["ixev__sample"] = ctx.sample.numpy()
    Pc0, _ = cspn_oracle_params(False)
    o = O.sample_index(Pc0, arch, O.NoiseSource(replay=rec), n=1, evidence=ev)
    note(case, "index sample + evidence", ctx.sample, o.sample)
    rec = []
    torch.manual_seed(17)
    zero_grads(m)
    with recorded_noise(ref_layers, ref_dist, rec):
        ctx = m.sample(mode="onehot", condition=cond, evidence=ev)
        H, _ = m.recursive_entropy_approx(condition=cond, sample_size=3, marginal_mask=failed)
    loss = -(0.1 * H + ctx.sample.squeeze(0).squeeze(0).squeeze(-1).sum(-1)).mean()
    loss.backward()
    A.update(noise_arrays("sac", rec))
    A["sac__sample"] = ctx.sample.detach().numpy()
    A["sac__H"] = H.detach().numpy()
    g = grads_of(m)
    A.update({f"sac__{k}": v for k, v in g.items()})
    Pc, raw = cspn_oracle_params(True)
    ns = O.NoiseSource(replay=rec)
    os_ = O.sample_onehot(Pc, arch, ns, n=1, evidence=ev)
    oH = O.recursive_entropy(Pc, arch, ns, sample_size=3, marginal_mask=failed, differentiable=True)
    (-(0.1 * oH + os_.sample.squeeze(0).squeeze(0).squeeze(-1).sum(-1)).mean()).backward()
    note(case, "SAC sample", ctx.sample.detach(), os_.sample.detach())
    note(case, "SAC entropy", H.detach(), oH.detach())
    for k, t in raw.items():
        note(case, f"SAC grad {k}", g["grad__" + k], t.grad)
    save(case, A)

    # ------------------------------------------------------------------ case 4: BASELINE config 1 (MNIST-shaped), 4 rows
    case = "ratspn_cfg1"
    lk = {}
    m = build_ratspn(0, 784, 3, 5, 10, 10, 1, False, lk)
    arch = O.Arch(F=784, D=3, R=5, S=10, I=10, C=1)
    A = dict(sd_arrays(m))
    A["meta"] = np.array([784, 3, 5, 10, 10, 1, 0])
    x = torch.rand(4, 1, 784, 1, 1)
    A["x"] = x.numpy()
    zero_grads(m)
    ll = m(x)
    (-ll.mean()).backward()
    A["ll"] = ll.detach().numpy()
    A.update(grads_of(m))
    Pg, raw = oracle_params_ratspn(m, arch, lk, requires_grad=True)
    o = O.forward(Pg, arch, x)
    (-o.mean()).backward()
    note(case, "forward", ll.detach(), o.detach())
    for k, t in raw.items():
        note(case, f"grad {k}", A["grad__" + k], t.grad)
    rec = []
    torch.manual_seed(18)
    with recorded_noise(ref_layers, ref_dist, rec):
        ctx = m.sample(mode="index", n=3)
    # the root Gumbel tensor is small here (R*S^2 = 500)
    A.update(noise_arrays("samp", rec))
    A["samp__sample"] = ctx.sample.detach().numpy()
    A["samp__rep"] = ctx.repetition_indices.numpy()
    P, _ = oracle_params_ratspn(m, arch, lk)
    o = O.sample_index(P, arch, O.NoiseSource(replay=rec), n=3)
    note(case, "sample", ctx.sample.detach(), o.sample)
    save(case, A)

    with open(os.path.join(GOLD, "PIN_REPORT.txt"), "w") as f:
        f.write("oracle/spn_oracle.py vs the unmodified reference (/root/reference), generated by oracle/gen_golden.py\n")
        f.write(f"torch {torch.__version__}, numpy {np.__version__}\n\n")
        f.write("\n".join(REPORT) + "\n")


if __name__ == "__main__":
    main()
