"""CPU: the oracle restatement against the golden vectors produced by the unmodified reference
(oracle/gen_golden.py).  Forward / gradients / sampling are expected bit-exact in fp32 on the same torch build and
within float re-association tolerance otherwise; entropy within 1e-5 relative."""
import pytest
import torch

from oracle import spn_oracle as O
from tests.helpers import Case, assert_close, oracle_params

RT, AT = 2e-5, 2e-5


@pytest.mark.parametrize("name,layers", [("ratspn_small", [0, 1, 2, 3, 4]), ("ratspn_pad_tanh", [0, 1, 2, 5, 8])])
def test_forward_layers(name, layers):
    c = Case(name)
    P, _ = oracle_params(c)
    for li in layers:
        assert_close(O.forward(P, c.arch, c.t("x"), layer_index=li), c.t(f"ll_layer{li}"), RT, AT, f"layer {li}")


def test_forward_nan_marginalisation_and_full_nan_is_zero():
    c = Case("ratspn_small")
    P, _ = oracle_params(c)
    out = O.forward(P, c.arch, c.t("x_nan"))
    assert_close(out, c.t("ll_nan"), RT, AT, "NaN input")
    assert out[2].abs().max() < 1e-5  # row 2 is all-NaN: fully marginalised => LL = 0 (invariant 3)


def test_forward_input_with_channel_and_repetition_dims():
    c = Case("ratspn_small")
    P, _ = oracle_params(c)
    out = O.forward(P, c.arch, c.t("x_ir"), layer_index=2, x_needs_permutation=False)
    assert_close(out, c.t("ll_ir_layer2"), RT, AT)


@pytest.mark.parametrize("name,xkey", [("ratspn_small", "x_nan"), ("ratspn_pad_tanh", "x"), ("ratspn_cfg1", "x")])
def test_backward(name, xkey):
    c = Case(name)
    P, raw = oracle_params(c, requires_grad=True)
    (-O.forward(P, c.arch, c.t(xkey)).mean()).backward()
    ref = c.grads()
    assert set(raw) == set(ref)
    for k, t in raw.items():
        assert_close(t.grad, ref[k], 1e-4, 1e-7, f"grad {k}")


@pytest.mark.parametrize("tag,kw", [("samp", dict(n=5)), ("mpe", dict(n=2, is_mpe=True)),
                                    ("cls", dict(n=3, class_index=torch.tensor([1])))])
def test_index_sampling_bit_exact(tag, kw):
    c = Case("ratspn_small")
    P, _ = oracle_params(c)
    o = O.sample_index(P, c.arch, O.NoiseSource(replay=c.noise(tag)), **kw)
    assert torch.equal(o.repetition_indices, c.t(f"{tag}__rep"))
    assert torch.equal(o.parent_indices, c.t(f"{tag}__parent"))
    assert_close(o.sample, c.t(f"{tag}__sample"), 1e-6, 1e-6)


@pytest.mark.parametrize("tag,kw", [("ev", dict(n=3, evidence=True)), ("evmpe", dict(n=1, evidence=True, is_mpe=True)),
                                    ("l3", dict(n=2, layer_index=3)), ("l2", dict(n=2, layer_index=2)),
                                    ("l7raw", dict(n=2, layer_index=7, postprocess=False)),
                                    ("l0raw", dict(n=3, layer_index=0, postprocess=False))])
def test_sampling_evidence_and_inner_layers(tag, kw):
    c = Case("ratspn_pad_tanh")
    P, _ = oracle_params(c)
    kw = dict(kw)
    if kw.pop("evidence", False):
        kw["evidence"] = c.t("evidence")
    o = O.sample_index(P, c.arch, O.NoiseSource(replay=c.noise(tag)), **kw)
    assert_close(o.sample, c.t(f"{tag}__sample"), 1e-5, 1e-6)
    if "evidence" in kw:  # evidence is preserved (invariant 6)
        ev = torch.tanh(c.t("evidence").clamp(-6, 6)).expand(kw["n"], *c.t("evidence").shape).movedim(-2, 0)
        keep = ~torch.isnan(ev)
        assert torch.equal(o.sample[keep], ev[keep])


@pytest.mark.parametrize("tag,mask", [("ent", False), ("entmask", True)])
def test_recursive_entropy_no_grad(tag, mask):
    c = Case("ratspn_pad_tanh")
    P, _ = oracle_params(c)
    mm = c.t("marginal_mask") if mask else None
    H = O.recursive_entropy(P, c.arch, O.NoiseSource(replay=c.noise(tag)), sample_size=4, marginal_mask=mm)
    assert_close(H, c.t(f"{tag}__H"), 1e-5, 1e-5)


@pytest.mark.parametrize("tag,aux", [("entgrad", False), ("entgradaux", True)])
def test_recursive_entropy_gradients(tag, aux):
    c = Case("ratspn_pad_tanh")
    P, raw = oracle_params(c, requires_grad=True)
    H = O.recursive_entropy(P, c.arch, O.NoiseSource(replay=c.noise(tag)), sample_size=4, aux_with_grad=aux,
                            marginal_mask=c.t("marginal_mask"), differentiable=True)
    (-H.sum()).backward()
    assert_close(H, c.t(f"{tag}__H"), 1e-5, 1e-5)
    ref = c.grads(f"{tag}__grad__")
    for k, t in raw.items():
        assert_close(t.grad, ref[k], 1e-3, 1e-5, f"{tag} grad {k}")


def test_onehot_sampling_with_evidence_gradients():
    c = Case("ratspn_pad_tanh")
    P, raw = oracle_params(c, requires_grad=True)
    o = O.sample_onehot(P, c.arch, O.NoiseSource(replay=c.noise("ohev")), n=2, evidence=c.t("evidence"))
    ref = c.t("ohev__sample")
    assert_close(o.sample, ref, 1e-6, 1e-6)
    cw = torch.linspace(-1.0, 1.0, ref.numel()).view_as(ref)
    (o.sample * cw).sum().backward()
    g = c.grads("ohev__grad__")
    for k, t in raw.items():
        assert_close(t.grad if t.grad is not None else torch.zeros_like(t), g[k], 1e-4, 1e-6, f"ohev grad {k}")


def _cspn_params(c, requires_grad):
    sd = c.state_dict()
    raw = {}
    if requires_grad:
        for k in list(sd):
            if sd[k].dtype.is_floating_point and (k.endswith(".weight") or k.endswith(".bias")):
                sd[k] = sd[k].clone().requires_grad_(True)
                raw[k] = sd[k]
    return O.cspn_params(sd, c.arch, c.t("cond"), min_sigma=0.001, max_sigma=1.0), raw


def test_cspn_set_params_forward_and_gradients():
    c = Case("cspn_small")
    P, raw = _cspn_params(c, True)
    for k in ("mu", "sigma", "lw_root"):
        assert_close(P[k], c.t("setparams__" + k), 1e-6, 1e-6, k)
    assert_close(P["lw_root"].exp().sum(2), torch.ones(4, 1, 1, 1), 1e-5, 1e-5, "weights normalised")
    out = O.forward(P, c.arch, c.t("x"))
    assert_close(out, c.t("ll"), RT, AT)
    (-out.mean()).backward()
    g = c.grads("ll__grad__")
    for k, t in raw.items():
        assert_close(t.grad, g[k], 1e-4, 1e-7, f"ll grad {k}")


def test_cspn_sac_update_pattern():
    c = Case("cspn_small")
    P0, _ = _cspn_params(c, False)
    o = O.sample_index(P0, c.arch, O.NoiseSource(replay=c.noise("ixev")), n=1, evidence=c.t("evidence"))
    assert_close(o.sample, c.t("ixev__sample"), 1e-6, 1e-6)
    P, raw = _cspn_params(c, True)
    ns = O.NoiseSource(replay=c.noise("sac"))
    s = O.sample_onehot(P, c.arch, ns, n=1, evidence=c.t("evidence"))
    H = O.recursive_entropy(P, c.arch, ns, sample_size=3, marginal_mask=c.t("failed"), differentiable=True)
    assert_close(s.sample, c.t("sac__sample"), 1e-6, 1e-6)
    assert_close(H, c.t("sac__H"), 1e-5, 1e-5)
    (-(0.1 * H + s.sample.squeeze(0).squeeze(0).squeeze(-1).sum(-1)).mean()).backward()
    g = c.grads("sac__grad__")
    for k, t in raw.items():
        assert_close(t.grad, g[k], 1e-3, 1e-6, f"sac grad {k}")


def test_cspn_at_baseline_config3_shape():
    """BASELINE config 3 at its real shape (obs 376 / 17 actions, R = I = S = 3, D = 4, [64] hidden; fixture recorded
    from the unmodified reference by oracle/gen_golden_cfg3.py): LL + gradients, rollout sampling with evidence (incl.
    MPE), and the policy-update pattern (onehot sample + recursive entropy with the failed-joint mask, one backward)."""
    c = Case("cspn_cfg3")
    P, raw = _cspn_params(c, True)
    out = O.forward(P, c.arch, c.t("x"))
    assert_close(out, c.t("ll"), RT, AT)
    (-out.mean()).backward()
    g = c.grads("ll__grad__")
    for k, t in raw.items():
        assert_close(t.grad, g[k], 1e-4, 1e-7, f"ll grad {k}")
    P0, _ = _cspn_params(c, False)
    for tag, mpe in (("ixev", False), ("ixevmpe", True)):
        o = O.sample_index(P0, c.arch, O.NoiseSource(replay=c.noise(tag)), n=1, evidence=c.t("evidence"), is_mpe=mpe)
        assert_close(o.sample, c.t(f"{tag}__sample"), 1e-6, 1e-6, tag)
    P, raw = _cspn_params(c, True)
    ns = O.NoiseSource(replay=c.noise("sac"))
    s = O.sample_onehot(P, c.arch, ns, n=1, evidence=c.t("evidence"))
    H = O.recursive_entropy(P, c.arch, ns, sample_size=5, marginal_mask=c.t("failed"), differentiable=True)
    assert_close(s.sample, c.t("sac__sample"), 1e-6, 1e-6)
    assert_close(H, c.t("sac__H"), 1e-5, 1e-5)
    (-(0.1 * H + s.sample.squeeze(0).squeeze(0).squeeze(-1).sum(-1)).mean()).backward()
    g = c.grads("sac__grad__")
    for k, t in raw.items():
        scale = float(g[k].abs().max())
        assert_close(t.grad, g[k], 0.0, 1e-4 * scale + 1e-9, f"sac grad {k}")


def test_structural_invariants():
    c = Case("ratspn_small")
    P, _ = oracle_params(c)
    inv = O.invert_permutation(P["perm"])
    x = torch.randn(3, 1, c.F, 1, c.R)
    back = torch.gather(O.apply_permutation(x, P["perm"]), -3, inv.unsqueeze(-2).expand(3, 1, c.F, 1, c.R))
    assert torch.equal(back, x)  # inverse permutation undoes the permutation (invariant 2)
    for lw in P["lw"] + [P["lw_root"]]:
        assert torch.allclose(lw.exp().sum(2), torch.ones(()), atol=1e-5)  # invariant 4
    assert O.forward(P, c.arch, c.t("x")).shape == (7, 1, 1, c.C, 1)  # invariant 7
