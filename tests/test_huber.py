"""``huber_entropy_lb`` (entropy lower bound of Huber et al. 2008; rat_spn.py:825-951, cspn.py:112-115) against golden
vectors recorded from the unmodified reference (oracle/gen_golden_huber.py -> tests/golden/huber.npz).  The bound only
combines parameters with tensor ops (no circuit evaluation), so it is checked on the host as well as on the GPU.
Tolerances: values rtol 2e-5 / atol 2e-5; gradients 2e-3 of the largest magnitude per tensor."""
import os

import numpy as np
import pytest
import torch

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "huber.npz"))


def _t(key, device="cpu"):
    return torch.from_numpy(GOLD[key]).to(device)


def _ratspn(device):
    import ratspn_b200 as rb
    cfg = rb.RatSpnConfig()
    cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = 16, 3, 3, 3, 4, 2
    cfg.dropout, cfg.leaf_base_class, cfg.tanh_squash = 0.0, rb.RatNormal, False
    cfg.leaf_base_kwargs = {"min_sigma": 0.05, "max_sigma": 1.5}
    m = rb.RatSpn(cfg)
    m.load_state_dict({k[len("rat__sd__"):]: _t(k) for k in GOLD.files if k.startswith("rat__sd__")})
    return m.to(device)


def _cspn(device):
    import ratspn_b200 as rb
    c = rb.CspnConfig()
    c.F_cond = (5,)
    c.F, c.R, c.D, c.I, c.S, c.C = 8, 2, 2, 3, 3, 1
    c.dropout, c.leaf_base_class, c.tanh_squash = 0.0, rb.RatNormal, True
    c.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
    c.feat_layers, c.sum_param_layers, c.dist_param_layers = [8], [8], [8]
    c.cond_layers_inner_act = torch.nn.LeakyReLU
    m = rb.CSPN(c)
    m.load_state_dict({k[len("cspn__sd__"):]: _t(k) for k in GOLD.files if k.startswith("cspn__sd__")})
    return m.to(device)


def _close(got, key, what):
    ref = _t(key, got.device)
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    torch.testing.assert_close(got, ref, rtol=2e-5, atol=2e-5, msg=lambda s: f"{what}: {s}")


def _grads_close(model, prefix):
    for k, p in model.named_parameters():
        key = prefix + k
        if key not in GOLD.files:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        ref = _t(key, p.device)
        scale = float(ref.abs().max()) + 1e-12
        err = float((p.grad - ref).abs().max())
        assert err <= 2e-3 * scale + 1e-7, f"grad {k}: err {err:.3e} scale {scale:.3e}"


def _check(device):
    m = _ratspn(device)
    for li in (2, 4, 6):
        lb, log = m.huber_entropy_lb(layer_index=li, verbose=False)
        assert log == {}
        _close(lb, f"rat__lb_layer{li}", f"layer {li}")
    lb, log = m.huber_entropy_lb(verbose=True)
    _close(lb, "rat__lb_default", "default")
    assert sorted(log.keys()) == list(GOLD["rat__log_keys"])
    np.testing.assert_allclose([log[k] for k in sorted(log.keys())], GOLD["rat__log_vals"], rtol=1e-4, atol=1e-4)
    lb, _ = m.huber_entropy_lb(verbose=False, marginal_mask=_t("rat__mask", device))
    _close(lb, "rat__lb_mask", "marginal mask")
    lb, _ = m.huber_entropy_lb(verbose=False, detach_weights=True, add_sub_weight_ent=True)
    _close(lb, "rat__lb_flags", "detach + add_sub")
    m.zero_grad(set_to_none=True)
    lb, _ = m.huber_entropy_lb(verbose=False, add_sub_weight_ent=True)
    (lb * torch.tensor([1.0, -0.5], device=device)).sum().backward()
    _grads_close(m, "rat__grad__")
    with pytest.raises(AssertionError):
        m.huber_entropy_lb(marginal_mask=torch.zeros(2, 3, dtype=torch.bool, device=device))

    c = _cspn(device)
    lb, _ = c.huber_entropy_lb(condition=_t("cspn__cond", device), verbose=False)
    _close(lb, "cspn__lb", "cspn")
    (-lb.mean()).backward()
    _grads_close(c, "cspn__grad__")


def test_huber_entropy_lb_vs_reference_golden_host():
    _check("cpu")


@pytest.mark.gpu
def test_huber_entropy_lb_vs_reference_golden_gpu():
    _check("cuda")


def test_index_helpers_round_trip():
    from ratspn_b200 import utils
    shape = (3, 4, 5)
    for flat in (0, 7, 33, 59):
        idx = utils.flat_index_to_tensor_index(flat, shape)
        assert idx == list(np.unravel_index(flat, shape))
        assert utils.tensor_index_to_flat_index(idx, shape) == flat
    assert utils.flat_index_to_tensor_index(torch.tensor(33), shape) == [1, 2, 3]


def test_oracle_huber_bound_is_pinned_to_the_reference_and_checks_the_product_on_random_models():
    """oracle/spn_oracle.huber_entropy_lb (per-repetition-pair restatement) against the golden vectors of the
    unmodified reference, then product vs oracle on seeded random architectures (host tensors, fp32 vs fp64)."""
    from oracle import spn_oracle as O
    import ratspn_b200 as rb
    arch = O.Arch(F=16, D=3, R=3, S=3, I=4, C=2)
    sd = {k[len("rat__sd__"):]: _t(k) for k in GOLD.files if k.startswith("rat__sd__")}
    raw = {}
    for k in list(sd):
        if k.split(".")[-1] in ("weight_param", "mean_param", "std_param") and not k.startswith("_sampling_root"):
            sd[k] = sd[k].double().requires_grad_(True)
            raw[k] = sd[k]
    P = O.params_from_ratspn_state(sd, arch, min_sigma=0.05, max_sigma=1.5, dtype=torch.float64)
    for li in (2, 4, 6):
        got = O.huber_entropy_lb(P, arch, layer_index=li)
        torch.testing.assert_close(got.float(), _t(f"rat__lb_layer{li}"), rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(O.huber_entropy_lb(P, arch, marginal_mask=_t("rat__mask")).float(), _t("rat__lb_mask"),
                               rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(O.huber_entropy_lb(P, arch, detach_weights=True, add_sub_weight_ent=True).float(),
                               _t("rat__lb_flags"), rtol=2e-5, atol=2e-5)
    lb = O.huber_entropy_lb(P, arch, add_sub_weight_ent=True)
    (lb * torch.tensor([1.0, -0.5], dtype=torch.float64)).sum().backward()
    for k, p in raw.items():
        key = "rat__grad__" + k
        if key in GOLD.files:
            ref = _t(key).double()
            assert float((p.grad - ref).abs().max()) <= 2e-4 * float(ref.abs().max()) + 1e-8, k
    # product vs oracle on random models
    for seed, (F, D, R, S, I, C) in enumerate([(12, 2, 2, 3, 2, 1), (8, 3, 3, 2, 3, 2), (20, 2, 2, 4, 3, 3)]):
        torch.manual_seed(40 + seed)
        np.random.seed(40 + seed)
        cfg = rb.RatSpnConfig()
        cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = F, D, R, S, I, C
        cfg.dropout, cfg.leaf_base_class, cfg.leaf_base_kwargs, cfg.tanh_squash = 0.0, rb.RatNormal, {}, False
        m = rb.RatSpn(cfg)
        a = O.Arch(F=F, D=D, R=R, S=S, I=I, C=C)
        Pm = O.params_from_ratspn_state({k: v.detach().double() for k, v in m.state_dict().items()}, a,
                                        dtype=torch.float64)
        mask = torch.zeros(1, F, dtype=torch.bool)
        mask[0, 1::4] = True
        for li in range(2, a.max_layer_index + 1, 2):
            got, _ = m.huber_entropy_lb(layer_index=li, verbose=False, marginal_mask=mask)
            ref = O.huber_entropy_lb(Pm, a, layer_index=li, marginal_mask=mask)
            torch.testing.assert_close(got.double(), ref, rtol=1e-4, atol=1e-4)
