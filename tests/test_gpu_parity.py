"""GPU parity tests: the CUDA path (through the Python mirror of the reference API -> ctypes -> libspn_b200.so) against
(a) the golden vectors recorded from the unmodified reference and (b) the CPU oracle on seeded inputs.

Tolerances (stated by BASELINE.json's north star: |d log-lik| <= 1e-4 relative in fp32; indices bit-exact):
  log-likelihoods        rtol 1e-4 (we hold 2e-5) + atol 1e-4
  parameter gradients    rtol 2e-3 of the largest gradient magnitude of the tensor
  sampled indices        exact;   sampled values rtol/atol 1e-5
"""
import numpy as np
import pytest
import torch

from oracle import spn_oracle as O
from tests.helpers import Case, assert_close, build_cspn, build_ratspn, oracle_params

pytestmark = pytest.mark.gpu
DEV = "cuda"
LL_RT, LL_AT = 2e-5, 1e-4


def grad_close(got, ref, what):
    scale = float(ref.abs().max()) + 1e-12
    assert_close(got, ref, 0.0, 2e-3 * scale, what)


def model_grads(model):
    return {k: p.grad for k, p in model.named_parameters() if p.requires_grad and p.grad is not None}


def test_library_is_loaded_and_reports_sm100():
    import ratspn_b200 as rb
    assert "sm_100a" in rb._abi.version()
    assert torch.cuda.get_device_capability(0)[0] == 10


@pytest.mark.parametrize("name,layers", [("ratspn_small", [0, 1, 2, 3, 4]), ("ratspn_pad_tanh", [0, 1, 2, 5, 8])])
def test_forward_layers_vs_golden(name, layers):
    c = Case(name)
    m = build_ratspn(c, DEV)
    x = c.t("x").to(DEV)
    for li in layers:
        out = m(x, layer_index=li)
        assert_close(out, c.t(f"ll_layer{li}"), LL_RT, LL_AT, f"{name} layer {li}")


def test_forward_nan_and_channel_rep_inputs():
    c = Case("ratspn_small")
    m = build_ratspn(c, DEV)
    out = m(c.t("x_nan").to(DEV))
    assert_close(out, c.t("ll_nan"), LL_RT, LL_AT, "NaN input")
    assert out[2].abs().max().item() < 1e-5  # fully marginalised row
    out = m(c.t("x_ir").to(DEV), layer_index=2, x_needs_permutation=False)
    assert_close(out, c.t("ll_ir_layer2"), LL_RT, LL_AT, "x with I and R dims")


@pytest.mark.parametrize("name,xkey", [("ratspn_small", "x_nan"), ("ratspn_pad_tanh", "x"), ("ratspn_cfg1", "x")])
def test_backward_vs_golden(name, xkey):
    c = Case(name)
    m = build_ratspn(c, DEV)
    out = m(c.t(xkey).to(DEV))
    if c.has("ll"):
        assert_close(out, c.t("ll"), LL_RT, LL_AT, f"{name} forward")
    (-out.mean()).backward()
    ref = c.grads()
    got = model_grads(m)
    assert set(got) == set(ref)
    for k in ref:
        grad_close(got[k], ref[k], f"{name} grad {k}")


def test_backward_wrt_input_matches_oracle_autograd():
    c = Case("ratspn_pad_tanh")
    m = build_ratspn(c, DEV)
    P, _ = oracle_params(c)
    x = torch.atanh(torch.rand(5, 1, c.F, 1, c.R, generator=torch.Generator().manual_seed(5)) * 1.8 - 0.9)
    xo = x.clone().requires_grad_(True)
    O.forward(P, c.arch, xo, x_needs_permutation=False).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    m(xg, x_needs_permutation=False).sum().backward()
    grad_close(xg.grad, xo.grad, "d LL / d x")
    # broadcast input (one value per feature, permuted inside the kernel): gradient must be reduced over (i, r)
    x1 = torch.atanh(torch.rand(3, 1, c.F, 1, 1, generator=torch.Generator().manual_seed(6)) * 1.8 - 0.9)
    xo = x1.clone().requires_grad_(True)
    O.forward(P, c.arch, xo).sum().backward()
    xg = x1.to(DEV).requires_grad_(True)
    m(xg).sum().backward()
    grad_close(xg.grad, xo.grad, "d LL / d x (broadcast + permutation)")


def test_standalone_layers_match_oracle():
    import ratspn_b200 as rb
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 2, 5, 3, 2, generator=g)  # odd scope count -> zero padded to 6... pow2 = 8
    xp = rb.CrossProduct(in_features=5, in_channels=3, num_repetitions=2).to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    out = xp(xg)
    xo = x.clone().requires_grad_(True)
    ref = O.cross_product(xo)
    assert_close(out, ref, 1e-6, 1e-6, "CrossProduct.forward")
    wgt = torch.randn(out.shape, generator=g)
    (out * wgt.to(DEV)).sum().backward()
    (ref * wgt).sum().backward()
    grad_close(xg.grad, xo.grad, "CrossProduct backward")
    s = rb.Sum(in_channels=9, in_features=4, out_channels=3, ratspn=True, num_repetitions=2).to(DEV)
    h = torch.randn(4, 1, 4, 9, 2, generator=g) * 3
    h[0, 0, 1, :, 0] = float("-inf")  # a fully impossible scope must give -inf like th.logsumexp
    lw = torch.log_softmax(s.weight_param.detach().cpu(), dim=2)
    assert_close(s(h.to(DEV)), O.sum_layer(h, lw), 1e-5, 1e-5, "Sum.forward")


def test_sum_underflow_rescue_is_exact():
    """Peaked weights on channels far below the input maximum: the max-shifted fast path underflows and the kernel must
    fall back to the exact log-sum-exp (what th.logsumexp computes in the reference, layers.py:123)."""
    import ratspn_b200 as rb
    s = rb.Sum(in_channels=4, in_features=1, out_channels=2, ratspn=False, num_repetitions=1).to(DEV)
    lw = torch.tensor([[-200.0, -1e-9], [-200.0, -90.0], [-1e-9, -90.0], [-200.0, -90.0]]).view(1, 1, 4, 2, 1)
    lw = torch.log_softmax(lw, dim=2)
    del s.weight_param
    s.weight_param = lw.to(DEV)
    h = torch.tensor([0.0, -50.0, -300.0, -120.0]).view(1, 1, 1, 4, 1)
    assert_close(s(h.to(DEV)), O.sum_layer(h, lw), 1e-6, 1e-5, "rescue path")


@pytest.mark.parametrize("tag,kw", [("samp", dict(n=5)), ("mpe", dict(n=2, is_mpe=True)),
                                    ("cls", dict(n=3, class_index=torch.tensor([1])))])
def test_index_sampling_bit_exact_vs_golden(tag, kw):
    c = Case("ratspn_small")
    m = build_ratspn(c, DEV)
    kw = dict(kw)
    if "class_index" in kw:
        kw["class_index"] = kw["class_index"].to(DEV)
    ctx = m.sample(mode="index", noise=c.noise(tag), **kw)
    assert torch.equal(ctx.repetition_indices.cpu(), c.t(f"{tag}__rep"))
    assert torch.equal(ctx.parent_indices.cpu(), c.t(f"{tag}__parent"))
    assert_close(ctx.sample, c.t(f"{tag}__sample"), 1e-5, 1e-5, tag)


@pytest.mark.parametrize("tag,kw", [("ev", dict(n=3, evidence=True)), ("evmpe", dict(n=1, evidence=True, is_mpe=True)),
                                    ("l3", dict(n=2, layer_index=3)), ("l2", dict(n=2, layer_index=2)),
                                    ("l7raw", dict(n=2, layer_index=7, do_sample_postprocessing=False)),
                                    ("l0raw", dict(n=3, layer_index=0, do_sample_postprocessing=False))])
def test_sampling_evidence_and_inner_layers_vs_golden(tag, kw):
    c = Case("ratspn_pad_tanh")
    m = build_ratspn(c, DEV)
    kw = dict(kw)
    if kw.pop("evidence", False):
        kw["evidence"] = c.t("evidence").to(DEV)
    ctx = m.sample(mode="index", noise=c.noise(tag), **kw)
    assert_close(ctx.sample, c.t(f"{tag}__sample"), 1e-5, 1e-5, tag)
    if c.has(f"{tag}__rep"):
        assert torch.equal(ctx.repetition_indices.cpu(), c.t(f"{tag}__rep"))


def test_cfg1_sampling_vs_golden_and_postprocessing_paths_agree():
    c = Case("ratspn_cfg1")
    m = build_ratspn(c, DEV)
    ctx = m.sample(mode="index", n=3, noise=c.noise("samp"))
    assert torch.equal(ctx.repetition_indices.cpu(), c.t("samp__rep"))
    assert_close(ctx.sample, c.t("samp__sample"), 1e-5, 1e-5)
    raw = m.sample(mode="index", n=3, noise=c.noise("samp"), do_sample_postprocessing=False)
    post = m.sample_postprocessing(raw)  # torch-op post-processing must equal the fused kernel epilogue
    assert torch.equal(post.sample, ctx.sample)


def test_onehot_equals_index_under_mpe():
    """Reference test invariant (tests/test_RatSpn.py:21-24)."""
    for name in ("ratspn_small", "ratspn_pad_tanh"):
        m = build_ratspn(Case(name), DEV)
        a = m.sample(mode="onehot", is_mpe=True).sample
        b = m.sample(mode="index", is_mpe=True).sample
        assert torch.equal(a, b)


def test_philox_sampling_statistics():
    """In-kernel RNG: root-channel frequencies follow exp(log w) and leaf draws have the selected (mean, std)."""
    c = Case("ratspn_small")
    m = build_ratspn(c, DEV)
    torch.manual_seed(0)
    n = 200000
    ctx = m.sample(mode="index", n=n, class_index=torch.tensor([0], device=DEV), do_sample_postprocessing=False)
    rep = ctx.repetition_indices.view(-1)
    lwr = torch.log_softmax(m.root.weight_param.detach(), dim=2)[0, 0, :, 0, 0]  # [R*ic]
    p_rep = lwr.exp().view(c.R, -1).sum(1)
    freq = torch.bincount(rep, minlength=c.R).float() / n
    assert (freq - p_rep).abs().max().item() < 5e-3
    # z-scores of the raw samples given their selected leaf must be standard normal
    P, _ = oracle_params(c)
    ch = ctx.parent_indices.view(n, c.F)
    mu = P["mu"][0].to(DEV)[torch.arange(c.F, device=DEV).view(1, -1), ch, rep.view(-1, 1)]
    sg = P["sigma"][0].to(DEV)[torch.arange(c.F, device=DEV).view(1, -1), ch, rep.view(-1, 1)]
    z = (ctx.sample.view(n, c.F) - mu) / sg
    assert abs(z.mean().item()) < 5e-3 and abs(z.std().item() - 1.0) < 5e-3
    ctx2 = m.sample(mode="index", n=16, class_index=torch.tensor([0], device=DEV))
    torch.manual_seed(0)
    a = m.sample(mode="index", n=16).sample
    torch.manual_seed(0)
    b = m.sample(mode="index", n=16).sample
    assert torch.equal(a, b) and not torch.equal(a[:, :8], a[:, 8:])  # reproducible under the torch seed


@pytest.mark.parametrize("tag,mask", [("ent", False), ("entmask", True)])
def test_recursive_entropy_no_grad_vs_golden(tag, mask):
    c = Case("ratspn_pad_tanh")
    m = build_ratspn(c, DEV)
    m._test_noise = c.noise(tag)
    with torch.no_grad():
        H, _ = m.recursive_entropy_approx(sample_size=4, marginal_mask=c.t("marginal_mask").to(DEV) if mask else None)
    assert len(m._test_noise) == 0
    assert_close(H, c.t(f"{tag}__H"), 1e-4, 1e-4, tag)


@pytest.mark.parametrize("verbose", [True, False])  # False: fused spn_entropy_combine kernels; True: logged composite
@pytest.mark.parametrize("tag,aux", [("entgrad", False), ("entgradaux", True)])
def test_recursive_entropy_gradients_vs_golden(tag, aux, verbose):
    c = Case("ratspn_pad_tanh")
    m = build_ratspn(c, DEV)
    m._test_noise = c.noise(tag)
    H, log = m.recursive_entropy_approx(sample_size=4, aux_with_grad=aux, marginal_mask=c.t("marginal_mask").to(DEV),
                                        verbose=verbose)
    assert_close(H, c.t(f"{tag}__H"), 1e-4, 1e-4, tag)
    assert any(k.startswith("lay8/rep0/node_entropy") for k in log) == verbose
    (-H.sum()).backward()
    ref = c.grads(f"{tag}__grad__")
    got = model_grads(m)
    for k in ref:
        grad_close(got[k], ref[k], f"{tag} grad {k}")


def test_onehot_sampling_with_evidence_gradients_vs_golden():
    c = Case("ratspn_pad_tanh")
    m = build_ratspn(c, DEV)
    ctx = m.sample(mode="onehot", n=2, evidence=c.t("evidence").to(DEV), noise=c.noise("ohev"))
    ref = c.t("ohev__sample")
    assert_close(ctx.sample, ref, 1e-5, 1e-5)
    cw = torch.linspace(-1.0, 1.0, ref.numel()).view_as(ref).to(DEV)
    (ctx.sample * cw).sum().backward()
    g = c.grads("ohev__grad__")
    got = model_grads(m)
    for k in g:
        grad_close(got.get(k, torch.zeros_like(g[k])), g[k], f"ohev grad {k}")


def test_cspn_forward_gradients_and_sac_pattern_vs_golden():
    c = Case("cspn_small")
    m = build_cspn(c, DEV)
    cond = c.t("cond").to(DEV)
    out = m(c.t("x").to(DEV), condition=cond)
    assert_close(out, c.t("ll"), LL_RT, LL_AT, "CSPN forward")
    assert_close(m.root.weight_param, c.t("setparams__lw_root"), 1e-5, 1e-6, "set_params root")
    assert_close(m._leaf.base_leaf.std_param, c.t("setparams__sigma"), 1e-5, 1e-6, "set_params sigma")
    (-out.mean()).backward()
    ref = c.grads("ll__grad__")
    got = model_grads(m)
    for k in ref:
        grad_close(got[k], ref[k], f"cspn ll grad {k}")
    # index sampling with evidence (rollout path, sb3.py:196-203)
    ev = c.t("evidence").to(DEV)
    with torch.no_grad():
        ctx = m.sample(mode="index", condition=cond, evidence=ev, noise=c.noise("ixev"))
    assert_close(ctx.sample, c.t("ixev__sample"), 1e-5, 1e-5, "index + evidence")
    # policy update pattern: onehot sample + recursive entropy with marginal mask, backward (sb3.py:146-184, :336-344)
    m.zero_grad()
    m._test_noise = c.noise("sac")
    ctx = m.sample(mode="onehot", condition=cond, evidence=ev)
    H, _ = m.recursive_entropy_approx(condition=cond, sample_size=3, marginal_mask=c.t("failed").to(DEV))
    assert len(m._test_noise) == 0
    assert_close(ctx.sample, c.t("sac__sample"), 1e-5, 1e-5, "SAC sample")
    assert_close(H, c.t("sac__H"), 1e-4, 1e-4, "SAC entropy")
    (-(0.1 * H + ctx.sample.squeeze(0).squeeze(0).squeeze(-1).sum(-1)).mean()).backward()
    ref = c.grads("sac__grad__")
    got = model_grads(m)
    for k in ref:
        grad_close(got[k], ref[k], f"sac grad {k}")


def test_seeded_random_architectures_vs_oracle():
    """Forward + backward + index sampling on a few seeded architectures (incl. D = 1, odd scope counts, C > 1)."""
    import ratspn_b200 as rb
    for seed, (F, D, R, S, I, C) in enumerate([(6, 1, 2, 3, 4, 3), (21, 3, 4, 5, 2, 2), (50, 4, 2, 3, 3, 1),
                                               (12, 2, 7, 2, 6, 1)]):
        torch.manual_seed(100 + seed)
        np.random.seed(100 + seed)
        cfg = rb.RatSpnConfig()
        cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = F, D, R, S, I, C
        cfg.dropout, cfg.leaf_base_class, cfg.leaf_base_kwargs, cfg.tanh_squash = 0.0, rb.RatNormal, {}, False
        m = rb.RatSpn(cfg)
        arch = O.Arch(F=F, D=D, R=R, S=S, I=I, C=C)
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        raw = {}
        for k in list(sd):
            if k.split(".")[-1] in ("weight_param", "mean_param", "std_param") and not k.startswith("_sampling_root"):
                sd[k] = sd[k].requires_grad_(True)
                raw[k] = sd[k]
        P = O.params_from_ratspn_state(sd, arch)
        x = torch.randn(9, 1, F, 1, 1)
        x[3, 0, ::3] = float("nan")
        ref = O.forward(P, arch, x)
        (-ref.mean()).backward()
        m = m.to(DEV)
        out = m(x.to(DEV))
        assert_close(out, ref, LL_RT, LL_AT, f"arch {seed} forward")
        (-out.mean()).backward()
        got = model_grads(m)
        for k, t in raw.items():
            grad_close(got[k], t.grad, f"arch {seed} grad {k}")
        ns = O.NoiseSource(generator=torch.Generator().manual_seed(seed))
        with torch.no_grad():
            Pn = O.params_from_ratspn_state({k: v.detach() for k, v in sd.items()}, arch)
            o = O.sample_index(Pn, arch, ns, n=(4,))
        ctx = m.sample(mode="index", n=4, noise=list(ns.record))
        assert torch.equal(ctx.repetition_indices.cpu(), o.repetition_indices)
        assert torch.equal(ctx.parent_indices.cpu(), o.parent_indices)
        assert_close(ctx.sample, o.sample, 1e-5, 1e-5, f"arch {seed} sample")


def test_baseline_config2_width_properties_and_oracle_slice():
    """BASELINE config 2 width (F=784, D=5, R=40, S=I=40): size-independent properties at a large batch plus a 3-row
    comparison with the micro-batched oracle (the reference cannot hold this width beyond a handful of rows)."""
    import ratspn_b200 as rb
    torch.manual_seed(0)
    np.random.seed(0)
    cfg = rb.RatSpnConfig()
    cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = 784, 5, 40, 40, 40, 1
    cfg.dropout, cfg.leaf_base_class, cfg.leaf_base_kwargs, cfg.tanh_squash = 0.0, rb.RatNormal, {}, False
    m = rb.RatSpn(cfg)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(DEV)
    B = 768
    x = torch.rand(B, 1, 784, 1, 1, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        full = m(x.to(DEV))
        halves = torch.cat([m(x[:300].to(DEV)), m(x[300:].to(DEV))])  # both parts take the same (tensor-core) path
        assert torch.equal(full, halves)  # rows are independent: batching must not change any value
        allnan = m(torch.full((2, 1, 784, 1, 1), float("nan"), device=DEV))
        assert allnan.abs().max().item() < 1e-4  # full marginalisation: only normalised weights remain
        swapped = m(x.to(DEV).transpose(0, 1))  # batch in the w dim (mnist_gen_train.py:670-675) == batch in n dim
        assert torch.equal(swapped.view(-1), full.view(-1))
    arch = O.Arch(F=784, D=5, R=40, S=40, I=40, C=1)
    P = O.params_from_ratspn_state(sd, arch)
    with torch.no_grad():
        ref = O.forward(P, arch, x[:3], microbatch=1)
    assert_close(full[:3], ref, LL_RT, LL_AT, "config-2 width, 3 rows vs oracle")


def test_cspn_conditioned_block_shares_one_gating_pass_vs_golden():
    """``with cspn.conditioned(obs)``: sample + entropy on one evaluation of the gating network; same values and
    gradients as the reference's policy-update pattern (which evaluates it twice)."""
    c = Case("cspn_small")
    m = build_cspn(c, DEV)
    cond = c.t("cond").to(DEV)
    ev = c.t("evidence").to(DEV)
    m._test_noise = c.noise("sac")
    calls = []
    orig = m.set_params
    m.set_params = lambda x: (calls.append(1), orig(x))[1]
    with m.conditioned(cond):
        ctx = m.sample(mode="onehot", condition=cond, evidence=ev)
        H, _ = m.recursive_entropy_approx(condition=cond, sample_size=3, marginal_mask=c.t("failed").to(DEV))
    assert len(calls) == 1 and len(m._test_noise) == 0
    assert_close(ctx.sample, c.t("sac__sample"), 1e-5, 1e-5, "SAC sample")
    assert_close(H, c.t("sac__H"), 1e-4, 1e-4, "SAC entropy")
    (-(0.1 * H + ctx.sample.squeeze(0).squeeze(0).squeeze(-1).sum(-1)).mean()).backward()
    ref = c.grads("sac__grad__")
    got = model_grads(m)
    for k in ref:
        grad_close(got[k], ref[k], f"sac grad {k}")
    m._test_noise = None
    m.sample(mode="index", condition=cond)   # outside the block the gating network runs again
    assert len(calls) == 2
