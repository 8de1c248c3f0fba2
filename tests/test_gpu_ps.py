"""GPU: TMA-staged per-sample-weight (CSPN) Sum forward against the general exact kernel (bit-level agreement is not
required -- different summation order -- but both are exact fp32 log-sum-exp: <= 1e-5 relative)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _general(x, lw):
    """spn_sum_fwd (general kernel) on the same tensors, bypassing the dispatch in ops.SumFn."""
    from ratspn_b200 import ops
    from ratspn_b200._abi import call, ptr, stream_ptr
    xprod, sx, d_in, c, sw, ints, so, out_shape = ops._sum_geometry(x, lw, "xsum", "ref")
    out = torch.empty(out_shape, device=x.device, dtype=torch.float32)
    call("spn_sum_fwd", xprod, ptr(x), *sx, d_in, c, ptr(lw), *sw, *ints, ptr(out), *so, stream_ptr(x.device))
    return out


@pytest.mark.parametrize("n,w,d_in,c,oc,R", [(1, 48, 4, 16, 16, 8), (3, 20, 3, 8, 16, 4), (1, 7, 2, 16, 8, 8), (2, 5, 8, 12, 64, 16),
                                           (40, 6, 4, 16, 16, 8), (17, 5, 3, 8, 16, 4), (70, 3, 2, 32, 8, 8)])  # >= 16 rows per set: rows kernel
def test_per_sample_sum_forward_matches_general_kernel(n, w, d_in, c, oc, R):
    from ratspn_b200 import ops
    from ratspn_b200 import _abi
    assert _abi.lib().spn_sum_ps_supported(c, oc, R)
    g = torch.Generator().manual_seed(n * 100 + w)
    x = (torch.randn(n, w, d_in, c, R, generator=g) * 3.0 - 10.0)
    x[0, 0, 0, :, 0] = float("-inf")     # impossible child for one repetition -> -inf outputs there
    x[0, 1, 1, 2, :] = 200.0             # dominant channel
    d_out = max(1, (1 << (d_in - 1).bit_length()) // 2)
    raw = torch.randn(w, d_out, c * c, oc, R, generator=g) * 2.0
    raw[1, 0, 2 * c:3 * c] = -400.0      # ... whose products carry negligible weight: linear-domain underflow -> rescue
    lw = torch.log_softmax(raw, dim=2)
    x, lw = x.to(DEV), lw.to(DEV)
    ref = _general(x, lw)
    got = ops.sum_forward(x, lw, "xsum")
    if c * c * oc * R * 4 >= 16384:
        pass  # the dispatch took the streaming kernel
    assert got.shape == ref.shape
    both_inf = torch.isinf(ref) & torch.isinf(got) & (ref == got)
    assert not torch.isnan(got).any()
    err = torch.where(both_inf, torch.zeros_like(ref), (got - ref).abs())
    assert err.max().item() <= 1e-5 * ref[torch.isfinite(ref)].abs().max().item() + 1e-5
    assert torch.isinf(got[0, 0, 0, :, 0]).all()


def test_per_sample_sum_gradients_flow_through_general_backward():
    from ratspn_b200 import ops
    n, w, d_in, c, oc, R = 1, 6, 2, 16, 16, 8
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, w, d_in, c, R, generator=g).to(DEV).requires_grad_(True)
    raw = torch.randn(w, 1, c * c, oc, R, generator=g).to(DEV).requires_grad_(True)
    out = ops.sum_forward(x, torch.log_softmax(raw, dim=2), "xsum")
    out.sum().backward()
    # every output is a normalised mixture: d out / d raw sums to zero over the in-channels, d out / d x sums to oc per child
    assert raw.grad.sum(dim=2).abs().max().item() < 1e-4
    assert (x.grad.sum(dim=3) - oc).abs().max().item() < 1e-3


def _general_backward(x, lw, out, g):
    """spn_sum_bwd_x / spn_sum_bwd_w (general exact kernels) on the same tensors."""
    from ratspn_b200 import ops
    return ops._sum_backward(x, lw, out, g, "xsum", "ref", True, True)


@pytest.mark.parametrize("w,d_in,c,oc,R", [(48, 4, 16, 16, 8), (9, 3, 8, 16, 4), (5, 2, 32, 8, 8), (6, 7, 16, 32, 2), (3, 2, 16, 4, 32)])
def test_fused_per_sample_backward_matches_general_kernels(w, d_in, c, oc, R):
    """spn_sum_ps_bwd (one pass over the weights: dlw + both child gradients) against the exact general backward
    kernels; tolerance 2e-5 of the tensor's largest magnitude (ex2.approx vs expf, different summation order)."""
    from ratspn_b200 import _abi, ops
    assert _abi.lib().spn_sum_ps_bwd_supported(c, oc, R)
    gen = torch.Generator().manual_seed(w * 10 + d_in)
    x = (torch.randn(1, w, d_in, c, R, generator=gen) * 3.0 - 5.0)
    x[0, 0, 0, :, 0] = float("-inf")     # impossible left child in repetition 0 of conditional 0 -> -inf outputs, zero gradients
    d_out = max(1, (1 << (d_in - 1).bit_length()) // 2)
    lw = torch.log_softmax(torch.randn(w, d_out, c * c, oc, R, generator=gen) * 2.0, dim=2)
    gout = torch.randn(1, w, d_out, oc, R, generator=gen)
    x, lw, gout = x.to(DEV).requires_grad_(True), lw.to(DEV).requires_grad_(True), gout.to(DEV)
    out = ops.sum_forward(x, lw, "xsum")
    assert out.grad_fn is not None and out.grad_fn.ps_bwd[0], "the dispatch must have taken the streaming path"
    out.backward(gout)
    dx_ref, dlw_ref = _general_backward(x.detach(), lw.detach(), out.detach(), gout)
    for got, ref, what in ((x.grad, dx_ref, "dx"), (lw.grad, dlw_ref, "dlw")):
        assert not torch.isnan(got).any(), what
        scale = float(ref.abs().max())
        assert float((got - ref).abs().max()) <= 2e-5 * scale + 1e-7, f"{what}: {float((got - ref).abs().max()):.3e} vs scale {scale:.3e}"
    assert float(x.grad[0, 0, 0, :, 0].abs().max()) == 0.0 or torch.isfinite(x.grad[0, 0, 0, :, 0]).all()


@pytest.mark.parametrize("n", [1, 20])
def test_cspn_model_streaming_paths_match_general_kernels(n, monkeypatch):
    """Whole CSPN (gating net -> per-conditional weights) through the streaming kernels (n = 1: fused backward; n = 20:
    many-rows forward + general backward) against the same model forced onto the general exact kernels."""
    import numpy as np
    import ratspn_b200 as rb
    torch.manual_seed(3)
    np.random.seed(3)
    cfg = rb.CspnConfig()
    cfg.F_cond = (6,)
    cfg.F, cfg.R, cfg.D, cfg.I, cfg.S, cfg.C = 32, 4, 3, 8, 16, 1
    cfg.dropout = 0.0
    cfg.leaf_base_class = rb.RatNormal
    cfg.leaf_base_kwargs = {"min_sigma": 0.05, "max_sigma": 1.0}
    cfg.tanh_squash = False
    cfg.feat_layers, cfg.sum_param_layers, cfg.dist_param_layers = [8], [8], [8]
    cfg.cond_layers_inner_act = torch.nn.LeakyReLU
    m = rb.CSPN(cfg).to(DEV)
    w = 12
    cond = torch.randn(w, 6, device=DEV)
    x = torch.randn(n, w, 32, 1, 1, device=DEV)
    x[0, 1, ::5] = float("nan")
    res = []
    for general in (False, True):
        from ratspn_b200 import ops
        monkeypatch.setattr(ops, "FORCE_GENERAL_SUM", general)
        m.zero_grad(set_to_none=True)
        ll = m(x, condition=cond)
        (-ll.mean()).backward()
        with torch.no_grad():
            torch.manual_seed(11)
            H, _ = m.recursive_entropy_approx(condition=cond, sample_size=20)
        res.append((ll.detach().clone(), H.clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
    (ll_a, H_a, g_a), (ll_b, H_b, g_b) = res
    torch.testing.assert_close(ll_a, ll_b, rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(H_a, H_b, rtol=1e-4, atol=1e-3)
    assert set(g_a) == set(g_b)
    for k in g_b:
        scale = float(g_b[k].abs().max()) + 1e-12
        assert float((g_a[k] - g_b[k]).abs().max()) <= 2e-4 * scale + 1e-7, k


@pytest.mark.parametrize("n", [1, 3])
def test_cspn_fused_weight_normalisation_matches_reference_semantics(n):
    """CspnConfig.fuse_weight_norm: the raw head outputs go into spn_sum_ps_fwd_raw / _bwd_raw (log_softmax fused).
    Same LL, same gradients w.r.t. the gating network, same MPE sample and entropy as the unfused model."""
    import numpy as np
    import ratspn_b200 as rb

    def build(fused):
        torch.manual_seed(5)
        np.random.seed(5)
        cfg = rb.CspnConfig()
        cfg.F_cond = (6,)
        cfg.F, cfg.R, cfg.D, cfg.I, cfg.S, cfg.C = 32, 4, 3, 8, 16, 1
        cfg.dropout = 0.0
        cfg.leaf_base_class = rb.RatNormal
        cfg.leaf_base_kwargs = {"min_sigma": 0.05, "max_sigma": 1.0}
        cfg.tanh_squash = False
        cfg.feat_layers, cfg.sum_param_layers, cfg.dist_param_layers = [8], [8], [8]
        cfg.cond_layers_inner_act = torch.nn.LeakyReLU
        cfg.fuse_weight_norm = fused
        return rb.CSPN(cfg).to(DEV)

    ref, fus = build(False), build(True)
    fus.load_state_dict(ref.state_dict())
    w = 10
    cond = torch.randn(w, 6, device=DEV) * 3.0   # large head outputs: the running-maximum rescaling is exercised
    x = torch.randn(n, w, 32, 1, 1, device=DEV)
    out = []
    for m in (ref, fus):
        m.zero_grad(set_to_none=True)
        ll = m(x, condition=cond)
        (-ll.mean()).backward()
        with torch.no_grad():
            mpe = m.sample(mode="index", condition=cond, is_mpe=True).sample
            torch.manual_seed(9)
            H, _ = m.recursive_entropy_approx(condition=cond, sample_size=4)
        out.append((ll.detach(), mpe, H, {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
    if n == 1:
        assert fus._inner_layers[1]._raw_weights and not ref._inner_layers[1]._raw_weights
        # weight_param holds the raw head outputs, .weights normalises on read
        assert float(fus._inner_layers[1].weights.exp().sum(2).sub(1).abs().max()) < 1e-4
    (ll_a, mpe_a, H_a, g_a), (ll_b, mpe_b, H_b, g_b) = out
    torch.testing.assert_close(ll_b, ll_a, rtol=1e-5, atol=2e-4)
    torch.testing.assert_close(mpe_b, mpe_a)
    torch.testing.assert_close(H_b, H_a, rtol=1e-4, atol=1e-3)
    assert set(g_a) == set(g_b)
    for k in g_a:
        scale = float(g_a[k].abs().max()) + 1e-12
        assert float((g_a[k] - g_b[k]).abs().max()) <= 3e-4 * scale + 1e-7, k
