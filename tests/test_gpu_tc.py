"""GPU: the tcgen05 tensor-core Sum path against the exact fp32 kernel (same inputs, internal layout)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def exact_reference(x_drbc, lw):
    """Exact fp32 result through the general kernel: convert [d,R,B,c] -> reference layout [B,1,d,c,R]."""
    from ratspn_b200 import ops
    x_ref = x_drbc.permute(2, 0, 3, 1).unsqueeze(1).contiguous()
    out = ops.sum_forward(x_ref, lw, "xsum")  # [B,1,d_out,oc,R]
    return out.squeeze(1).permute(1, 3, 0, 2).contiguous()  # [d_out,R,B,oc]


@pytest.mark.parametrize("c,d_in,R,B", [(16, 4, 3, 300), (16, 3, 2, 128), (32, 2, 2, 257), (40, 4, 3, 300), (40, 7, 2, 1024)])
def test_tc_forward_matches_exact_kernel(c, d_in, R, B):
    from ratspn_b200 import ops
    assert ops.tc_supported(c, c, R)
    g = torch.Generator().manual_seed(c * 1000 + B)
    x = (torch.randn(d_in, R, B, c, generator=g) * 4.0 - 30.0).to(DEV)
    d_out = max(1, (1 << (d_in - 1).bit_length()) // 2)
    lw = torch.log_softmax(torch.randn(1, d_out, c * c, c, R, generator=g), dim=2).to(DEV)
    wimg = ops.tc_prepare(lw)
    out = ops.tc_sum_forward(x, wimg, lw)
    torch.cuda.synchronize()
    assert ops.tc_aborts() == 0, "an mbarrier wait timed out inside the tensor-core kernel"
    ref = exact_reference(x, lw)
    err = (out - ref).abs().max().item()
    # bf16 operands, fp32 accumulation: absolute error of a log of a K-term sum; bound 2e-2 nats, typically ~1e-3
    assert err < 2e-2, f"max abs err {err}"
    assert (out - ref).abs().mean().item() < 2e-3


def test_tc_forward_keeps_logsumexp_semantics_on_underflow_and_minus_inf():
    from ratspn_b200 import ops
    c, d_in, R, B = 16, 2, 1, 128
    g = torch.Generator().manual_seed(1)
    x = torch.randn(d_in, R, B, c, generator=g)
    x[0, 0, 0, :] = float("-inf")           # an impossible left scope -> -inf output, not NaN
    x[0, 0, 1, 0] = 400.0                   # one dominant channel: every other term underflows in linear space
    raw = torch.randn(1, 1, c * c, c, R, generator=g)
    raw[0, 0, :c, :, 0] = -300.0            # ... and the dominant channel's products have negligible weight
    lw = torch.log_softmax(raw, dim=2)
    x, lw = x.to(DEV), lw.to(DEV)
    out = ops.tc_sum_forward(x, ops.tc_prepare(lw), lw)
    ref = exact_reference(x, lw)
    assert torch.isinf(out[0, 0, 0]).all() and (out[0, 0, 0] < 0).all()
    assert not torch.isnan(out).any()
    assert (out[:, :, 1:] - ref[:, :, 1:]).abs().max().item() < 2e-2
    assert ops.tc_aborts() == 0


def test_model_forward_backward_on_tensor_core_path_vs_oracle():
    """Whole-model check of the tensor-core path (internal layout, tcgen05 Sum layers, exact root) against the CPU
    oracle: |d LL| <= 1e-4 relative (BASELINE.json tolerance), parameter gradients within 2% of the tensor's max."""
    import numpy as np
    import ratspn_b200 as rb
    from oracle import spn_oracle as O
    torch.manual_seed(5)
    np.random.seed(5)
    F, D, R, S, I, C = 64, 3, 3, 16, 16, 2
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
    B = 300
    x = torch.randn(B, 1, F, 1, 1)
    x[7, 0, ::5] = float("nan")
    ref = O.forward(O.params_from_ratspn_state(sd, arch), arch, x)
    (-ref.mean()).backward()
    m = m.to(DEV)
    assert m._tensor_core_path_ok(x.to(DEV), False)
    out = m(x.to(DEV))
    rel = ((out.cpu() - ref).abs() / ref.abs().clamp_min(1.0)).max().item()
    assert rel < 1e-4, f"relative LL error {rel}"
    (-out.mean()).backward()
    for k, p in m.named_parameters():
        if p.requires_grad:
            g, r = p.grad.cpu(), raw[k].grad
            assert (g - r).abs().max().item() <= 2e-2 * r.abs().max().item() + 1e-9, k
    # partial forwards through the internal layout agree with the exact path
    m.use_tensor_cores = False
    exact2 = m(x.to(DEV), layer_index=2)
    exact3 = m(x.to(DEV), layer_index=3)
    m.use_tensor_cores = True
    assert (m(x.to(DEV), layer_index=2) - exact2).abs().max().item() < 2e-2
    assert (m(x.to(DEV), layer_index=3) - exact3).abs().max().item() < 4e-2
    assert torch.equal(m(x.to(DEV), layer_index=0), m._leaf(x.to(DEV), perm32=m._perm32()[0]))
    from ratspn_b200 import ops
    assert ops.tc_aborts() == 0


@pytest.mark.parametrize("c,d_in,R,B", [(16, 4, 3, 300), (16, 3, 2, 128), (32, 2, 2, 257), (40, 4, 3, 300), (40, 7, 2, 1024)])
def test_tc_backward_matches_exact_kernels(c, d_in, R, B):
    """dX and dW tensor-core kernels against the exact fp32 backward kernels on the same inputs."""
    from ratspn_b200 import ops
    g = torch.Generator().manual_seed(c * 77 + B)
    x = (torch.randn(d_in, R, B, c, generator=g) * 4.0 - 30.0).to(DEV)
    d_out = max(1, (1 << (d_in - 1).bit_length()) // 2)
    lw = torch.log_softmax(torch.randn(1, d_out, c * c, c, R, generator=g), dim=2).to(DEV)
    go = torch.randn(d_out, R, B, c, generator=g).to(DEV)
    # exact: general kernels through autograd on the internal layout
    xe = x.clone().requires_grad_(True)
    lwe = lw.clone().requires_grad_(True)
    oe = ops.sum_forward(xe, lwe, "xsum", "drbc")
    (oe * go).sum().backward()
    # tensor cores
    xt = x.clone().requires_grad_(True)
    lwt = lw.clone().requires_grad_(True)
    ot = ops.sum_forward_tc(xt, lwt)
    (ot * go).sum().backward()
    torch.cuda.synchronize()
    assert ops.tc_aborts() == 0
    assert (ot - oe).abs().max().item() < 2e-2
    sx = xe.grad.abs().max().item()
    sw = lwe.grad.abs().max().item()
    ex = (xt.grad - xe.grad).abs().max().item()
    ew = (lwt.grad - lwe.grad).abs().max().item()
    assert ex <= 2e-2 * sx, f"dx: max err {ex} vs scale {sx}"
    assert ew <= 2e-2 * sw, f"dlw: max err {ew} vs scale {sw}"
    assert (xt.grad - xe.grad).abs().mean().item() <= 2e-3 * sx
    assert (lwt.grad - lwe.grad).abs().mean().item() <= 2e-3 * sw


@pytest.mark.parametrize("c,d_in,R,B", [(16, 4, 3, 300), (40, 4, 3, 300), (40, 3, 40, 640)])
def test_tc_fused_weight_normalisation_forward_and_backward(c, d_in, R, B):
    """raw = True: the kernels take the unnormalised parameter and fuse F.log_softmax(dim=2) (layers.py:85) and its
    backward; compared with autograd through torch's log_softmax + the exact fp32 kernels."""
    from ratspn_b200 import ops
    g = torch.Generator().manual_seed(c * 13 + B)
    x = (torch.randn(d_in, R, B, c, generator=g) * 4.0 - 30.0).to(DEV)
    d_out = max(1, (1 << (d_in - 1).bit_length()) // 2)
    raw = (torch.randn(1, d_out, c * c, c, R, generator=g) * 1.5 + 0.7).to(DEV)
    go = torch.randn(d_out, R, B, c, generator=g).to(DEV)
    pe = raw.clone().requires_grad_(True)
    oe = ops.sum_forward(x, torch.log_softmax(pe, dim=2), "xsum", "drbc")
    (oe * go).sum().backward()
    pt = raw.clone().requires_grad_(True)
    cache = {}
    ot = ops.sum_forward_tc(x, pt, True, cache)
    (ot * go).sum().backward()
    torch.cuda.synchronize()
    assert ops.tc_aborts() == 0
    logz = torch.logsumexp(raw, dim=2).squeeze(0)  # [d, oc, R]
    assert (cache["logz"] - logz).abs().max().item() < 1e-5
    assert (ot - oe).abs().max().item() < 2e-2
    sw = pe.grad.abs().max().item()
    assert (pt.grad - pe.grad).abs().max().item() <= 2e-2 * sw
    assert (pt.grad - pe.grad).abs().mean().item() <= 2e-3 * sw
    # the image is re-used while the parameter is unchanged and rebuilt after an in-place update
    img = cache["wimg"]
    ops.sum_forward_tc(x, pt, True, cache)
    assert cache["wimg"] is img
    with torch.no_grad():
        pt.add_(0.5 * torch.randn(pt.shape, device=DEV))
    o2 = ops.sum_forward_tc(x, pt, True, cache)
    assert cache["wimg"] is not img
    ref2 = ops.sum_forward(x, torch.log_softmax(pt.detach(), dim=2), "xsum", "drbc")
    assert (o2 - ref2).abs().max().item() < 2e-2
