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
