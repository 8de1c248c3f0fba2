"""GPU: dedicated shared-weight root kernels (internal layout) against the general exact kernels."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("c,R,C,B,d_in", [(8, 3, 1, 100, 2), (40, 5, 1, 300, 2), (16, 4, 3, 257, 2), (64, 2, 2, 130, 2),
                                          (8, 2, 2, 64, 1)])
def test_root_shared_forward_backward_match_general_kernels(c, R, C, B, d_in):
    from ratspn_b200 import ops
    assert ops.root_shared_supported(c, C, R)
    g = torch.Generator().manual_seed(c * 31 + B)
    x = (torch.randn(d_in, R, B, c, generator=g) * 4.0 - 30.0)
    x[0, :, 3, :] = float("-inf")            # an impossible row: out = -inf, zero gradients, no NaN
    x[d_in - 1, 1, 5, 0] = 300.0             # a dominant channel
    x = x.to(DEV)
    raw = torch.randn(1, 1, R * c * c, C, 1, generator=g).to(DEV)
    go = torch.randn(B, C, generator=g).to(DEV)
    go[3] = 0.0

    xe = x.clone().requires_grad_(True)
    pe = raw.clone().requires_grad_(True)
    oe = ops.sum_forward(xe, torch.log_softmax(pe, dim=2), "xroot", "drbc").reshape(B, C)
    keep = torch.ones(B, dtype=torch.bool, device=DEV)
    keep[3] = False
    (oe[keep] * go[keep]).sum().backward()

    xt = x.clone().requires_grad_(True)
    pt = raw.clone().requires_grad_(True)
    ot = ops.root_forward_shared(xt, torch.log_softmax(pt, dim=2))
    assert torch.isinf(ot[3]).all() and (ot[3] < 0).all() and not torch.isnan(ot).any()
    (ot[keep] * go[keep]).sum().backward()
    assert (ot[keep] - oe[keep]).abs().max().item() <= 1e-4 * oe[keep].abs().max().item() + 1e-4
    sx = xe.grad.abs().max().item()
    sw = pe.grad.abs().max().item()
    assert not torch.isnan(xt.grad).any() and not torch.isnan(pt.grad).any()
    assert (xt.grad - xe.grad).abs().max().item() <= 1e-4 * sx + 1e-7
    assert (pt.grad - pe.grad).abs().max().item() <= 1e-4 * sw + 1e-7
