"""GPU: the flat single-launch Adam (spn_adam_step via parallel.FlatAdam) against torch.optim.Adam on the same
parameters and gradients (semantics of mnist_experiments/mnist_gen_train.py:599); rtol 1e-5 after 5 steps."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_flat_adam_matches_torch_adam():
    from ratspn_b200.parallel import FlatAdam, FlatGradAllReducer
    g = torch.Generator().manual_seed(0)
    shapes = [(3, 5, 7), (129,), (64, 64), (1, 2, 9, 4, 3)]
    ours = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    red = FlatGradAllReducer(ours)
    opt = FlatAdam(red, lr=3e-3, betas=(0.9, 0.99), eps=1e-8)
    topt = torch.optim.Adam(ref, lr=3e-3, betas=(0.9, 0.99), eps=1e-8)
    for p in ours:
        assert p.data_ptr() % 256 == 0 and p.grad.data_ptr() % 256 == 0
    for step in range(5):
        grads = [torch.randn(s, generator=g).to(DEV) * (10.0 ** (step - 2)) for s in shapes]
        red.zero_()
        for p, r, gr in zip(ours, ref, grads):
            p.grad.add_(gr)
            r.grad = gr.clone()
        opt.step()
        topt.step()
    for p, r in zip(ours, ref):
        torch.testing.assert_close(p.detach(), r.detach(), rtol=1e-5, atol=1e-7)


def test_flat_adam_trains_a_ratspn_and_invalidates_operand_images():
    import numpy as np
    import ratspn_b200 as rb
    from ratspn_b200.parallel import FlatAdam, FlatGradAllReducer

    def build():
        torch.manual_seed(1)
        np.random.seed(1)
        cfg = rb.RatSpnConfig()
        cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = 64, 3, 2, 16, 16, 1
        cfg.dropout, cfg.leaf_base_class, cfg.leaf_base_kwargs, cfg.tanh_squash = 0.0, rb.RatNormal, {}, False
        return rb.RatSpn(cfg).to(DEV)

    x = torch.randn(256, 1, 64, 1, 1, device=DEV)
    m1, m2 = build(), build()
    assert m1._tensor_core_path_ok(x, False)
    red = FlatGradAllReducer(m1.parameters())
    opt = FlatAdam(red, lr=1e-2, module=m1)
    topt = torch.optim.Adam(m2.parameters(), lr=1e-2)
    l1, l2 = [], []
    for _ in range(6):
        red.zero_()
        loss = -m1(x).mean()
        loss.backward()
        opt.step()
        l1.append(float(loss))
        topt.zero_grad()
        loss2 = -m2(x).mean()
        loss2.backward()
        topt.step()
        l2.append(float(loss2))
    assert l1[-1] < l1[0]
    import numpy
    numpy.testing.assert_allclose(l1, l2, rtol=2e-4)
