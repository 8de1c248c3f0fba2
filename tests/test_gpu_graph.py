"""The hot path is CUDA-graph capturable (graphs.GraphedCall): a captured LL forward/backward replays to the eager
result, captured sampling draws fresh noise on every replay (device-resident seed), and the tensor-core operand images
are rebuilt inside the graph so that in-graph weight updates are seen by the next replay."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cspn():
    import ratspn_b200 as rb
    torch.manual_seed(0)
    np.random.seed(0)
    cfg = rb.CspnConfig()
    cfg.F_cond = (12,)
    cfg.F, cfg.R, cfg.D, cfg.I, cfg.S, cfg.C = 9, 3, 3, 3, 3, 1
    cfg.dropout = 0.0
    cfg.leaf_base_class = rb.RatNormal
    cfg.leaf_base_kwargs = {"min_sigma": 0.01, "max_sigma": 1.0}
    cfg.tanh_squash = True
    cfg.feat_layers, cfg.sum_param_layers, cfg.dist_param_layers = [16], [16], [16]
    cfg.cond_layers_inner_act = torch.nn.LeakyReLU
    return rb.CSPN(cfg).to(DEV)


def test_graphed_policy_update_matches_eager_and_resamples():
    import ratspn_b200 as rb
    m = _cspn()
    B = 32
    obs = torch.randn(B, 12, device=DEV)
    xs = torch.randn(1, B, 9, 1, 1, device=DEV).clamp(-2, 2)
    ev = torch.full((1, B, 9, 1, 1), float("nan"), device=DEV)
    ev[0, ::3, 2] = 0.25
    failed = ~torch.isnan(ev.view(B, 9))
    params = [p for p in m.parameters() if p.requires_grad]

    def ll_step():
        for p in params:
            p.grad = None if p.grad is None else p.grad.zero_()
        ll = m(xs, condition=obs)
        (-ll.mean()).backward()
        return ll

    ll_eager = ll_step().detach().clone()
    g_eager = [p.grad.clone() for p in params]
    del ll_step  # (and with it nothing of the eager autograd graph stays alive)

    def ll_step2():
        for p in params:
            p.grad.zero_()
        ll = m(xs, condition=obs)
        (-ll.mean()).backward()
        return ll

    g = rb.GraphedCall(ll_step2)
    out = g()
    torch.cuda.synchronize()
    torch.testing.assert_close(out, ll_eager, rtol=1e-6, atol=1e-6)
    for p, ge in zip(params, g_eager):
        torch.testing.assert_close(p.grad, ge, rtol=1e-4, atol=1e-6)
    obs.copy_(torch.randn(B, 12, device=DEV))  # a new batch through the static input
    out2 = g().clone()
    assert not torch.allclose(out2, ll_eager)

    def update():
        for p in params:
            p.grad.zero_()
        ctx = m.sample(mode="onehot", condition=obs, evidence=ev)
        H, _ = m.recursive_entropy_approx(condition=obs, sample_size=3, marginal_mask=failed)
        act = ctx.sample.reshape(B, 9)
        (-(0.1 * H + act.sum(-1)).mean()).backward()
        return act, H

    gu = rb.GraphedCall(update)
    a1, h1 = [t.clone() for t in gu()]
    g1 = [p.grad.clone() for p in params]
    a2, h2 = [t.clone() for t in gu()]
    torch.cuda.synchronize()
    assert torch.isfinite(a1).all() and torch.isfinite(h1).all() and all(torch.isfinite(t).all() for t in g1)
    assert not torch.equal(a1, a2), "replays must draw fresh noise"
    filled = failed.view(B, 9)
    torch.testing.assert_close(a1[filled], torch.tanh(ev.view(B, 9))[filled])  # evidence is kept in every replay
    assert (a1.abs() <= 1).all()


def test_graphed_tensor_core_training_steps_see_weight_updates():
    import ratspn_b200 as rb

    def build():
        torch.manual_seed(1)
        np.random.seed(1)
        cfg = rb.RatSpnConfig()
        cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = 64, 3, 2, 16, 16, 1
        cfg.dropout, cfg.leaf_base_class, cfg.leaf_base_kwargs, cfg.tanh_squash = 0.0, rb.RatNormal, {}, False
        return rb.RatSpn(cfg).to(DEV)

    x = torch.randn(256, 1, 64, 1, 1, device=DEV)
    lr = 0.05

    def make_step(m):
        params = [p for p in m.parameters() if p.requires_grad]
        for p in params:
            p.grad = torch.zeros_like(p)

        def step():
            for p in params:
                p.grad.zero_()
            ll = m(x)
            (-ll.mean()).backward()
            with torch.no_grad():
                for p in params:
                    p.add_(p.grad, alpha=-lr)
            return ll.detach().mean()
        return step

    m_e = build()
    assert m_e._tensor_core_path_ok(x, False)
    s_e = make_step(m_e)
    eager = [float(s_e()) for _ in range(4 + 3)]  # 3 warm-up calls + capture call happen before the replays
    m_g = build()
    g = rb.GraphedCall(make_step(m_g), warmup=3)  # 3 warm-up steps + 1 captured (not executed) step
    graphed = [float(g()) for _ in range(4)]
    torch.cuda.synchronize()
    # the capture pass itself does not execute, so replay i corresponds to eager step 3 + i
    np.testing.assert_allclose(graphed, eager[3:7], rtol=2e-4)
    assert graphed[-1] > graphed[0]
