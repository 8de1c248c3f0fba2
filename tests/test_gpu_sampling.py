"""GPU: the in-channel-contiguous weight copy used when many samples are drawn from wide shared-weight layers
(ops.sampling_weights) must give bit-identical indices to the parameter layout, for inner layers and the root."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("mpe", [False, True])
def test_sample_sum_layouts_agree_inner_and_root(mpe):
    from ratspn_b200 import ops
    g = torch.Generator().manual_seed(0)
    d, c, oc, R, Q = 4, 9, 7, 3, 150
    ic = c * c
    lw = torch.log_softmax(torch.randn(1, d, ic, oc, R, generator=g), dim=2).to(DEV)
    # parent choices of the CrossProduct above (pairs of this layer's out-channels), one repetition per sample
    pa = torch.randint(0, oc * oc, (Q, d // 2, 1), generator=g).to(torch.int32).to(DEV)
    rep = torch.randint(0, R, (Q,), generator=g).to(torch.int32).to(DEV)
    gum = None if mpe else (-torch.empty(Q, d, ic, 1).exponential_(generator=g).log()).to(DEV)
    xc = torch.randn(Q, 2 * d, c, R, generator=g).to(DEV)  # evidence activations of the children
    a = ops.sample_sum(lw, 0, Q, d, 1, pa, (d // 2, 1, 0), True, rep, xc, Q, gum, mpe)
    b = ops.sample_sum(ops.sampling_weights(lw), 0, Q, d, 1, pa, (d // 2, 1, 0), True, rep, xc, Q, gum, mpe, transposed=True)
    assert torch.equal(a, b)
    # root: K = R * ic in-channels, C classes
    C = 5
    lwr = torch.log_softmax(torch.randn(1, 1, R * ic, C, 1, generator=g), dim=2).to(DEV)
    node = torch.randint(0, C, (Q,), generator=g).to(torch.int32).to(DEV)
    gr = None if mpe else (-torch.empty(Q, 1, R * ic, 1).exponential_(generator=g).log()).to(DEV)
    a = ops.sample_sum(lwr, R, Q, 1, 1, node, (1, 0, 0), False, None, None, 1, gr, mpe)
    b = ops.sample_sum(ops.sampling_weights(lwr), R, Q, 1, 1, node, (1, 0, 0), False, None, None, 1, gr, mpe, transposed=True)
    assert torch.equal(a, b)


def test_model_mpe_is_independent_of_the_number_of_samples():
    """n = 1 uses the parameter layout, n = 128 the in-channel-contiguous copy: same MPE sample."""
    import ratspn_b200 as rb
    torch.manual_seed(0)
    np.random.seed(0)
    cfg = rb.RatSpnConfig()
    cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = 32, 3, 3, 9, 8, 2
    cfg.dropout, cfg.leaf_base_class, cfg.leaf_base_kwargs, cfg.tanh_squash = 0.0, rb.RatNormal, {}, False
    m = rb.RatSpn(cfg).to(DEV)
    with torch.no_grad():
        one = m.sample(mode="index", n=1, is_mpe=True).sample      # [C, 1, 1, F, 1]
        many = m.sample(mode="index", n=128, is_mpe=True).sample   # [C, 128, 1, F, 1]
    assert torch.equal(many, one.expand_as(many))


def test_in_kernel_noise_four_per_philox_block_is_unbiased_and_deterministic():
    """Rs = 1, ic % 4 == 0 takes one Philox block per four in-channels.  Uniform logits -> uniform choices (a reused or
    shifted noise word would bias them); same seed -> same draws; different seed -> different draws."""
    from ratspn_b200 import ops
    d, ic, oc, R, Q = 2, 16, 4, 3, 60000
    lw = torch.full((1, d, ic, oc, R), -float(np.log(ic)), device=DEV)
    pa = torch.zeros(Q, 1, 1, dtype=torch.int32, device=DEV)
    rep = torch.zeros(Q, dtype=torch.int32, device=DEV)
    a = ops.sample_sum(lw, 0, Q, d, 1, pa, (1, 1, 0), True, rep, None, 1, None, False, seed=1234)
    b = ops.sample_sum(lw, 0, Q, d, 1, pa, (1, 1, 0), True, rep, None, 1, None, False, seed=1234)
    c = ops.sample_sum(lw, 0, Q, d, 1, pa, (1, 1, 0), True, rep, None, 1, None, False, seed=99)
    assert torch.equal(a, b) and not torch.equal(a, c)
    counts = torch.bincount(a.flatten().long(), minlength=ic).double().cpu().numpy()
    expected = Q * d / ic
    chi2 = float(((counts - expected) ** 2 / expected).sum())
    assert chi2 < 50.0, (chi2, counts)      # 15 degrees of freedom: P(chi2 > 50) ~ 1e-5
    # neighbouring samples / scopes are independent: the two scopes of one sample agree with probability 1 / ic
    agree = float((a[:, 0, 0] == a[:, 1, 0]).double().mean())
    assert abs(agree - 1.0 / ic) < 0.01


@pytest.mark.parametrize("Rs,ic,with_xc", [(4, 16, False), (8, 64, True), (4, 9, True)])
def test_four_repetitions_per_group_variant_draws_the_same_indices(Rs, ic, with_xc, monkeypatch):
    """Layers sampled in every repetition with in-kernel noise (entropy path, R % 4 == 0) use one Philox block for four
    repetitions; indices must equal those of the one-repetition-per-group kernel (``one_rep_per_group=True``) for the same seed."""
    from ratspn_b200 import ops
    g = torch.Generator().manual_seed(Rs * 100 + ic)
    c = int(round(ic ** 0.5))
    d, oc, Q, w = 4, 5, 333, 3
    lw = torch.log_softmax(torch.randn(w, d, ic, oc, Rs, generator=g), dim=2).to(DEV)   # per-conditional weights
    pa = torch.randint(0, oc * oc, (Q, d // 2, Rs), generator=g).to(torch.int32).to(DEV)
    xc = torch.randn(Q, 2 * d, c, Rs, generator=g).to(DEV) if with_xc else None
    args = (lw, 0, Q, d, Rs, pa, (d // 2 * Rs, Rs, 1), True, None, xc, Q, None, False)
    a = ops.sample_sum(*args, seed=4242, q_offset=17)
    b = ops.sample_sum(*args, seed=4242, q_offset=17, one_rep_per_group=True)
    assert a.shape == (Q, d, Rs) and torch.equal(a, b)
    assert int(a.min()) >= 0 and int(a.max()) < ic and a.float().std() > 0
