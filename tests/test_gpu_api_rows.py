"""GPU: API rows of SURVEY 8 that had no test in round 1 -- batch-sharded sampling (8e), dropout (layers.py:105-108,
distributions.py:51-56), split_by_scope post-processing (rat_spn.py:523-536), naive_entropy_approx (rat_spn.py:799-823)."""
import numpy as np
import pytest
import torch

from oracle import spn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ratspn(F=32, D=3, R=4, S=6, I=5, C=1, dropout=0.0, seed=0):
    import ratspn_b200 as rb
    torch.manual_seed(seed)
    np.random.seed(seed)
    cfg = rb.RatSpnConfig()
    cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = F, D, R, S, I, C
    cfg.dropout, cfg.leaf_base_class, cfg.leaf_base_kwargs, cfg.tanh_squash = dropout, rb.RatNormal, {}, False
    return rb.RatSpn(cfg).to(DEV)


def test_sharded_index_sampling_reproduces_the_single_gpu_draw():
    """the in-kernel Philox noise is keyed by (seed, global sample index): two shards drawing samples [0, 40) and
    [40, 96) with the job's seed return exactly the rows of the single call -- indices and values bit-identical."""
    m = _ratspn(F=64, D=4, R=5, S=8, I=8)
    full = m.sample(mode="index", n=96, seed=1234)
    a = m.sample(mode="index", n=40, seed=1234, sample_offset=0)
    b = m.sample(mode="index", n=56, seed=1234, sample_offset=40)
    assert torch.equal(torch.cat([a.sample, b.sample], dim=1), full.sample)
    assert torch.equal(torch.cat([a.parent_indices, b.parent_indices], dim=1), full.parent_indices)
    assert torch.equal(torch.cat([a.repetition_indices, b.repetition_indices], dim=1), full.repetition_indices)
    other = m.sample(mode="index", n=96, seed=99)
    assert not torch.equal(other.sample, full.sample)


def test_sum_and_leaf_dropout_follow_the_reference_semantics():
    """Training mode with dropout p: a Sum layer sets random inputs to -inf (layers.py:105-108), a leaf sets random
    log-densities to 0 (distributions.py:51-56); eval mode applies neither.  The masks are reproduced from the same
    generator state and applied to the oracle's layer functions."""
    import ratspn_b200 as rb
    p = 0.3
    layer = rb.Sum(in_channels=6, in_features=4, out_channels=5, ratspn=True, num_repetitions=3, dropout=p).to(DEV)
    x = torch.randn(7, 1, 4, 6, 3, device=DEV)
    layer.train()
    torch.manual_seed(11)
    mask = torch.rand_like(x) < p
    torch.manual_seed(11)
    out = layer(x)
    ref = O.sum_layer(x.masked_fill(mask, float("-inf")).cpu(), torch.log_softmax(layer.weight_param.detach().cpu(), dim=2))
    torch.testing.assert_close(out.detach().cpu(), ref, rtol=1e-5, atol=1e-5)
    assert 0.2 < mask.float().mean().item() < 0.4
    layer.eval()
    ref = O.sum_layer(x.cpu(), torch.log_softmax(layer.weight_param.detach().cpu(), dim=2))
    torch.testing.assert_close(layer(x).detach().cpu(), ref, rtol=1e-5, atol=1e-5)

    leaf = rb.RatNormal(in_features=8, out_channels=4, ratspn=True, num_repetitions=3, dropout=p).to(DEV)
    xl = torch.randn(5, 1, 8, 1, 1, device=DEV)
    leaf.eval()
    dense = leaf(xl)
    leaf.train()
    torch.manual_seed(3)
    lmask = torch.rand_like(dense) < p
    torch.manual_seed(3)
    dropped = leaf(xl)
    torch.testing.assert_close(dropped, dense.masked_fill(lmask, 0.0))


def test_model_forward_with_dropout_runs_layer_by_layer_and_is_stochastic_only_in_training():
    m = _ratspn(dropout=0.2)
    x = torch.randn(300, 1, 32, 1, 1, device=DEV)
    m.eval()
    a, b = m(x), m(x)
    assert torch.equal(a, b)
    m.train()
    torch.manual_seed(0)
    c = m(x)
    torch.manual_seed(1)
    d = m(x)
    assert torch.isfinite(c).all() and not torch.equal(c, d)
    torch.manual_seed(0)
    assert torch.equal(m(x), c)          # same generator state, same masks
    (-c.mean()).backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)


def test_split_by_scope_postprocessing():
    """rat_spn.py:523-536: samples drawn at an inner layer are repeated once per scope of that layer and everything outside
    the scope is NaN; the kept entries are the plain post-processed sample."""
    m = _ratspn(F=32, D=3, R=4, S=6, I=5)
    torch.manual_seed(5)
    ctx = m.sample(mode="index", n=3, layer_index=2, do_sample_postprocessing=False)
    scopes = ctx.scopes
    assert scopes == m.layer_index_to_obj(2).in_features == 4
    raw = ctx.sample.clone()
    plain = m.sample_postprocessing(
        type(ctx)(**{**ctx.__dict__, "sample": raw.clone()}), split_by_scope=False).sample
    split = m.sample_postprocessing(ctx, split_by_scope=True)
    assert split.is_split_by_scope and split.sample.shape[1] == scopes
    s = split.sample
    assert s.shape[0] == plain.shape[0] and s.shape[2:] == plain.shape[1:]
    finite = torch.isfinite(s)
    # every feature belongs to exactly one scope, and where it is kept the value is the unsplit sample's
    assert torch.equal(finite.sum(dim=1), torch.ones_like(finite.sum(dim=1)))
    merged = torch.nan_to_num(s, nan=0.0).sum(dim=1)
    torch.testing.assert_close(merged, plain)


def test_naive_entropy_approx_is_a_monte_carlo_estimate_of_the_entropy():
    """rat_spn.py:799-823: -E[log p(x)] from the model's own samples.  Checked against the closed form for a model whose
    mixture components coincide (all leaves N(0.3, 0.5^2) -> a product of F independent Gaussians), and against the
    Huber lower bound on a random model."""
    import math
    m = _ratspn(F=16, D=2, R=2, S=3, I=3)
    base = m._leaf.base_leaf
    with torch.no_grad():
        base.mean_param.fill_(0.3)
        base.std_param.fill_(0.1)
        sig = float(base.stds.flatten()[0])   # after the bounding of distributions.py:366-382
        assert float((base.stds - sig).abs().max()) == 0.0
    exact = 16 * 0.5 * math.log(2 * math.pi * math.e * sig ** 2)
    torch.manual_seed(0)
    est = m.naive_entropy_approx(sample_size=4000).item()
    assert abs(est - exact) < 0.15, (est, exact)

    m2 = _ratspn(F=16, D=2, R=2, S=3, I=3, seed=4)
    torch.manual_seed(1)
    est2 = m2.naive_entropy_approx(sample_size=4000).item()
    lb, _ = m2.huber_entropy_lb(verbose=False)
    assert float(lb.flatten()[0]) <= est2 + 0.3
