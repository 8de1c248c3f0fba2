"""CPU: host-side logic of the Python mirror -- configuration validation, module structure / state_dict keys,
permutation helpers, the Sample context and argument checking (no kernel is launched)."""
import numpy as np
import pytest
import torch

import ratspn_b200 as rb
from ratspn_b200.type_checks import InvalidTypeException, OutOfBoundsException, check_valid
from tests.helpers import Case


def make_cfg(F=17, D=4, R=3, S=3, I=3, C=1, tanh=True):
    cfg = rb.RatSpnConfig()
    cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = F, D, R, S, I, C
    cfg.dropout, cfg.leaf_base_class, cfg.tanh_squash = 0.0, rb.RatNormal, tanh
    cfg.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
    return cfg


def test_config_guards():
    cfg = rb.RatSpnConfig()
    with pytest.raises(AttributeError):
        cfg.not_a_field = 1
    cfg = make_cfg(F=8, D=4)
    with pytest.raises(Exception):
        cfg.assert_valid()  # 2**D > F (rat_spn.py:76-77)
    cfg = make_cfg(F=3, D=1, R=3)
    with pytest.raises(AssertionError):
        rb.RatSpn(cfg)  # R == F is rejected (rat_spn.py:104)
    cfg = make_cfg()
    assert cfg.F == cfg.in_features == 17


def test_check_valid_semantics():
    assert check_valid(3, int, 1) == 3
    assert check_valid(np.int64(3), int, 1) == 3 and isinstance(check_valid(np.int64(3), int, 1), int)
    assert check_valid(0.0, float, 0.0, 1.0) == 0.0
    assert check_valid(None, int, allow_none=True) is None
    with pytest.raises(InvalidTypeException):
        check_valid(1.5, int)
    with pytest.raises(OutOfBoundsException):
        check_valid(5, int, 1, 5)  # upper bound is exclusive
    with pytest.raises(Exception):
        check_valid(None, int)


def test_layer_stack_and_state_dict_keys_match_reference_fixture():
    c = Case("ratspn_pad_tanh")
    m = rb.RatSpn(make_cfg())
    ref_keys = set(c.state_dict())
    assert set(m.state_dict()) == ref_keys
    for k, v in c.state_dict().items():
        assert tuple(m.state_dict()[k].shape) == tuple(v.shape), k
    m.load_state_dict(c.state_dict(), strict=True)
    kinds = [type(l).__name__ for l in m._inner_layers]
    assert kinds == ["CrossProduct", "Sum", "CrossProduct", "Sum", "CrossProduct", "Sum", "CrossProduct"]
    assert m.max_layer_index == 8 and m.num_layers == 9 and m.sum_layer_indices == [2, 4, 6, 8]
    assert m._leaf.cardinality == 2 and m._leaf.pad == 1 and m._leaf.out_features == 9
    assert m._inner_layers[0]._pad == 7 and m._inner_layers[0].in_features == 16
    assert m.layer_index_to_obj(0) is m._leaf and m.layer_index_to_obj(8) is m.root
    with pytest.raises(IndexError):
        m.layer_index_to_obj(9)
    assert tuple(m.root.weight_param.shape) == (1, 1, 27, 1, 1)
    assert tuple(m.root_weights_split_by_rep.shape) == (1, 1, 9, 1, 3)


def test_cspn_structure_and_state_dict_keys_match_reference_fixture():
    c = Case("cspn_small")
    from tests.helpers import build_cspn
    m = build_cspn(c, "cpu")
    assert set(m.state_dict()) == set(c.state_dict())
    assert not isinstance(m.root.weight_param, torch.nn.Parameter)  # the circuit owns no parameters
    m.set_params(c.t("cond"))
    B = 4
    assert tuple(m.root.weight_param.shape) == (B, 1, 2 * 4, 1, 1)
    assert tuple(m._leaf.base_leaf.mean_param.shape) == (B, 9, 3, 2)
    assert torch.allclose(m.root.weight_param.exp().sum(2), torch.ones(()), atol=1e-5)
    assert torch.allclose(m.root.weight_param, c.t("setparams__lw_root"), atol=1e-6)
    assert torch.allclose(m._leaf.base_leaf.mean_param, c.t("setparams__mu"), atol=1e-6)
    s = m._leaf.base_leaf.std_param
    assert (s >= 0.001).all() and (s <= 1.0).all()
    assert all(l.num_conditionals == B for l in m._inner_layers if isinstance(l, rb.CrossProduct))


def test_permutation_roundtrip_and_numpy_seed_semantics():
    np.random.seed(7)
    m = rb.RatSpn(make_cfg())
    np.random.seed(7)
    expect = np.stack([np.random.permutation(17) for _ in range(3)], axis=-1)
    assert np.array_equal(m.permutation.numpy(), expect)  # same numpy draw order as rat_spn.py:147-149
    x = torch.randn(4, 1, 17, 2, 3)
    assert torch.equal(m.apply_inv_permutation(m.apply_permutation(x)), x)
    x1 = torch.randn(4, 1, 17, 1, 1)
    assert m.apply_permutation(x1).shape == (4, 1, 17, 1, 3)


def test_sample_context_dataclass():
    ctx = rb.Sample(n=(3,), sampling_mode="index")
    assert ctx.is_root
    with pytest.raises(AttributeError):
        ctx.bogus = 1
    ctx.parent_indices = torch.zeros(2, 3, 1, 4, 1, dtype=torch.long)
    ctx.assert_correct_indices()
    assert not ctx.is_root
    oh = rb.Sample(n=(3,), sampling_mode="onehot", parent_indices=torch.eye(4).view(1, 1, 1, 4, 4, 1).expand(2, 3, 1, 4, 4, 1))
    oh.assert_correct_indices()


def test_crossproduct_index_unravelling_and_scope_split():
    xp = rb.CrossProduct(in_features=5, in_channels=3, num_repetitions=2)
    assert xp._pad == 3 and xp.in_features == 8
    assert xp.unraveled_channel_indices.tolist()[:4] == [[0, 0], [0, 1], [0, 2], [1, 0]]
    assert tuple(xp.one_hot_in_channel_mapping.shape) == (9, 2, 3)
    ctx = rb.Sample(n=(1,), sampling_mode="index", parent_indices=torch.tensor([7, 2, 4, 0]).view(1, 1, 1, 4, 1))
    out = xp.sample(mode="index", ctx=ctx).parent_indices.view(-1).tolist()
    assert out == [2, 1, 0, 2, 1]  # (7 -> 2,1), (2 -> 0,2), (4 -> 1,1) truncated to the 5 real scopes
    left, right = xp.split_shuffled_scopes(torch.arange(5.0).view(1, 5, 1), scope_dim=1)
    assert left.view(-1).tolist() == [0, 2, 4, 0] and right.view(-1).tolist() == [1, 3, 0, 0]


def test_forward_argument_checks_happen_before_any_kernel():
    m = rb.RatSpn(make_cfg())
    with pytest.raises(AssertionError):
        m(torch.randn(2, 17))  # fewer than 5 dims (rat_spn.py:213)
    with pytest.raises(AssertionError):
        m(torch.randn(2, 1, 16, 1, 1))  # wrong feature count
    with pytest.raises(AssertionError):
        m(torch.randn(2, 1, 17, 1, 1).double())  # dtype
    with pytest.raises(AssertionError):
        m.sample(mode=None)


def test_cspn_conditioned_block_evaluates_the_gating_network_once():
    """CSPN.conditioned(): calls that pass the pinned tensor re-use the installed parameters; host-only check through
    huber_entropy_lb (parameters -> tensor ops, no kernels)."""
    import ratspn_b200 as rb
    torch.manual_seed(0)
    cfg = rb.CspnConfig()
    cfg.F_cond = (5,)
    cfg.F, cfg.R, cfg.D, cfg.I, cfg.S, cfg.C = 8, 2, 2, 3, 3, 1
    cfg.dropout, cfg.leaf_base_class, cfg.tanh_squash = 0.0, rb.RatNormal, True
    cfg.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
    cfg.feat_layers, cfg.sum_param_layers, cfg.dist_param_layers = [8], [8], [8]
    m = rb.CSPN(cfg)
    cond, other = torch.randn(4, 5), torch.randn(4, 5)
    calls = []
    orig = m.set_params
    m.set_params = lambda x: (calls.append(x), orig(x))[1]
    with m.conditioned(cond):
        a, _ = m.huber_entropy_lb(condition=cond, verbose=False)
        b, _ = m.huber_entropy_lb(condition=cond, verbose=False)
        assert len(calls) == 1 and torch.equal(a, b)
        c, _ = m.huber_entropy_lb(condition=other, verbose=False)   # a different tensor is evaluated as usual
        assert len(calls) == 2 and not torch.equal(a, c)
    assert m._pinned_condition is None
    m.huber_entropy_lb(condition=cond, verbose=False)
    assert len(calls) == 3
    # fused weight normalisation: weight_param holds the raw head outputs, .weights normalises on read
    cfg.fuse_weight_norm = True
    f = rb.CSPN(cfg)
    f.load_state_dict(m.state_dict())
    f.set_params(cond)
    lay = f._inner_layers[1]
    assert lay._raw_weights and not f.root._raw_weights
    assert float(lay.weight_param.exp().sum(2).sub(1).abs().max()) > 1e-3
    assert float(lay.weights.exp().sum(2).sub(1).abs().max()) < 1e-5
    m.set_params(cond)
    assert torch.allclose(lay.weights, m._inner_layers[1].weights, atol=1e-6)


def test_flat_grad_buffer_slices_are_256_byte_aligned_views():
    """Every parameter's .grad is a view into the flat buffer starting on a 64-element (256-byte) boundary (the kernels
    use 16-byte vector / TMA accesses on whole tensors); rebind() restores the views after set_to_none."""
    from ratspn_b200.parallel import FlatGradAllReducer
    ps = [torch.nn.Parameter(torch.randn(s)) for s in [(3, 5), (7,), (64,), (2, 2, 2)]]
    red = FlatGradAllReducer(ps)
    assert red.offsets == [0, 64, 128, 192] and red.numel == 256
    base = red.flat.data_ptr()
    for p, off in zip(ps, red.offsets):
        assert p.grad.data_ptr() == base + 4 * off and p.grad.shape == p.shape
    (ps[0].sum() * 2 + ps[3].sum()).backward()
    assert float(red.flat[:15].sum()) == 30.0 and float(red.flat[192:200].sum()) == 8.0 and float(red.flat[15:192].abs().sum()) == 0.0
    for p in ps:
        p.grad = None
    red.rebind()
    assert all(p.grad.data_ptr() == base + 4 * off for p, off in zip(ps, red.offsets))
    red.zero_()
    assert float(red.flat.abs().sum()) == 0.0
