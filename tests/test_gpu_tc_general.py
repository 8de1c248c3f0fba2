"""GPU: the generalised tensor-core Sum path -- any channel counts that are multiples of 4 up to 64 (c != oc included,
padded internally to a multiple of 8), streamed weight images for wide layers (P >= 48), and several weight sets with
many rows per set (per-conditional weights of a CSPN) -- against the CPU oracle (oracle/spn_oracle.py: cross_product +
sum_layer and their autograd), evaluated in row micro-batches.

Tolerances (bf16 operands, fp32 accumulation): forward <= 2e-2 nats absolute (typically 1e-3) on log-likelihoods of
magnitude 10 - 100; gradients <= 5e-3 of the tensor's largest magnitude, mean error <= 5e-4 of it.
"""
import numpy as np
import pytest
import torch

from oracle import spn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def to_ref_layout(x_drbc, nsets):
    """internal [d, R, nsets*Bs, c] (set-major rows) -> reference layout [Bs, nsets, d, c, R]."""
    d, R, Btot, c = x_drbc.shape
    return x_drbc.view(d, R, nsets, Btot // nsets, c).permute(3, 2, 0, 4, 1).contiguous()


def from_ref_layout(t):
    """[Bs, nsets, d, c, R] -> [d, R, nsets*Bs, c]."""
    Bs, nsets, d, c, R = t.shape
    return t.permute(2, 4, 1, 0, 3).reshape(d, R, nsets * Bs, c).contiguous()


def oracle_forward_backward(x_drbc, lw, go, micro=32):
    """out and the gradients of sum(out * go) w.r.t. x and lw through the oracle, rows in micro-batches."""
    nsets = lw.shape[0]
    x = to_ref_layout(x_drbc.cpu(), nsets).requires_grad_(True)
    lwc = lw.cpu().clone().requires_grad_(True)
    g = to_ref_layout(go.cpu(), nsets)
    outs = []
    for i in range(0, x.shape[0], micro):
        o = O.sum_layer(O.cross_product(x[i:i + micro]), lwc)
        (o * g[i:i + micro]).sum().backward()
        outs.append(o.detach())
    return from_ref_layout(torch.cat(outs)), from_ref_layout(x.grad), lwc.grad


def check(got, ref, what, tol=5e-3, mean_tol=5e-4):
    got, ref = got.detach().cpu().double(), ref.double()
    scale = ref.abs().max().item() + 1e-30
    err = (got - ref).abs()
    assert err.max().item() <= tol * scale, f"{what}: max err {err.max().item() / scale:.3e} of max-norm"
    assert err.mean().item() <= mean_tol * scale, f"{what}: mean err {err.mean().item() / scale:.3e} of max-norm"


CASES = [  # c, oc, d_in, R, Bs, nsets
    (8, 8, 2, 3, 200, 1), (16, 16, 3, 2, 130, 1), (12, 20, 4, 2, 300, 1), (20, 24, 2, 3, 257, 1), (24, 16, 2, 2, 140, 1),
    (32, 32, 2, 2, 257, 1), (40, 40, 4, 2, 300, 1), (40, 8, 2, 2, 129, 1), (48, 48, 2, 1, 260, 1), (56, 40, 2, 1, 140, 1),
    (64, 64, 2, 2, 200, 1), (16, 16, 4, 2, 150, 3), (24, 32, 2, 2, 131, 2), (64, 64, 2, 1, 129, 2),
]


@pytest.mark.parametrize("c,oc,d_in,R,Bs,nsets", CASES)
def test_tc_forward_backward_general_shapes_vs_oracle(c, oc, d_in, R, Bs, nsets):
    from ratspn_b200 import ops
    assert ops.tc_supported(c, oc, R)
    g = torch.Generator().manual_seed(c * 1000 + oc * 10 + nsets)
    d_out = max(1, (1 << (d_in - 1).bit_length()) // 2)
    x = torch.randn(d_in, R, nsets * Bs, c, generator=g) * 4.0 - 30.0
    lw = torch.log_softmax(torch.randn(nsets, d_out, c * c, oc, R, generator=g) * 1.5, dim=2)
    go = torch.randn(d_out, R, nsets * Bs, oc, generator=g)
    ref_out, ref_dx, ref_dlw = oracle_forward_backward(x, lw, go)
    xt = x.to(DEV).requires_grad_(True)
    lwt = lw.to(DEV).requires_grad_(True)
    out = ops.sum_forward_tc(xt, lwt)
    (out * go.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    assert ops.tc_aborts() == 0
    err = (out.detach().cpu() - ref_out).abs()
    assert err.max().item() < 2e-2 and err.mean().item() < 2e-3, f"forward: max {err.max().item():.3e} mean {err.mean().item():.3e}"
    check(xt.grad, ref_dx, f"dx c={c} oc={oc} sets={nsets}")
    check(lwt.grad, ref_dlw, f"dlw c={c} oc={oc} sets={nsets}")
    # an impossible left scope (all channels -inf): -inf output like th.logsumexp, no NaN, and no gradient from that row
    xi = x.clone()
    xi[0, 0, 3, :] = float("-inf")
    with torch.no_grad():
        ro = from_ref_layout(O.sum_layer(O.cross_product(to_ref_layout(xi, nsets)[:8]), lw))
    xit = xi.to(DEV).requires_grad_(True)
    oi = ops.sum_forward_tc(xit, lw.to(DEV))
    (oi * go.to(DEV)).sum().backward()
    o8 = oi.detach().cpu().view(d_out, R, nsets, Bs, oc)[:, :, :, :8].reshape(d_out, R, nsets * 8, oc)
    assert torch.equal(torch.isinf(o8), torch.isinf(ro)) and not torch.isnan(oi).any()
    assert torch.isinf(oi[0, 0, 3]).all() and (oi[0, 0, 3] < 0).all()
    assert not torch.isnan(xit.grad).any() and xit.grad[:2, 0, 3].abs().max().item() == 0.0


@pytest.mark.parametrize("c,oc,R,Bs,nsets", [(20, 12, 3, 300, 1), (48, 48, 2, 260, 2)])
def test_tc_fused_weight_normalisation_general_shapes(c, oc, R, Bs, nsets):
    """raw = True on padded / streamed shapes: unnormalised parameters in, log_softmax forward and backward fused."""
    from ratspn_b200 import ops
    g = torch.Generator().manual_seed(c + oc)
    d_in, d_out = 2, 1
    x = torch.randn(d_in, R, nsets * Bs, c, generator=g) * 3.0 - 20.0
    raw = torch.randn(nsets, d_out, c * c, oc, R, generator=g) * 1.5 + 0.3
    go = torch.randn(d_out, R, nsets * Bs, oc, generator=g)
    rawc = raw.clone().requires_grad_(True)
    xr = to_ref_layout(x, nsets)
    gr = to_ref_layout(go, nsets)
    for i in range(0, Bs, 32):
        o = O.sum_layer(O.cross_product(xr[i:i + 32]), torch.log_softmax(rawc, dim=2))
        (o * gr[i:i + 32]).sum().backward()
    pt = raw.to(DEV).requires_grad_(True)
    out = ops.sum_forward_tc(x.to(DEV), pt, True, {})
    (out * go.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    # the fused softmax backward subtracts softmax * column-sum from the linear-domain gradient: the bf16 operand errors of
    # both terms add up, so the bound is twice the plain weight gradient's
    check(pt.grad, rawc.grad, f"d raw c={c} oc={oc}", tol=1e-2, mean_tol=1e-3)


def test_model_with_unequal_leaf_and_sum_channels_takes_the_tensor_core_path():
    """I = 20, S = 24 (RatSpnConfig allows I != S, rat_spn.py:43-53): the bottom layer pairs leaf channels (c = 20 -> oc = 24),
    the upper ones sum channels (24 -> 24); LL and gradients vs the oracle."""
    import ratspn_b200 as rb
    torch.manual_seed(7)
    np.random.seed(7)
    F, D, R, S, I, C = 64, 3, 3, 24, 20, 2
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
    x[5, 0, ::4] = float("nan")
    ref = O.forward(O.params_from_ratspn_state(sd, arch), arch, x, microbatch=None)
    (-ref.mean()).backward()
    m = m.to(DEV)
    assert m._tensor_core_path_ok(x.to(DEV), False), "I = 20 / S = 24 must take the tensor-core path"
    out = m(x.to(DEV))
    rel = ((out.cpu() - ref).abs() / ref.abs().clamp_min(1.0)).max().item()
    assert rel < 1e-4, f"relative LL error {rel}"
    (-out.mean()).backward()
    for k, p in m.named_parameters():
        if p.requires_grad:
            # weight parameters are stored unnormalised: their gradient is the difference of two nearly equal terms (softmax
            # backward), which doubles the bf16 operand error of the activations below them
            tol = 2e-2 if k.endswith("weight_param") else 1e-2
            check(p.grad, raw[k].grad, k, tol=tol, mean_tol=1e-3)


@pytest.mark.parametrize("fuse", [False, True])
def test_cspn_with_many_rows_per_conditional_takes_the_tensor_core_path(fuse):
    """CSPN (gating net -> per-conditional weights, cspn.py:254-301) evaluated on 300 rows per conditional: the inner Sum
    layers run on the tcgen05 kernels batched over the weight sets.  LL and the gradients w.r.t. the gating network
    against the same model on the exact fp32 kernels (which are pinned to the reference's golden vectors)."""
    import ratspn_b200 as rb
    torch.manual_seed(5)
    np.random.seed(5)
    cfg = rb.CspnConfig()
    cfg.F_cond = (6,)
    cfg.F, cfg.R, cfg.D, cfg.I, cfg.S, cfg.C = 32, 3, 3, 8, 16, 1
    cfg.dropout = 0.0
    cfg.leaf_base_class = rb.RatNormal
    cfg.leaf_base_kwargs = {"min_sigma": 0.05, "max_sigma": 1.0}
    cfg.tanh_squash = False
    cfg.feat_layers, cfg.sum_param_layers, cfg.dist_param_layers = [8], [8], [8]
    cfg.cond_layers_inner_act = torch.nn.LeakyReLU
    cfg.fuse_weight_norm = fuse
    m = rb.CSPN(cfg).to(DEV)
    w, n = 3, 300
    cond = torch.randn(w, 6, device=DEV)
    x = torch.randn(n, w, 32, 1, 1, device=DEV)
    x[7, 1, ::5] = float("nan")
    res = []
    for tc in (True, False):
        m.use_tensor_cores = tc
        m.zero_grad(set_to_none=True)
        m.set_params(cond)
        assert m._tensor_core_sets_path_ok(x, False, m.max_layer_index) == tc
        ll = m(x, condition=None)
        (-ll.mean()).backward()
        res.append((ll.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
    (ll_t, g_t), (ll_e, g_e) = res
    rel = ((ll_t - ll_e).abs() / ll_e.abs().clamp_min(1.0)).max().item()
    assert rel < 1e-4, f"relative LL error {rel}"
    assert set(g_t) == set(g_e)
    for k in g_e:
        check(g_t[k], g_e[k].cpu(), k, tol=2e-2, mean_tol=2e-3)


def test_cspn_recursive_entropy_with_gradients_takes_the_tensor_core_path():
    """rat_spn.py:625-750 on a CSPN: every node's samples (K nodes x n samples per conditional = 320 rows per weight set) are
    scored under the circuit below the layer -- dense per-conditional contractions, on the tensor-core kernels batched over
    the weight sets.  Entropy value and its gradients w.r.t. the gating network against the exact fp32 kernels, same noise."""
    import ratspn_b200 as rb
    torch.manual_seed(9)
    np.random.seed(9)
    cfg = rb.CspnConfig()
    cfg.F_cond = (6,)
    cfg.F, cfg.R, cfg.D, cfg.I, cfg.S, cfg.C = 32, 2, 3, 16, 16, 1
    cfg.dropout = 0.0
    cfg.leaf_base_class = rb.RatNormal
    cfg.leaf_base_kwargs = {"min_sigma": 0.05, "max_sigma": 1.0}
    cfg.tanh_squash = False
    cfg.feat_layers, cfg.sum_param_layers, cfg.dist_param_layers = [8], [8], [8]
    cfg.cond_layers_inner_act = torch.nn.LeakyReLU
    m = rb.CSPN(cfg).to(DEV)
    cond = torch.randn(2, 6, device=DEV)
    res = []
    for tc in (True, False):
        m.use_tensor_cores = tc
        m.zero_grad(set_to_none=True)
        torch.manual_seed(21)
        H, _ = m.recursive_entropy_approx(condition=cond, sample_size=20)
        H.sum().backward()
        res.append((H.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
    (H_t, g_t), (H_e, g_e) = res
    assert float((H_t - H_e).abs().max()) <= 2e-3 * float(H_e.abs().max()) + 1e-3, (H_t, H_e)
    assert set(g_t) == set(g_e)
    for k in g_e:
        check(g_t[k], g_e[k].cpu(), k, tol=3e-2, mean_tol=3e-3)
