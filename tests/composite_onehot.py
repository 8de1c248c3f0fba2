"""Tensor-op restatement of the reference's one-hot sampling bookkeeping on the autograd tape (layers.py:181-234,
:495-503, distributions.py:281-301, rat_spn.py:437-468).  TEST SCAFFOLDING: tests/test_gpu_diffsample.py uses it as a
second opinion for the fused straight-through backward kernels of the product (ratspn_b200.diffsample.sample_onehot)."""
from typing import List, Optional

import torch as th

from ratspn_b200 import ops
from ratspn_b200.layers import CrossProduct, Sum
from ratspn_b200.utils import Sample


def _straight_through(logits: th.Tensor, gumbel: Optional[th.Tensor]) -> th.Tensor:
    """One-hot of argmax(logits + gumbel) along dim -2; gradient of softmax(logits + gumbel) (tau = 1).
    MPE (gumbel None): plain one-hot of the argmax, no gradient (layers.py:206-214)."""
    z = logits if gumbel is None else logits + gumbel
    hard = th.zeros_like(z).scatter_(-2, z.argmax(dim=-2, keepdim=True), 1.0)
    if gumbel is None:
        return hard
    soft = th.softmax(z, dim=-2)
    return hard - soft.detach() + soft


def _sum_input_from_fused_cache(layer: Sum) -> Optional[th.Tensor]:
    """Posterior offset of a Sum layer = its input on the evidence = CrossProduct of the cached activations."""
    if not layer._is_input_cache_enabled or layer._input_cache_fused is None:
        return None
    return ops.cross_product(layer._input_cache_fused)


def sample_onehot_composite(spn, ctx: Sample, class_index, evidence, layer_index: int,
                            noise: Optional[List[th.Tensor]]) -> Sample:
    """Tensor-op restatement of the reference's one-hot bookkeeping on the autograd tape.  Not on the product path:
    tests/test_gpu_diffsample.py uses it as a second opinion for the fused backward kernels."""
    cfg = spn.config
    dev = spn.device
    n = tuple(ctx.n)
    W = spn._num_weight_sets()
    w = evidence.shape[-4] if evidence is not None else (class_index.shape[0] if class_index is not None else W)

    def gumbel(shape):
        if ctx.is_mpe:
            return None
        if noise is not None:
            t = noise.pop(0)
            assert tuple(t.shape) == tuple(shape), f"noise shape {tuple(t.shape)} != {tuple(shape)}"
            return t.to(dev)
        return -th.empty(shape, device=dev).exponential_().log()

    def expand_nodes(lw):  # [W, d, ic, oc, R] -> [oc, *n, w, d, ic, R]
        Wl, d, ic, oc, R = lw.shape
        t = lw.expand(w, d, ic, oc, R).permute(3, 0, 1, 2, 4)
        return t.reshape(oc, *([1] * len(n)), w, d, ic, R).expand(oc, *n, w, d, ic, R)

    start = layer_index
    has_rep_dim = True
    oh = None
    if layer_index == spn.max_layer_index:
        ctx.scopes = 1
        lwr = spn.root.weights  # [W, 1, K, C, 1]
        K = lwr.shape[2]
        logits = expand_nodes(lwr)
        if class_index is not None:  # rat_spn.py:437-442: one fixed root node per conditional
            sel = th.nn.functional.one_hot(th.as_tensor(class_index, device=dev), cfg.C).to(lwr.dtype)  # [w, C]
            logits = (logits * sel.t().reshape(cfg.C, *([1] * len(n)), w, 1, 1, 1)).sum(0, keepdim=True)
        cache = _sum_input_from_fused_cache(spn.root)
        if cache is not None:  # [*evb, w, 1, c*c, R] -> channels r*ic + i
            logits = logits + cache.transpose(-1, -2).flatten(-2, -1).unsqueeze(-1)
        oh = _straight_through(logits, gumbel(logits.shape))  # [nodes, *n, w, 1, K, 1]
        oh = oh.reshape(*oh.shape[:-2], cfg.R, K // cfg.R).transpose(-1, -2)  # [.., ic, R]
        start = layer_index - 1
        has_rep_dim = False
    elif layer_index > 0:
        layer = spn.layer_index_to_obj(layer_index)
        if isinstance(layer, CrossProduct):
            ctx.scopes = layer.in_features // layer.cardinality
            cc = layer.in_channels ** 2
            eye = th.eye(cc, device=dev)
            oh = eye.view(cc, *([1] * len(n)), 1, 1, cc, 1).expand(cc, *n, w, ctx.scopes, cc, cfg.R)
        else:
            ctx.scopes = layer.in_features
            logits = expand_nodes(layer.weights)
            cache = _sum_input_from_fused_cache(layer)
            if cache is not None:
                logits = logits + cache
            oh = _straight_through(logits, gumbel(logits.shape))
            start = layer_index - 1
    ctx.has_rep_dim = has_rep_dim

    for li in range(start, 0, -1):
        layer = spn._inner_layers[li - 1]
        if isinstance(layer, CrossProduct):  # layers.py:495-508
            c = layer.in_channels
            t = oh.reshape(*oh.shape[:-2], c, c, oh.shape[-1])  # [.., d, l, r', r]
            oh = th.stack((t.sum(dim=-2), t.sum(dim=-3)), dim=-3)  # [.., d, 2, c, r]
            oh = oh.reshape(*oh.shape[:-4], oh.shape[-4] * 2, c, oh.shape[-1])
            if layer._pad:
                oh = oh[..., :-layer._pad, :, :]
        else:  # layers.py:181-186, :198-203, :233-234
            lw = layer.weights
            logits = (lw * oh.unsqueeze(-3)).sum(-2)  # [.., d, ic, r]
            unused = (oh.sum(-2, keepdim=True) == 0.0).expand_as(logits)
            cache = _sum_input_from_fused_cache(layer)
            if cache is not None:
                logits = logits + (cache.unsqueeze(-2) * oh.unsqueeze(-3)).sum(-2)
            new = _straight_through(logits, gumbel(logits.shape))
            oh = th.where(unused, th.zeros_like(new), new)

    base = spn._leaf.base_leaf
    mu, sigma = base.means, base.stds
    if layer_index == 0:
        assert len(n) == 1
        mu_e = mu.expand(*n, w, *mu.shape[1:])
        if ctx.is_mpe:
            y = mu_e
        else:
            eps = noise.pop(0).to(dev) if noise is not None else th.randn(mu_e.shape, device=dev)
            y = mu_e + sigma.expand_as(mu_e) * eps
        ctx.scopes = spn._leaf.in_features // spn._leaf.cardinality
        ctx.sample = y.permute(3, 0, 1, 2, 4)
        return ctx
    oh_f = th.repeat_interleave(oh, spn._leaf.cardinality, dim=-3)[..., :cfg.F, :, :]  # [.., F, I, r]
    mu_s = (mu * oh_f).sum(-2)
    sg_s = (sigma * oh_f).sum(-2)
    if not has_rep_dim:
        mu_s, sg_s = mu_s.sum(-1), sg_s.sum(-1)
    if ctx.is_mpe:
        y = mu_s
    else:
        eps = noise.pop(0).to(dev) if noise is not None else th.randn(mu_s.shape, device=dev)
        y = mu_s + sg_s * eps
    ctx.parent_indices = oh_f
    ctx.sample = y
    return ctx
