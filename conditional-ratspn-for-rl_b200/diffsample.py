"""Differentiable ('onehot') ancestral sampling: same forward values as index mode, gradients by the straight-through
Gumbel-softmax estimator (reference: layers.py:181-186, :206-234, :495-503, distributions.py:281-301,
rat_spn.py:437-468).

``sample_onehot`` is the product path: the forward is the index-mode kernel pipeline (spn_sample_sum / spn_sample_leaf:
the one-hots of the reference are never materialised, only the chosen indices), the backward is one
spn_sample_leaf_bwd + one spn_sample_sum_bwd per Sum layer, bottom-up, which re-create the Gumbel-softmax
probabilities from the same noise and push dL/d(one-hot) up the selected path.  Gradients reach the sum weights, the
leaf parameters and the evidence activations cached by ``provide_evidence`` (and through them a CSPN's gating network).
"""
from typing import List, Optional

import torch as th

import numpy as np

from . import ops
from .layers import CrossProduct, Sum
from .utils import Sample


class _Plan:
    """Host-side description of one top-down pass (everything that is not a differentiable tensor)."""
    __slots__ = ("steps", "Q", "Qe", "Rs", "R", "pa", "gumbel", "eps", "seed", "mpe", "card", "inv32", "perm32", "ev",
                 "ev_strides", "squash", "ch", "rep")


class DiffSampleFn(th.autograd.Function):
    """y = sample(plan; mu, sigma, lw_0..lw_{k-1}, xc_0..xc_{k-1}), Sum layers listed top -> bottom."""

    @staticmethod
    def forward(ctx, plan: _Plan, mu, sigma, *rest):
        k = len(plan.steps)
        lws = [ops._f32c(t.detach()) for t in rest[:k]]
        xcs = [None if t is None else ops._f32c(t.detach()) for t in rest[k:]]
        mu, sigma = ops._f32c(mu.detach()), ops._f32c(sigma.detach())
        Q, Qe, Rs = plan.Q, plan.Qe, plan.Rs
        parent, pstr, rep = plan.pa, (1, 0, 0), None
        rec = []
        for j, (is_root, d, unravel) in enumerate(plan.steps):
            g = plan.gumbel[j]
            if is_root:
                ic_top = lws[j].shape[2] // plan.R
                k0 = ops.sample_sum(lws[j], plan.R, Q, 1, 1, parent, pstr, False, None, xcs[j], Qe, g, plan.mpe,
                                    plan.seed).view(Q)
                rep = th.div(k0, ic_top, rounding_mode='floor').to(th.int32)
                rec.append((parent, pstr, False, rep, g))
                parent, pstr = (k0 - rep * ic_top).to(th.int32).view(Q, 1, 1), (1, 1, 1)
            else:
                if g is not None and rep is not None and g.shape[-1] != 1:  # supplied noise covers every repetition
                    ic = lws[j].shape[2]
                    g = g.reshape(Q, d, ic, -1).gather(3, rep.long().view(Q, 1, 1, 1).expand(Q, d, ic, 1)).contiguous()
                out = ops.sample_sum(lws[j], 0, Q, d, Rs, parent, pstr, unravel, rep, xcs[j], Qe, g, plan.mpe, plan.seed)
                rec.append((parent, pstr, unravel, rep, g))
                parent, pstr = out, (d * Rs, Rs, 1)
        y, ch = ops.sample_leaf(mu, sigma, Q, plan.card, Rs, parent, pstr, rep, plan.eps, plan.mpe, plan.inv32, plan.ev,
                                plan.ev_strides, Qe, plan.squash, True, plan.seed)
        plan.ch, plan.rep = ch, rep
        ctx.plan, ctx.rec, ctx.lws, ctx.xcs, ctx.leaf = plan, rec, lws, xcs, (mu, sigma)
        # the output must go through save_for_backward: a plain attribute would close a reference cycle (node -> y -> node)
        # that keeps the whole autograd graph (and its AccumulateGrad nodes) alive until the cyclic GC runs
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        plan, rec, lws, xcs = ctx.plan, ctx.rec, ctx.lws, ctx.xcs
        mu, sigma = ctx.leaf
        (y,) = ctx.saved_tensors
        k = len(plan.steps)
        Q, Qe, Rs = plan.Q, plan.Qe, plan.Rs
        through_sums = k > 0 and not plan.mpe
        allreps = k == 1 and plan.steps[0][0]  # D = 1: nothing below the root clears the unselected repetitions
        dmu, dsigma, g0 = ops.sample_leaf_backward(mu, sigma, Q, plan.card, Rs, plan.ch, plan.rep, plan.eps, plan.mpe,
                                                   plan.seed, plan.perm32, plan.ev, plan.ev_strides, Qe, plan.squash, y,
                                                   ops._f32c(gy), through_sums, allreps)
        dlws, dxcs = [None] * k, [None] * k
        if through_sums:
            gc, gc_d = g0, g0.shape[1]
            for j in range(k - 1, -1, -1):
                is_root, d, _ = plan.steps[j]
                parent, pstr, unravel, rep, g = rec[j]
                dlws[j] = th.zeros_like(lws[j])
                dxcs[j] = None if xcs[j] is None else th.zeros_like(xcs[j])
                gp = ops.sample_sum_backward(lws[j], plan.R if is_root else 0, Q, d, 1 if is_root else Rs, parent, pstr,
                                             unravel, rep, xcs[j], Qe, g, plan.seed, gc, gc_d, dlws[j], dxcs[j], j > 0,
                                             allreps)
                gc, gc_d = gp, d
        return (None, dmu, None if plan.mpe else dsigma, *dlws, *dxcs)


def sample_onehot(spn, ctx: Sample, class_index, evidence, layer_index: int, noise: Optional[List[th.Tensor]],
                  fuse_post: bool = False):
    """Differentiable sampling on the fused kernels; returns (ctx, post-processing already applied?)."""
    cfg = spn.config
    dev = spn.device
    n = tuple(ctx.n)
    if layer_index == 0:
        return spn._sample_leaf_root(ctx, _num_conditionals(spn, class_index, evidence), noise), False
    w = _num_conditionals(spn, class_index, evidence)
    plan = _Plan()
    plan.mpe, plan.R, plan.card = bool(ctx.is_mpe), cfg.R, spn._leaf.cardinality
    plan.Qe = int(np.prod(evidence.shape[:-3])) if evidence is not None else 1
    plan.seed = ops.new_seed(dev, plan.mpe or noise is not None)

    def take(shape):
        if plan.mpe or noise is None:
            return None
        t = noise.pop(0)
        assert tuple(t.shape) == tuple(shape), f"noise shape {tuple(t.shape)} != {tuple(shape)}"
        return t.to(device=dev, dtype=th.float32).contiguous()

    steps, layers, gumbel = [], [], []
    start = layer_index
    if layer_index == spn.max_layer_index:
        ctx.scopes = 1
        K = spn.root.weights.shape[2]
        if class_index is not None:
            nodes = 1
            node = th.as_tensor(class_index, device=dev).to(th.int32).view(*([1] * (len(n) + 1)), w)
        else:
            nodes = cfg.C
            node = th.arange(nodes, device=dev, dtype=th.int32).view(nodes, *([1] * (len(n) + 1)))
        lead = (nodes, *n, w)
        steps.append((True, 1, False))
        layers.append(spn.root)
        gumbel.append(take((*lead, 1, K, 1)))
        plan.Rs = 1
        start = layer_index - 1
    else:
        layer = spn.layer_index_to_obj(layer_index)
        nodes = layer.in_channels ** 2 if isinstance(layer, CrossProduct) else layer.out_channels
        node = th.arange(nodes, device=dev, dtype=th.int32).view(nodes, *([1] * (len(n) + 1)))
        lead = (nodes, *n, w)
        plan.Rs = cfg.R
        if isinstance(layer, CrossProduct):
            ctx.scopes = layer.in_features // layer.cardinality
        else:
            ctx.scopes = layer.in_features
            steps.append((False, layer.in_features, False))
            layers.append(layer)
            gumbel.append(take((*lead, layer.in_features, layer.in_channels, cfg.R)))
        start = layer_index - 1
    for li in range(start, 0, -1):
        if li % 2 == 1:
            continue
        layer = spn._inner_layers[li - 1]
        steps.append((False, layer.in_features, True))
        layers.append(layer)
        gumbel.append(take((*lead, layer.in_features, layer.in_channels, cfg.R)))
    is_root = layer_index == spn.max_layer_index
    plan.eps = take((*lead, cfg.F) if is_root else (*lead, cfg.F, cfg.R))
    plan.steps, plan.gumbel = steps, gumbel
    plan.Q = int(np.prod(lead))
    plan.pa = node.expand(*lead).reshape(plan.Q).contiguous()
    perm32, inv32 = spn._perm32()
    plan.inv32, plan.perm32 = (inv32, perm32) if fuse_post else (None, None)
    plan.ev, plan.ev_strides = None, (0, 0, 0)
    if fuse_post and evidence is not None:
        assert nodes == 1, "sampling with evidence needs a single root node (C = 1 or class_index)"
        plan.ev = evidence.to(th.float32).contiguous()
        r_ev = plan.ev.shape[-1]
        plan.ev_strides = (cfg.F * plan.ev.shape[-2] * r_ev, plan.ev.shape[-2] * r_ev, 1 if r_ev > 1 else 0)
    plan.squash = bool(fuse_post and cfg.tanh_squash and ctx.needs_squashing)

    base = spn._leaf.base_leaf
    xcs = [l._input_cache_fused if l._is_input_cache_enabled else None for l in layers]
    y = DiffSampleFn.apply(plan, base.means, base.stds, *[l.weights for l in layers], *xcs)

    Rs, F_, I_ = plan.Rs, cfg.F, base.means.shape[2]
    ch = plan.ch.view(plan.Q, F_, 1, Rs).long()
    if is_root:  # reference layout [.., F, I, R]: all-zero one-hots in the repetitions that were not selected
        oh = th.zeros(plan.Q, F_, I_, cfg.R, device=dev)
        oh.view(plan.Q, F_, I_ * cfg.R).scatter_(2, ch.view(plan.Q, F_, 1) * cfg.R + plan.rep.long().view(plan.Q, 1, 1), 1.0)
        ctx.parent_indices = oh.view(*lead, F_, I_, cfg.R)
        ctx.onehot_repetition = plan.rep.view(*lead).long()
    else:
        ctx.parent_indices = th.zeros(plan.Q, F_, I_, Rs, device=dev).scatter_(2, ch, 1.0).view(*lead, F_, I_, Rs)
    ctx.has_rep_dim = not is_root
    if fuse_post:
        ctx.sample = y.view(*lead, F_, Rs)
        ctx.permutation_inverted = True
        ctx.has_rep_dim = True
        ctx.evidence_filled_in = evidence is not None
        if plan.squash:
            ctx.needs_squashing = False
    else:
        ctx.sample = y.view(*lead, F_) if is_root else y.view(*lead, F_, Rs)
    return ctx, fuse_post


def _num_conditionals(spn, class_index, evidence) -> int:
    if evidence is not None:
        return evidence.shape[-4]
    if class_index is not None:
        return class_index.shape[0]
    return spn._num_weight_sets()
