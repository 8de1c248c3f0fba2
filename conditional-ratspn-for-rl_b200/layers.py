"""Sum / Product / CrossProduct layers with the constructor signatures, attributes, tensor layouts and
``state_dict`` keys of the reference's ``layers.py`` -- but every arithmetic step is a call into libspn_b200.so.

Activation layout (same as the reference, layers.py:95-96):  [*batch_dims, w (conditionals), d (scopes), c (channels),
r (repetitions)], fp32.  Sum log-weights: [W, d, ic, oc, r] with W = 1 (RatSpn) or W = w (CSPN).

``RatSpn.forward`` does not chain these modules one by one: it fuses every CrossProduct into the Sum above it
(ops.sum_forward(mode='xsum')).  The stand-alone ``forward`` methods below exist because the reference API calls
layers individually as well (rat_spn.py:673, :707).
"""
import logging

import numpy as np
import torch as th
from torch import nn
from torch.nn import functional as F

from . import ops
from .type_checks import check_valid
from .utils import Sample

logger = logging.getLogger(__name__)


class AbstractLayer(nn.Module):
    def __init__(self, in_features: int, num_repetitions: int = 1):
        super().__init__()
        self.in_features = check_valid(in_features, int, 1)
        self.num_repetitions = check_valid(num_repetitions, int, 1)


class Sum(AbstractLayer):
    """Sum nodes: out[.., s, o, r] = logsumexp_i(x[.., s, i, r] + log w[W, s, i, o, r])   (layers.py:22-128)."""

    def __init__(self, in_channels: int, in_features: int, out_channels: int, ratspn: bool, num_repetitions: int,
                 dropout: float = 0.0):
        super().__init__(in_features, num_repetitions)
        self.in_channels = check_valid(in_channels, int, 1)
        self.out_channels = check_valid(out_channels, int, 1)
        self.dropout = nn.Parameter(th.tensor(check_valid(dropout, float, 0.0, 1.0)), requires_grad=False)
        self._dropout_p = float(dropout)  # host copy: no device sync in forward
        self.ratspn = ratspn
        # one weight set (w = 1); a CSPN swaps this for per-conditional tensors in set_params()
        self.weight_param = nn.Parameter(
            th.randn(1, self.in_features, self.in_channels, self.out_channels, self.num_repetitions))
        self.out_shape = f"(N, {self.in_features}, {self.out_channels}, {self.num_repetitions})"
        self._is_input_cache_enabled = False
        self._input_cache = None        # stand-alone forward: the layer input itself (layers.py:102-103)
        self._input_cache_fused = None  # fused forward: the *CrossProduct input* this layer's input derives from
        self.consolidated_weights = None
        self._raw_weights = False       # CSPN with fused weight normalisation: weight_param holds the raw head outputs
        self._tc_cache = {}             # bf16 operand image + log-normaliser of weight_param (tensor-core path)

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_tc_cache"] = {}  # derived device buffers are not part of a checkpoint
        return state

    def __setstate__(self, state):
        """Also accepts a pickle written by the reference implementation (th.save(model), rat_spn.py:1046-1047): its
        Sum layers carry none of this implementation's host-side attributes, which get their defaults here."""
        super().__setstate__(state)
        d = self.__dict__
        d.setdefault("_dropout_p", float(self.dropout.detach().cpu()) if "dropout" in self._parameters else 0.0)
        d.setdefault("_input_cache_fused", None)
        d.setdefault("_raw_weights", False)
        d["_tc_cache"] = {}

    def invalidate_weight_cache(self):
        """Drop the cached operand image (needed only after writing to ``weight_param.data`` behind autograd's back;
        in-place updates of the Parameter itself are detected through its version counter)."""
        self._tc_cache.clear()

    def _enable_input_cache(self):
        self._is_input_cache_enabled = True

    def _disable_input_cache(self):
        self._is_input_cache_enabled = False
        self._input_cache = None
        self._input_cache_fused = None

    @property
    def weights(self) -> th.Tensor:
        """Log-weights [w, d, ic, oc, r]; normalised over ic on every read for a RatSpn (layers.py:81-88)."""
        if self.ratspn or self._raw_weights:
            return F.log_softmax(self.weight_param, dim=2)
        return self.weight_param

    def _weights_for(self, detach_params: bool) -> th.Tensor:
        lw = self.weights
        return lw.detach() if detach_params else lw

    def forward(self, x: th.Tensor, detach_params: bool = False) -> th.Tensor:
        """x [*batch, w, d, ic, r] -> [*batch, w, d, oc, r]."""
        if self._is_input_cache_enabled:
            self._input_cache = x.detach().clone()
        if self._dropout_p > 0.0 and self.training:
            # the reference overwrites dropped children with -inf in place (layers.py:105-108)
            x = x.masked_fill(th.rand_like(x) < self._dropout_p, float("-inf"))
        return ops.sum_forward(x, self._weights_for(detach_params), "sum")

    def sample(self, mode: str = None, ctx: Sample = None) -> Sample:
        """One top-down step in index mode: pick an in-channel per (scope, repetition) given the parent's choice
        (layers.py:130-237).  Differentiable ('onehot') sampling is provided at model level (RatSpn.sample)."""
        assert mode is not None, "A sampling mode must be provided!"
        if mode != "index":
            raise NotImplementedError("layer-wise sampling supports mode='index'; use RatSpn.sample for 'onehot'")
        lw = self.weights
        W, d, ic, oc, R = lw.shape
        dev = lw.device
        n = tuple(ctx.n)
        if ctx.is_root:
            w = W
            lead = (oc, *n, w)
            Q = int(np.prod(lead))
            node = th.arange(oc, device=dev, dtype=th.int32).view(oc, *([1] * (len(n) + 1))).expand(*lead)
            pa = node.reshape(Q, 1, 1).expand(Q, d, R).contiguous()
            rep, Rs = None, R
        else:
            pa_full = ctx.parent_indices
            lead = tuple(pa_full.shape[:-2])
            Q = int(np.prod(lead))
            Rs = pa_full.shape[-1]
            pa = pa_full.reshape(Q, d, Rs).to(th.int32).contiguous()
            rep = ctx.repetition_indices.reshape(Q).to(th.int32).contiguous() if ctx.repetition_indices is not None else None
        xc, Qe = None, 1
        if self._is_input_cache_enabled and self._input_cache is not None:
            raise NotImplementedError("layer-wise evidence sampling: use RatSpn.sample(evidence=...)")
        gumbel = None
        if not ctx.is_mpe:
            gumbel = -th.empty(Q, d, ic, Rs, device=dev).exponential_().log()
        out = ops.sample_sum(lw, 0, Q, d, Rs, pa, (d * Rs, Rs, 1), False, rep, xc, Qe, gumbel, ctx.is_mpe)
        ctx.parent_indices = out.view(*lead, d, Rs).long()
        return ctx

    def sample_index_style(self, ctx: Sample = None) -> Sample:
        return self.sample(mode='index', ctx=ctx)

    def sample_onehot_style(self, ctx: Sample = None) -> Sample:
        return self.sample(mode='onehot', ctx=ctx)

    def __repr__(self):
        return (f"Sum(in_channels={self.in_channels}, in_features={self.in_features}, "
                f"out_channels={self.out_channels}, dropout={self._dropout_p}, out_shape={self.out_shape})")


class Product(AbstractLayer):
    """Product over `cardinality` consecutive features (layers.py:251-367).  Inside the leaf this reduction is done
    by the leaf kernel; the module form is kept for the sampling fan-out and for API compatibility."""

    def __init__(self, in_features: int, cardinality: int, num_repetitions: int = 1):
        super().__init__(in_features, num_repetitions)
        self.cardinality = check_valid(cardinality, int, 1, in_features + 1)
        self._pad = (self.cardinality - self.in_features % self.cardinality) % self.cardinality
        self._out_features = int(np.ceil(self.in_features / self.cardinality))
        self.out_shape = f"(N, {self._out_features}, in_channels, {self.num_repetitions})"

    def forward(self, x: th.Tensor, reduction='sum', **kwargs):
        """x [*batch, w, F, c, r] -> [*batch, w, ceil(F / card), c, r].  Pure data movement + a short add chain on
        already materialised leaf outputs; only reached through direct calls (the model path is fused)."""
        if self.cardinality == x.shape[2]:
            return x.sum(2, keepdim=True) if reduction == 'sum' else x
        if self.cardinality == 1:
            return x
        if self._pad > 0:
            x = F.pad(x, pad=(0, 0, 0, 0, 0, self._pad), value=0)
        *n, w, d, c, r = x.shape
        x = x.view(*n, w, d // self.cardinality, self.cardinality, c, r)
        if reduction is None:
            return x
        if reduction == 'sum':
            return x.sum(dim=-3)
        raise NotImplementedError("No reduction other than sum is implemented. ")

    def repeat_by_cardinality(self, x, feature_dim=-2):
        x = th.repeat_interleave(x, repeats=self.cardinality, dim=feature_dim)
        if self._pad:
            x, _ = th.tensor_split(x, [-self._pad], dim=feature_dim)
        return x

    def sample(self, n: int = None, ctx: Sample = None) -> Sample:
        if ctx.is_root:
            if self.num_repetitions != 1:
                raise Exception("Cannot start sampling from Product layer with num_repetitions > 1 and no context given.")
            ctx.parent_indices = th.zeros(*ctx.n, 1, dtype=int)
            ctx.repetition_indices = th.zeros(*ctx.n, dtype=int)
            return ctx
        ctx.parent_indices = self.repeat_by_cardinality(ctx.parent_indices, -2 if ctx.sampling_mode == 'index' else -3)
        return ctx

    sample_index_style = sample_onehot_style = lambda self, ctx=None: self.sample(ctx=ctx)

    def __repr__(self):
        return f"Product(in_features={self.in_features}, cardinality={self.cardinality}, out_shape={self.out_shape})"


class CrossProduct(AbstractLayer):
    """Pairs scopes (2j, 2j+1) and forms all c*c channel combinations: out[l*c + r'] = left[l] + right[r']
    (layers.py:370-563).  Scope counts that are not a power of two are zero padded."""

    def __init__(self, in_features: int, in_channels: int, num_repetitions: int = 1):
        padded = 1 << max(0, int(in_features - 1).bit_length())
        self._pad = padded - in_features
        super().__init__(int(padded), num_repetitions)
        self.in_channels = check_valid(in_channels, int, 1)
        self.cardinality = check_valid(2, int, 2, in_features + 1)
        self._out_features = int(np.ceil(self.in_features / self.cardinality))
        pairs = th.stack(th.meshgrid(th.arange(self.in_channels), th.arange(self.in_channels), indexing="ij"), dim=-1)
        self.unraveled_channel_indices = nn.Parameter(pairs.reshape(-1, 2), requires_grad=False)
        self.one_hot_in_channel_mapping = nn.Parameter(
            F.one_hot(self.unraveled_channel_indices, num_classes=self.in_channels).float(), requires_grad=False)
        self.num_conditionals = 1  # set by CSPN.set_params; only needed when sampling starts at this layer
        self.out_shape = f"(N, {self._out_features}, {self.in_channels ** 2}, {self.num_repetitions})"

    def forward(self, x: th.Tensor, **kwargs) -> th.Tensor:
        """x [*batch, w, d, c, r] -> [*batch, w, pow2ceil(d)/2, c*c, r] (materialising kernel)."""
        out = ops.cross_product(x)
        assert out.shape[-3:] == (self.in_features // self.cardinality, self.in_channels ** 2, self.num_repetitions)
        return out

    def split_shuffled_scopes(self, t: th.Tensor, scope_dim: int):
        """[x_0L, x_0R, x_1L, x_1R, ...] along ``scope_dim`` -> (left scopes, right scopes)   (layers.py:545-563)."""
        scope_dim = scope_dim % t.dim()
        if self._pad > 0:
            shape = list(t.shape)
            shape[scope_dim] = self._pad
            t = th.cat((t, t.new_zeros(shape)), dim=scope_dim)
        idx = th.arange(t.size(scope_dim), device=t.device)
        return t.index_select(scope_dim, idx[0::2]), t.index_select(scope_dim, idx[1::2])

    def sample(self, mode: str = None, ctx: Sample = None) -> Sample:
        """Index mode: map each chosen pair channel back to its (left, right) child channels (layers.py:442-511)."""
        assert mode is not None, "A sampling mode must be provided!"
        if mode != "index":
            raise NotImplementedError("layer-wise sampling supports mode='index'; use RatSpn.sample for 'onehot'")
        c = self.in_channels
        if ctx.is_root:
            dev = self.unraveled_channel_indices.device
            k = th.arange(c * c, device=dev).view(c * c, *([1] * (len(ctx.n) + 3)))
            parent = k.expand(c * c, *ctx.n, self.num_conditionals, self.in_features // 2, self.num_repetitions)
        else:
            parent = ctx.parent_indices
        pair = th.stack((th.div(parent, c, rounding_mode="floor"), parent % c), dim=-2)  # [.., d, 2, r]
        idx = pair.reshape(*parent.shape[:-2], parent.shape[-2] * 2, parent.shape[-1])
        if self._pad:
            idx = idx[..., :-self._pad, :]
        ctx.parent_indices = idx
        return ctx

    def sample_index_style(self, ctx: Sample = None) -> Sample:
        return self.sample(mode='index', ctx=ctx)

    def sample_onehot_style(self, ctx: Sample = None) -> Sample:
        return self.sample(mode='onehot', ctx=ctx)

    def consolidate_weights(self, parent_weights):
        """Marginal weight a parent sum puts on each left / right child channel (layers.py:519-540)."""
        n, d, p_ic, p_oc, r = parent_weights.shape
        c = self.in_channels
        assert p_ic == c ** 2 and d * 2 == self.in_features
        pw = parent_weights.softmax(dim=2).view(n, d, c, c, p_oc, r)
        both = th.stack((pw.sum(dim=3), pw.sum(dim=2)), dim=2)
        return both.view(n, self.in_features, c, p_oc, r)

    def __repr__(self):
        return f"CrossProduct(in_features={self.in_features}, out_shape={self.out_shape})"
