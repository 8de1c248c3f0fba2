"""Leaf layers with the API of the reference's ``distributions.py``: ``Leaf``, ``RatNormal``,
``IndependentMultivariate`` and ``truncated_normal_``.  The density evaluation (and the product over the leaf
cardinality, and the per-repetition feature permutation when the model asks for it) is one CUDA kernel
(ops.leaf_forward -> spn_leaf_fwd).
"""
import logging
import math
from typing import Dict

import numpy as np
import torch as th
from torch import nn
from torch.nn import functional as F

from . import ops
from .layers import AbstractLayer, Product
from .type_checks import check_valid
from .utils import Sample

logger = logging.getLogger(__name__)


class Leaf(AbstractLayer):
    """Base class of input distributions; NaN inputs are marginalised (log-density 0)   (distributions.py:18-64)."""

    def __init__(self, in_features: int, out_channels: int, num_repetitions: int = 1, dropout=0.0):
        super().__init__(in_features=in_features, num_repetitions=num_repetitions)
        self.out_channels = check_valid(out_channels, int, 1)
        dropout = check_valid(dropout, float, 0.0, 1.0)
        self.dropout = nn.Parameter(th.tensor(dropout), requires_grad=False)
        self._dropout_p = float(dropout)
        self.out_shape = f"(N, {in_features}, {out_channels})"
        self.marginalization_constant = nn.Parameter(th.zeros(1), requires_grad=False)

    def __setstate__(self, state):
        super().__setstate__(state)   # a pickle written by the reference lacks the host copy of the dropout probability
        self.__dict__.setdefault("_dropout_p", float(self.dropout.detach().cpu()) if "dropout" in self._parameters else 0.0)

    def _apply_dropout(self, x: th.Tensor) -> th.Tensor:
        if self._dropout_p > 0.0 and self.training:
            x = x.masked_fill(th.rand_like(x) < self._dropout_p, 0.0)
        return x

    def __repr__(self):
        return (f"{self.__class__.__name__}(in_features={self.in_features}, out_channels={self.out_channels}, "
                f"dropout={self._dropout_p}, out_shape={self.out_shape})")


class RatNormal(Leaf):
    """Gaussian leaves with sigmoid-bounded std (and optionally mean)   (distributions.py:67-308)."""

    def __init__(self, in_features: int, out_channels: int, ratspn: bool, num_repetitions: int, dropout: float = 0.0,
                 tanh_squash: bool = False, min_sigma: float = 0.0, max_sigma: float = 1.0, min_mean: float = None,
                 max_mean: float = None, no_tanh_log_prob_correction: bool = False, stds_in_lin_space: bool = True,
                 stds_sigmoid_bound: bool = True):
        super().__init__(in_features, out_channels, num_repetitions, dropout)
        shape = (1, in_features, out_channels, num_repetitions)
        self.mean_param = nn.Parameter(th.randn(*shape))
        self._tanh_squash = tanh_squash
        self._no_tanh_log_prob_correction = no_tanh_log_prob_correction
        self._stds_in_lin_space = stds_in_lin_space
        self._stds_sigmoid_bound = stds_sigmoid_bound
        self._ratspn = ratspn
        if min_sigma is not None and max_sigma is not None:
            self.std_param = nn.Parameter(th.randn(*shape))
        else:
            self.std_param = nn.Parameter(th.rand(*shape))
        self.min_sigma = check_valid(min_sigma, float, 0.0, max_sigma, allow_none=False)
        self.max_sigma = check_valid(max_sigma, float, min_sigma, allow_none=False)
        self.min_log_sigma = math.log(self.min_sigma + 1e-8)
        self.max_log_sigma = math.log(self.max_sigma + 1e-8)
        if self._tanh_squash:
            assert min_mean is None and max_mean is None, \
                "When leaves are tanh-squashed, the min_mean and max_mean values are predefined and cannot be set."
            min_mean, max_mean = -6.0, 6.0
        self.min_mean = check_valid(min_mean, float, upper_bound=max_mean, allow_none=True)
        self.max_mean = check_valid(max_mean, float, min_mean, allow_none=True)
        self._means_sigmoid_bound = True
        if not (stds_in_lin_space and stds_sigmoid_bound):
            # the experiments only use the linear-space sigmoid bound; the kernels take linear-space sigma
            logger.warning("RatNormal: non-default std parameterisation is converted to linear space on every call")

    @property
    def device(self):
        return self.marginalization_constant.device

    @property
    def tanh_correction_enabled(self) -> bool:
        return bool(self._tanh_squash and not self._no_tanh_log_prob_correction)

    def bounded_means(self, means: th.Tensor = None):
        means = self.mean_param if means is None else means
        if self.min_mean is not None or self.max_mean is not None:
            if self._means_sigmoid_bound:
                return self.min_mean + (self.max_mean - self.min_mean) * th.sigmoid(means)
            return th.clamp(means, self.min_mean, self.max_mean)
        return means

    def bounded_stds(self, stds: th.Tensor = None):
        stds = self.std_param if stds is None else stds
        if self._stds_in_lin_space:
            if self._stds_sigmoid_bound:
                return self.min_sigma + (self.max_sigma - self.min_sigma) * th.sigmoid(stds)
            return F.softplus(stds) + self.min_sigma
        return F.softplus(stds) + self.min_log_sigma

    @property
    def means(self):
        # RatSpn: bounded on every read; CSPN: set_params stored already-bounded tensors (distributions.py:156-162)
        return self.bounded_means() if self._ratspn else self.mean_param

    @means.setter
    def means(self, means: th.Tensor):
        if isinstance(means, np.ndarray):
            means = th.as_tensor(means)
        means = self.bounded_means(means)
        self.mean_param = nn.Parameter(th.as_tensor(means, dtype=th.float, device=self.device))

    @property
    def stds(self):
        std_param = self.bounded_stds() if self._ratspn else self.std_param
        return std_param if self._stds_in_lin_space else std_param.exp()

    @property
    def log_stds(self):
        std_param = self.bounded_stds() if self._ratspn else self.std_param
        return std_param.log() if self._stds_in_lin_space else std_param

    @property
    def var(self):
        return self.stds ** 2

    def _params_for(self, detach_params: bool):
        mu, sigma = self.means, self.stds
        if detach_params:
            mu, sigma = mu.detach(), sigma.detach()
        return mu, sigma

    def forward(self, x, detach_params: bool = False):
        """x [*batch, w, F, 1|I, 1|R] (pre-squash values) -> log-density [*batch, w, F, I, R]   (:199-246)."""
        mu, sigma = self._params_for(detach_params)
        out = ops.leaf_forward(x, mu, sigma, None, 1, self.tanh_correction_enabled)
        return self._apply_dropout(out)

    def sample(self, mode: str = None, ctx: Sample = None):
        """Index mode: gather (mean, std) of the chosen channel / repetition and draw (distributions.py:248-302)."""
        if mode == "onehot":
            raise NotImplementedError("layer-wise sampling supports mode='index'; use RatSpn.sample for 'onehot'")
        means, stds = self.means, self.stds
        if ctx.is_root:
            mu = means.expand(*ctx.n, *means.shape)
            sg = stds.expand(*ctx.n, *stds.shape)
        else:
            pi = ctx.parent_indices
            mu = means.expand(*pi.shape[:-1], -1, -1)
            sg = stds.expand(*pi.shape[:-1], -1, -1)
            if ctx.repetition_indices is not None:
                ri = ctx.repetition_indices.view(*ctx.repetition_indices.shape, 1, 1, 1).expand(*mu.shape[:-1], 1)
                mu, sg = th.gather(mu, -1, ri), th.gather(sg, -1, ri)
            mu = th.gather(mu, -2, pi.unsqueeze(-2)).squeeze(-2)
            sg = th.gather(sg, -2, pi.unsqueeze(-2)).squeeze(-2)
            if ctx.repetition_indices is not None:
                mu, sg = mu.squeeze(-1), sg.squeeze(-1)
        return mu if ctx.is_mpe else mu + sg * th.randn_like(mu)

    def sample_index_style(self, ctx: Sample = None) -> th.Tensor:
        return self.sample(mode='index', ctx=ctx)

    def sample_onehot_style(self, ctx: Sample = None) -> th.Tensor:
        return self.sample(mode='onehot', ctx=ctx)


class IndependentMultivariate(Leaf):
    """Leaf distributions over `cardinality` consecutive features with diagonal covariance: a base leaf followed by a
    product (distributions.py:311-402).  Here both run in one kernel."""

    def __init__(self, in_features: int, out_channels: int, cardinality: int, ratspn: bool, num_repetitions: int,
                 dropout: float = 0.0, tanh_squash: bool = False, leaf_base_class: Leaf = RatNormal,
                 leaf_base_kwargs: Dict = None):
        super().__init__(in_features, out_channels, num_repetitions, dropout)
        self.base_leaf = leaf_base_class(out_channels=out_channels, in_features=in_features, dropout=dropout,
                                         num_repetitions=num_repetitions, tanh_squash=tanh_squash, ratspn=ratspn,
                                         **(leaf_base_kwargs or {}))
        self._pad = (cardinality - self.in_features % cardinality) % cardinality
        self.prod = Product(in_features=in_features + self._pad, cardinality=cardinality,
                            num_repetitions=num_repetitions)
        self.cardinality = check_valid(cardinality, int, 1, in_features + 1)
        self.out_shape = f"(N, {self.prod._out_features}, {out_channels}, {self.num_repetitions})"

    @property
    def out_features(self):
        return self.prod._out_features

    @property
    def pad(self):
        return self._pad

    def _init_weights(self):
        # kept for API parity; like the reference it acts on a temporary and leaves the parameters untouched
        # (rat_spn.py:320-329, distributions.py:362-364)
        if isinstance(self.base_leaf, RatNormal):
            truncated_normal_(self.base_leaf.stds, std=0.5)

    def pad_input(self, x: th.Tensor):
        if self._pad:
            x = F.pad(x, pad=[0, 0, 0, 0, 0, self._pad], mode="constant", value=0.0)
        return x

    def forward(self, x: th.Tensor, reduction='sum', detach_params: bool = False, perm32: th.Tensor = None,
                layout: str = "ref"):
        """x [*batch, w, F, 1|I, 1|R] -> [*batch, w, ceil(F/card), I, R].  ``perm32`` (int32 [F, R]) applies the
        model's per-repetition feature permutation inside the kernel."""
        base = self.base_leaf
        if reduction != 'sum' or base._dropout_p > 0.0 and self.training or not isinstance(base, RatNormal):
            if perm32 is not None:
                raise NotImplementedError("permutation fusion requires the default (sum) reduction")
            return self.prod(self.pad_input(base(x, detach_params=detach_params)), reduction=reduction)
        mu, sigma = base._params_for(detach_params)
        return ops.leaf_forward(x, mu, sigma, perm32, self.cardinality, base.tanh_correction_enabled, layout)

    def sample(self, mode: str = None, ctx: Sample = None):
        if not ctx.is_root:
            ctx = self.prod.sample(ctx=ctx)
            if self._pad:  # strip the padded feature slots (the reference slices the wrong dim in onehot mode)
                if ctx.sampling_mode == 'index' or mode == 'index':
                    ctx.parent_indices = ctx.parent_indices[..., :-self._pad, :]
                else:
                    ctx.parent_indices = ctx.parent_indices[..., :-self._pad, :, :]
        return self.base_leaf.sample(ctx=ctx, mode=mode)

    def sample_index_style(self, ctx: Sample = None) -> th.Tensor:
        return self.sample(ctx=ctx, mode='index')

    def sample_onehot_style(self, ctx: Sample = None) -> th.Tensor:
        return self.sample(ctx=ctx, mode='onehot')

    def __repr__(self):
        return (f"IndependentMultivariate(in_features={self.in_features}, out_channels={self.out_channels}, "
                f"dropout={self._dropout_p}, cardinality={self.cardinality}, out_shape={self.out_shape})")


def truncated_normal_(tensor, mean=0, std=0.1):
    """Fill ``tensor`` in place with N(mean, std) samples truncated at two standard deviations."""
    with th.no_grad():
        draws = th.randn(*tensor.shape, 4, device=tensor.device, dtype=tensor.dtype)
        ok = draws.abs() < 2
        first = ok.float().argmax(dim=-1, keepdim=True)
        tensor.copy_(draws.gather(-1, first).squeeze(-1) * std + mean)
    return tensor
