"""Conditional SPN with the public API of the reference's ``cspn.py``: a gating MLP maps each conditional to its own set
of sum weights and leaf parameters, which the RAT-SPN kernels then consume as per-sample (W = w) weights.

The gating network itself is a stack of ``nn.Linear`` layers (cuBLAS GEMMs, as in the reference and as scoped by the
north star); everything downstream of ``set_params`` -- log-likelihood, sampling, entropy -- runs on the per-sample
weight kernels of libspn_b200.so, where every weight is read exactly once (HBM-streaming).
"""
import logging
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Optional, Tuple, Type

import numpy as np
import torch as th
import torch.nn.functional as F
from torch import nn

from .layers import CrossProduct, Sum
from .rat_spn import RatSpn, RatSpnConfig

logger = logging.getLogger(__name__)


def print_cspn_params(cspn):
    def count_params(model):
        return sum(p.numel() for p in model.parameters() if p.requires_grad)
    print(f"Total params in CSPN: {count_params(cspn)}")
    print(f"Params to extract features from the conditional: {count_params(cspn.feat_layers)}")
    print(f"Params in MLP for the sum params, excluding the heads: {count_params(cspn.sum_layers)}")
    print(f"Params in the heads of the sum param MLPs: {sum(count_params(h) for h in cspn.sum_param_heads)}")
    print(f"Params in MLP for the dist params, excluding the heads: {count_params(cspn.dist_layers)}")
    print(f"Params in the heads of the dist param MLPs: "
          f"{count_params(cspn.dist_mean_head) + count_params(cspn.dist_std_head)}")


@dataclass
class CspnConfig(RatSpnConfig):
    is_ratspn: bool = False
    F_cond: tuple = 0
    feat_layers: list = None
    conv_kernel_size: int = 5
    conv_pooling_kernel_size: int = 3
    conv_pooling_stride: int = 3
    sum_param_layers: list = None
    dist_param_layers: list = None
    cond_layers_inner_act: Type[nn.Module] = nn.LeakyReLU
    # extension: keep the inner Sum layers' head outputs UNNORMALISED and fold F.log_softmax into the
    # streaming kernels; `layer.weight_param` is then raw (as for a RatSpn), `layer.weights` normalises on read
    fuse_weight_norm: bool = False

    def __setattr__(self, key, value):
        if hasattr(self, key):
            super().__setattr__(key, value)
        else:
            raise AttributeError(f"CspnConfig object has no attribute {key}")


def _mlp(sizes, inner_act, last_is_identity: bool) -> nn.Sequential:
    """Linear/activation stack; with ``last_is_identity`` the final Linear is followed by nn.Identity (so module
    indices, and therefore state_dict keys, match cspn.py:216-244)."""
    if len(sizes) < 2:
        return nn.Sequential(nn.Identity())
    mods = []
    for j in range(len(sizes) - 1):
        act = nn.Identity if (last_is_identity and j == len(sizes) - 2) else inner_act
        mods += [nn.Linear(sizes[j], sizes[j + 1]), act()]
    return nn.Sequential(*mods)


class CSPN(RatSpn):
    def __init__(self, config: CspnConfig):
        super().__init__(config=config)
        self.config: CspnConfig = config
        self.dist_std_head = None
        self.dist_mean_head = None
        self.dist_layers = None
        self.sum_param_heads = None
        self.sum_layers = None
        self.feat_layers = None
        self.replace_layer_params()
        self.create_feat_layers(config.F_cond)

    @property
    def device(self):
        return self.dist_mean_head.bias.device

    # ---------------------------------------------------------------------------------- construction
    def replace_layer_params(self):
        """The circuit owns no trainable parameters: Parameters become plain placeholder tensors that ``set_params``
        overwrites for every batch of conditionals (cspn.py:153-171)."""
        sums = [l for l in self._inner_layers if isinstance(l, Sum)] + [self.root, self._sampling_root]
        for layer in sums:
            placeholder = th.zeros_like(layer.weight_param)
            del layer.weight_param
            layer.weight_param = placeholder
        base = self._leaf.base_leaf
        placeholder = th.zeros_like(base.mean_param)
        del base.mean_param
        del base.std_param
        base.mean_param = placeholder
        base.std_param = placeholder

    def create_feat_layers(self, feature_dim: tuple):
        """feat_layers -> {sum_layers -> one head per Sum layer + root} and {dist_layers -> mean head, std head}
        (cspn.py:173-247).  Only flat (1-D) conditionals are supported, as in the reference."""
        assert len(feature_dim) == 1, \
            f"Don't know how to construct feature extraction layers for features of dim {len(feature_dim)}."
        cfg = self.config
        act = cfg.cond_layers_inner_act
        self.feat_layers = _mlp([feature_dim[0]] + list(cfg.feat_layers or []), act, last_is_identity=False)
        with th.no_grad():
            n_feat = int(np.prod(self.feat_layers(th.ones((1, *feature_dim))).shape))
        sum_sizes = [n_feat] + list(cfg.sum_param_layers or [])
        self.sum_layers = _mlp(sum_sizes, act, last_is_identity=True)
        heads = [nn.Linear(sum_sizes[-1], l.weight_param.numel()) for l in self._inner_layers if isinstance(l, Sum)]
        heads.append(nn.Linear(sum_sizes[-1], self.root.weight_param.numel()))
        self.sum_param_heads = nn.ModuleList(heads)
        dist_sizes = [n_feat] + list(cfg.dist_param_layers or [])
        self.dist_layers = _mlp(dist_sizes, act, last_is_identity=True)
        self.dist_mean_head = nn.Linear(dist_sizes[-1], self._leaf.base_leaf.mean_param.numel())
        self.dist_std_head = nn.Linear(dist_sizes[-1], self._leaf.base_leaf.std_param.numel())

    def create_one_hot_in_channel_mapping(self):
        for lay in self._inner_layers:
            if isinstance(lay, CrossProduct):
                lay.one_hot_in_channel_mapping = F.one_hot(lay.unraveled_channel_indices).float().requires_grad_(False)

    # ---------------------------------------------------------------------------------- per-sample parameters
    def set_params(self, feat_inp: th.Tensor):
        """Evaluate the gating network on ``feat_inp`` [B, *F_cond] and install, per conditional, log-normalised sum
        weights [B, d, ic, oc, r] and sigmoid-bounded leaf parameters [B, F, I, R]   (cspn.py:254-301)."""
        B = feat_inp.shape[0]
        self._release_param_graph()
        features = self.feat_layers(feat_inp).flatten(start_dim=1)
        hidden = self.sum_layers(features)
        head = iter(self.sum_param_heads)
        for layer in self._inner_layers:
            if isinstance(layer, Sum):
                shape = (B, layer.in_features, layer.in_channels, layer.out_channels, layer.num_repetitions)
                raw = next(head)(hidden).view(shape)
                layer._raw_weights = bool(self.config.fuse_weight_norm)
                layer.weight_param = raw if layer._raw_weights else F.log_softmax(raw, dim=2)
            else:
                layer.num_conditionals = B
        root = self.root
        shape = (B, root.in_features, root.in_channels, root.out_channels, root.num_repetitions)
        root.weight_param = F.log_softmax(next(head)(hidden).view(shape), dim=2)
        self._sampling_root.weight_param = th.full((B, 1, 1, 1, 1), 1 / self.config.C, device=self.device).log_()
        base = self._leaf.base_leaf
        shape = (B, base.in_features, self.config.I, self.config.R)
        hidden = self.dist_layers(features)
        base.mean_param = base.bounded_means(self.dist_mean_head(hidden).view(shape))
        base.std_param = base.bounded_stds(self.dist_std_head(hidden).view(shape))

    def _release_param_graph(self):
        """Drop the autograd graph of the previous ``set_params`` before building the next one.  The installed
        parameters are the only long-lived references to it; keeping it alive across calls would make the new graph
        share the old AccumulateGrad nodes (and their stream), which breaks CUDA-graph capture of an update."""
        base = self._leaf.base_leaf
        for owner, name in [(l, "weight_param") for l in list(self._inner_layers) + [self.root] if isinstance(l, Sum)] \
                + [(base, "mean_param"), (base, "std_param")]:
            t = getattr(owner, name, None)
            if isinstance(t, th.Tensor) and t.grad_fn is not None:
                setattr(owner, name, t.detach())

    def clear_params(self):
        dev = self.device
        for layer in list(self._inner_layers) + [self.root]:
            if isinstance(layer, Sum):
                layer.weight_param = th.zeros(0, layer.in_features, layer.in_channels, layer.out_channels,
                                              layer.num_repetitions, device=dev)
            else:
                layer.num_conditionals = 0
        self._sampling_root.weight_param = th.zeros(0, 1, 1, 1, 1, device=dev)
        base = self._leaf.base_leaf
        base.mean_param = th.zeros(0, base.in_features, self.config.I, self.config.R, device=dev)
        base.std_param = th.zeros(0, base.in_features, self.config.I, self.config.R, device=dev)

    _pinned_condition = None

    @contextmanager
    def conditioned(self, condition: th.Tensor):
        """Evaluate the gating network ONCE for several calls on the same conditionals (the SAC actor update
        calls ``sample`` and ``recursive_entropy_approx`` on the same observations, sb3.py:132-137 and :158-162):

            with cspn.conditioned(obs):
                ctx = cspn.sample(mode='onehot', condition=obs)
                ent, _ = cspn.recursive_entropy_approx(condition=obs)
            (loss built from both).backward()       # ONE backward: the calls share the gating network's graph

        Inside the block, calls that pass this very tensor as ``condition`` re-use the installed parameters; any other
        ``condition`` is evaluated as usual.  Opt-in because sharing the autograd graph between separate ``backward()``
        calls would need ``retain_graph``."""
        self.set_params(condition)
        self._pinned_condition = condition
        try:
            yield self
        finally:
            self._pinned_condition = None

    # ---------------------------------------------------------------------------------- condition= wrappers
    def forward(self, x: th.Tensor, condition: th.Tensor = None, **kwargs) -> th.Tensor:
        """Conditional log-likelihood log p(x | condition); x [batch, conditionals, F, 1, 1] (cspn.py:72-100)."""
        if condition is not None and condition is not self._pinned_condition:
            self.set_params(condition)
        weight_sets = self._leaf.base_leaf.mean_param.shape[0]
        if x.dim() == 3:
            assert x.shape[1] == weight_sets, \
                f"x provides samples for {x.shape[1]} conditionals but the CSPN weights were set for {weight_sets}"
        elif x.dim() == 2:
            assert x.shape[0] == weight_sets, \
                f"x provides {x.shape[0]} samples but the CSPN weights were set for {weight_sets} conditionals"
            x = x.unsqueeze(0)
        return super().forward(x, **kwargs)

    def recursive_entropy_approx(self, condition: th.Tensor = None, **kwargs) -> Tuple[th.Tensor, Optional[dict]]:
        if condition is not None and condition is not self._pinned_condition:
            self.set_params(condition)
        return super().recursive_entropy_approx(**kwargs)

    def naive_entropy_approx(self, condition: th.Tensor = None, **kwargs) -> th.Tensor:
        if condition is not None and condition is not self._pinned_condition:
            self.set_params(condition)
        return super().naive_entropy_approx(**kwargs)

    def huber_entropy_lb(self, condition: th.Tensor = None, **kwargs):
        """Huber et al. entropy lower bound of p(. | condition)   (cspn.py:112-115)."""
        if condition is not None and condition is not self._pinned_condition:
            self.set_params(condition)
        return super().huber_entropy_lb(**kwargs)

    def sample(self, mode: str = None, condition: th.Tensor = None, class_index=None, evidence: th.Tensor = None,
               **kwargs):
        """Sample x ~ p(. | condition)   (cspn.py:125-145)."""
        if condition is not None and condition is not self._pinned_condition:
            self.set_params(condition)
        assert class_index is None or condition.shape[0] == len(class_index), \
            "The batch size of the condition must equal the length of the class index list if they are provided!"
        return super().sample(mode=mode, class_index=class_index, evidence=evidence, **kwargs)

    def sample_index_style(self, **kwargs):
        return self.sample(mode='index', **kwargs)

    def sample_onehot_style(self, **kwargs):
        return self.sample(mode='onehot', **kwargs)

    def save(self, *args, **kwargs):
        """Pickle a fresh instance carrying only the state_dict (the installed per-sample tensors are not saved)."""
        save_model = CSPN(self.config)
        save_model.load_state_dict(self.state_dict())
        th.save(save_model, *args, **kwargs)
