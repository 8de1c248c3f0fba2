"""RAT-SPN with the public API of the reference's ``rat_spn.py`` (``RatSpnConfig``, ``RatSpn``), evaluated by the
sm_100a kernels of libspn_b200.so.

Differences in *how* (not what) things are computed:
  * forward: the per-repetition permutation, Gaussian density, tanh correction, NaN marginalisation and the leaf
    product are one kernel; every CrossProduct is fused into the Sum above it (the c*c outer sum is never
    materialised); the root merges repetitions inside the kernel (rat_spn.py:189-244 in the reference).
  * sampling (index mode): one kernel per Sum layer (CrossProduct un-ravelling fused) and one leaf kernel that also
    applies the post-processing (rat_spn.py:348-583).
  * no host<->device synchronisation on the hot path (the reference's asserts / .isnan().any() force several).
"""
import logging
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple, Type, Union

import numpy as np
import torch as th
from torch import nn
from torch.nn import functional as F

from . import ops
from .distributions import IndependentMultivariate, truncated_normal_
from .layers import CrossProduct, Sum
from .type_checks import check_valid
from .utils import Sample, provide_evidence

logger = logging.getLogger(__name__)


def invert_permutation(p: th.Tensor):
    """s[p[i]] = i for a permutation p of 0..len(p)-1."""
    s = th.empty_like(p)
    s[p] = th.arange(p.shape[0], dtype=p.dtype, device=p.device)
    return s


@dataclass
class RatSpnConfig:
    """Architecture of a RAT-SPN (names follow the RAT-SPN paper; rat_spn.py:26-83):
    in_features (alias F), tree depth D, sum nodes per region S, leaf distributions per region I, repetitions R,
    root heads C, dropout, leaf_base_class (+ kwargs), tanh_squash."""
    is_ratspn: bool = True
    in_features: int = None
    D: int = None
    S: int = None
    I: int = None
    R: int = None
    C: int = None
    dropout: float = None
    leaf_base_class: Type = None
    leaf_base_kwargs: Dict = None
    tanh_squash: bool = None

    @property
    def F(self):
        return self.in_features

    @F.setter
    def F(self, in_features):
        self.in_features = in_features

    def assert_valid(self):
        self.F = check_valid(self.F, int, 1)
        self.D = check_valid(self.D, int, 1, allow_none=True)
        self.C = check_valid(self.C, int, 1)
        self.S = check_valid(self.S, int, 1)
        self.R = check_valid(self.R, int, 1)
        self.I = check_valid(self.I, int, 1)
        self.dropout = check_valid(self.dropout, float, 0.0, 1.0)
        assert self.leaf_base_class is not None, Exception("RatSpnConfig.leaf_base_class parameter was not set!")
        if self.D is not None and 2 ** self.D > self.F:
            raise Exception(f"The tree depth D={self.D} must be <= {np.floor(np.log2(self.F))} (log2(in_features).")

    def __setattr__(self, key, value):
        if hasattr(self, key):
            super().__setattr__(key, value)
        else:
            raise AttributeError(f"RatSpnConfig object has no attribute {key}")


class RatSpn(nn.Module):
    """Layered tensorised RAT-SPN (https://arxiv.org/abs/1806.01910) on B200 kernels."""
    _inner_layers: nn.ModuleList

    def __init__(self, config: RatSpnConfig):
        super().__init__()
        config.assert_valid()
        assert config.R != config.F, "The number of repetitions can't be the number of input features to the SPN, sorry."
        self.config = config
        if self.config.D is None:
            self.config.D = int(np.log2(self.config.F))
        self._build()
        if self.config.is_ratspn:
            self._init_weights()
        self._make_random_repetition_permutation_indices()
        self.__max_layer_index = len(self._inner_layers) + 1
        self.num_layers = self.__max_layer_index + 1
        self._perm_cache = None
        self._test_noise = None

    # ------------------------------------------------------------------------------------------ construction

    def __setstate__(self, state):
        """Also accepts a whole-model pickle written by the reference implementation (rat_spn.py:1046-1047,
        cspn.py:324-327) once ``install_as_reference_modules()`` has mapped its module names onto this package."""
        super().__setstate__(state)
        self.__dict__.setdefault("_perm_cache", None)
        self.__dict__.setdefault("_test_noise", None)
        self.__dict__["_perm_cache"] = None

    def _build(self):
        """leaf -> X_0 -> [S_j -> X_j] for j = 1..D-1 -> root   (rat_spn.py:246-300)."""
        cfg = self.config
        cardinality = int(np.ceil(cfg.F / (2 ** cfg.D)))
        self._leaf = IndependentMultivariate(
            in_features=cfg.F, out_channels=cfg.I, num_repetitions=cfg.R, cardinality=cardinality,
            dropout=cfg.dropout, tanh_squash=cfg.tanh_squash, leaf_base_class=cfg.leaf_base_class,
            leaf_base_kwargs=cfg.leaf_base_kwargs, ratspn=cfg.is_ratspn)
        layers = [CrossProduct(in_features=self._leaf.out_features, in_channels=cfg.I, num_repetitions=cfg.R)]
        sum_in = cfg.I ** 2
        for i in range(cfg.D - 1, 0, -1):
            layers.append(Sum(in_features=2 ** i, in_channels=sum_in, num_repetitions=cfg.R, out_channels=cfg.S,
                              dropout=cfg.dropout, ratspn=cfg.is_ratspn))
            layers.append(CrossProduct(in_features=2 ** i, in_channels=cfg.S, num_repetitions=cfg.R))
            sum_in = cfg.S ** 2
        self._inner_layers = nn.ModuleList(layers)
        self.root = Sum(in_channels=cfg.R * sum_in, in_features=1, num_repetitions=1, out_channels=cfg.C,
                        ratspn=cfg.is_ratspn)
        self._sampling_root = Sum(in_channels=cfg.C, in_features=1, out_channels=1, num_repetitions=1,
                                  ratspn=cfg.is_ratspn)
        self._sampling_root.weight_param = nn.Parameter(
            th.ones(size=(1, 1, cfg.C, 1, 1)) * th.tensor(1 / cfg.C), requires_grad=False)

    def _init_weights(self):
        """Same call pattern as the reference, including its quirk: the truncated-normal init writes into the
        temporaries returned by the ``weights`` / ``stds`` properties, so parameters keep their randn values
        (rat_spn.py:320-329)."""
        for module in self.modules():
            if hasattr(module, "_init_weights") and module != self:
                module._init_weights()
            elif isinstance(module, Sum):
                truncated_normal_(module.weights, std=0.5)

    def _make_random_repetition_permutation_indices(self):
        """One random feature permutation per repetition, drawn from numpy's global RNG (rat_spn.py:143-153)."""
        perm = th.stack([th.as_tensor(np.random.permutation(self.config.F)) for _ in range(self.config.R)], dim=-1)
        inv = th.stack([invert_permutation(perm[:, r]) for r in range(self.config.R)], dim=-1)
        self.permutation = nn.Parameter(perm, requires_grad=False)
        self.inv_permutation = nn.Parameter(inv, requires_grad=False)

    def _perm32(self) -> Tuple[th.Tensor, th.Tensor]:
        """int32 device copies of (permutation, inv_permutation) for the kernels, refreshed when the buffers change."""
        key = (self.permutation.data_ptr(), self.permutation._version, self.inv_permutation.data_ptr(),
               self.inv_permutation._version, ops.PARAM_EPOCH)
        if self._perm_cache is None or self._perm_cache[0] != key:
            self._perm_cache = (key, self.permutation.data.to(th.int32).contiguous(),
                                self.inv_permutation.data.to(th.int32).contiguous())
        return self._perm_cache[1], self._perm_cache[2]

    # ------------------------------------------------------------------------------------------ small accessors
    @property
    def dtype(self):
        return self._leaf.base_leaf.marginalization_constant.dtype

    @property
    def device(self):
        return self._sampling_root.weight_param.device

    @property
    def max_layer_index(self):
        return self.__max_layer_index

    @property
    def sum_layer_indices(self):
        return list(range(2, self.max_layer_index + 1, 2))

    @property
    def is_ratspn(self):
        return self.config.is_ratspn

    def layer_index_to_obj(self, layer_index):
        if layer_index == 0:
            return self._leaf
        if layer_index < self.__max_layer_index:
            return self._inner_layers[layer_index - 1]
        if layer_index == self.__max_layer_index:
            return self.root
        raise IndexError(f"layer_index must take a value between 0 and {self.max_layer_index}, "
                         f"but it was {layer_index}!")

    @property
    def root_weights_split_by_rep(self):
        """Root log-weights viewed per repetition: [w, 1, ic, C, R]   (rat_spn.py:331-335)."""
        weights = self.root.weights
        w, d, _, oc, _ = weights.shape
        return weights.view(w, d, self.config.R, self.root.in_channels // self.config.R, oc).permute(0, 1, 3, 4, 2)

    means = property(lambda self: self._leaf.base_leaf.means,
                     lambda self, v: setattr(self._leaf.base_leaf, "means", v))
    stds = property(lambda self: self._leaf.base_leaf.stds)
    log_stds = property(lambda self: self._leaf.base_leaf.log_stds)
    var = property(lambda self: self._leaf.base_leaf.var)

    # ------------------------------------------------------------------------------------------ permutation
    def _gather_features(self, x: th.Tensor, index: th.Tensor) -> th.Tensor:
        assert x.size(-1) == 1 or x.size(-1) == self.config.R
        if x.size(-1) == 1:
            x = x.expand(*x.shape[:-1], self.config.R)
        return th.gather(x, dim=-3, index=index.unsqueeze(-2).expand_as(x))

    def apply_permutation(self, x: th.Tensor) -> th.Tensor:
        """x [.., F, *, 1|R] -> features shuffled per repetition (rat_spn.py:155-170).  The forward pass does this
        inside the leaf kernel; this method serves masks and external callers."""
        return self._gather_features(x, self.permutation)

    def apply_inv_permutation(self, x: th.Tensor) -> th.Tensor:
        return self._gather_features(x, self.inv_permutation)

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, x: th.Tensor, layer_index: int = None, x_needs_permutation: bool = True,
                detach_params: bool = False) -> th.Tensor:
        """Log-likelihood of ``x`` [*batch_dims, conditionals, F, 1|I, 1|R] evaluated up to ``layer_index``
        (default: root, output [*batch_dims, w, 1, C, 1]); same contract as rat_spn.py:189-244."""
        cfg = self.config
        if layer_index is None:
            layer_index = self.max_layer_index
        assert x.dtype == self.dtype, f"x has data type {x.dtype} but must have data type {self.dtype}"
        assert x.dim() >= 5, "Input needs at least 5 dims. Did you read the docstring of RatSpn.forward()?"
        assert x.size(-3) == cfg.F and x.size(-2) in (1, cfg.I) and x.size(-1) in (1, cfg.R), \
            f"Shape of last three dims is {tuple(x.shape[-3:])} but must be ({cfg.F}, 1 or {cfg.I}, 1 or {cfg.R})"
        perm32 = self._perm32()[0] if x_needs_permutation else None
        if self.training and cfg.dropout > 0.0:
            return self._forward_layerwise(x, layer_index, x_needs_permutation, detach_params)

        caching = self.root._is_input_cache_enabled
        # (a pass that ends below the first Sum layer stays on the exact fp32 leaf kernels)
        if layer_index >= 2 and self._tensor_core_path_ok(x, caching):
            return self._forward_tensor_core(x, layer_index, perm32, detach_params)
        if self._tensor_core_sets_path_ok(x, caching, layer_index):
            return self._forward_tensor_core_sets(x, layer_index, perm32, detach_params)
        h = self._leaf(x, detach_params=detach_params, perm32=perm32)
        for li in range(1, min(layer_index, self.max_layer_index - 1) + 1):
            layer = self._inner_layers[li - 1]
            if li % 2 == 1:
                if li == layer_index:  # the pass ends on a CrossProduct: materialise it
                    h = layer(h)
                continue
            if caching:
                layer._input_cache_fused = h.detach()
            if layer._raw_weights and ops.sum_raw_supported(h, layer.weight_param):
                # CSPN with fused weight normalisation: the raw head outputs go straight into the streaming kernels
                h = ops.sum_forward_raw(h, layer.weight_param.detach() if detach_params else layer.weight_param)
            else:
                h = ops.sum_forward(h, layer._weights_for(detach_params), "xsum")
        if layer_index == self.max_layer_index:
            if caching:
                self.root._input_cache_fused = h.detach()
            h = ops.sum_forward(h, self.root._weights_for(detach_params), "xroot")
        return h

    # rows below which the exact fp32 kernels are used even when the tensor-core path is available (tiles are 128 rows)
    tensor_core_min_rows = 256
    use_tensor_cores = True

    def _tensor_core_path_ok(self, x: th.Tensor, caching: bool) -> bool:
        """Shared weights (RatSpn), supported channel counts, enough rows to fill 128-row tiles, no evidence cache."""
        cfg = self.config
        if not (self.use_tensor_cores and cfg.is_ratspn) or caching or cfg.D < 2:
            return False
        rows = x.numel() // (x.shape[-3] * x.shape[-2] * x.shape[-1])
        if rows < self.tensor_core_min_rows:
            return False
        if cfg.D > 2 and cfg.I != cfg.S:
            pass  # S_1 pairs leaf channels (c = I), the layers above pair sum channels (c = S); both must be supported
        return ops.tc_supported(cfg.I, cfg.S, cfg.R) and (cfg.D == 2 or ops.tc_supported(cfg.S, cfg.S, cfg.R))

    def _tensor_core_sets_path_ok(self, x: th.Tensor, caching: bool, layer_index: int) -> bool:
        """Per-conditional weights (CSPN, cspn.py:254-301) with MANY rows per weight set -- the entropy estimate scores
        K * n samples per conditional, a wide conditional circuit a whole batch: every inner Sum layer is then one dense
        contraction per weight set and runs on the tcgen05 kernels batched over the sets."""
        cfg = self.config
        if cfg.is_ratspn or not self.use_tensor_cores or caching or cfg.D < 2 or layer_index < 2:
            return False
        w = x.shape[-4]
        rows_per_set = x.numel() // (w * x.shape[-3] * x.shape[-2] * x.shape[-1])
        if rows_per_set < self.tensor_core_min_rows:
            return False
        for layer in self._inner_layers:
            if isinstance(layer, Sum) and (layer.weight_param is None or layer.weight_param.shape[0] != w):
                return False
        return ops.tc_supported(cfg.I, cfg.S, cfg.R) and (cfg.D == 2 or ops.tc_supported(cfg.S, cfg.S, cfg.R))

    def _forward_tensor_core_sets(self, x, layer_index, perm32, detach_params):
        """Leaf and root on the per-conditional fp32 kernels (reference layout), the inner Sum layers up to ``layer_index``
        on the tensor-core kernels with the activations in the internal layout [scope, repetition, set-major rows,
        channel].  (``layer_index`` below the root: the recursive entropy estimate scores the samples of every node of
        a layer under the circuit below it, rat_spn.py:625-750.)"""
        batch_shape, w = x.shape[:-4], x.shape[-4]
        if w == 1 and x.shape[-1] == 1 and x.shape[-2] == 1:   # a single conditional: the leaf kernel writes the internal layout (as for a RatSpn)
            h = self._leaf(x, detach_params=detach_params, perm32=perm32, layout="drbc")    # [d0, R, rows, I]
            d0, R, rows, I = h.shape
        else:
            h = self._leaf(x, detach_params=detach_params, perm32=perm32)           # [*batch, w, d0, I, R]
            d0, I, R = h.shape[-3:]
            h = h.reshape(-1, w, d0, I, R).permute(2, 4, 1, 0, 3)                   # [d0, R, w, rows, I]
            rows = h.shape[3]
            h = h.reshape(d0, R, w * rows, I)
        for li in range(2, min(layer_index, self.max_layer_index - 1) + 1, 2):
            layer = self._inner_layers[li - 1]
            if layer._raw_weights:   # fused weight normalisation: the raw head outputs go in
                wt = layer.weight_param.detach() if detach_params else layer.weight_param
                h = ops.sum_forward_tc(h, wt, True, None)
            else:
                h = ops.sum_forward_tc(h, layer._weights_for(detach_params), False, None)
        d, c = h.shape[0], h.shape[3]
        if layer_index == self.max_layer_index and ops.root_shared_supported(c, self.config.C, R):
            # root on the internal layout, one weight set at a time (few sets with many rows each in this regime)
            lw_root = self.root._weights_for(detach_params)
            outs = [ops.root_forward_shared(h[:, :, i * rows:(i + 1) * rows], lw_root[i:i + 1]) for i in range(w)]
            out = outs[0] if w == 1 else th.stack(outs, dim=1)                      # [rows, (w,) C]
            return out.reshape(*batch_shape, w, 1, self.config.C, 1)
        h = h.reshape(d, R, w, rows, c).permute(3, 2, 0, 4, 1).reshape(*batch_shape, w, d, c, R)
        if layer_index == self.max_layer_index:
            return ops.sum_forward(h, self.root._weights_for(detach_params), "xroot")
        if layer_index % 2 == 1:   # the pass ends on a CrossProduct: materialise it
            h = self._inner_layers[layer_index - 1](h)
        return h

    def _forward_tensor_core(self, x, layer_index, perm32, detach_params):
        """Same evaluation with activations in the internal [scope, repetition, row, channel] layout and every inner
        Sum layer on the tcgen05 kernel; the root (N = C outputs, GEMV-shaped) stays on the exact fp32 kernel."""
        batch_shape = x.shape[:-3]  # (*batch_dims, w)
        h = self._leaf(x, detach_params=detach_params, perm32=perm32, layout="drbc")
        for li in range(2, min(layer_index, self.max_layer_index - 1) + 1, 2):
            layer = self._inner_layers[li - 1]
            # the unnormalised parameter goes in; log_softmax over the in-channels is fused into the kernels
            w = layer.weight_param.detach() if detach_params else layer.weight_param
            h = ops.sum_forward_tc(h, w, True, layer._tc_cache)
        if layer_index == self.max_layer_index:
            lw_root = self.root._weights_for(detach_params)
            if ops.root_shared_supported(h.shape[3], self.config.C, self.config.R):
                out = ops.root_forward_shared(h, lw_root)  # [B, C]
            else:
                out = ops.sum_forward(h, lw_root, "xroot", "drbc")  # [B, 1, 1, C, 1]
            return out.reshape(*batch_shape, 1, self.config.C, 1)
        h = h.permute(2, 0, 3, 1).reshape(*batch_shape, h.shape[0], h.shape[3], h.shape[1])  # -> [*batch, w, d, c, R]
        if layer_index % 2 == 1:
            h = self._inner_layers[layer_index - 1](h)
        return h

    def _forward_layerwise(self, x, layer_index, x_needs_permutation, detach_params):
        """Unfused evaluation (one module after the other); only used when dropout is active in training mode.  With the
        evidence cache enabled the Sum layers also receive the CrossProduct *input* their own input derives from
        (``_input_cache_fused``), which is what the sampling kernels reweight with -- like the reference
        (layers.py:102-103, :189-203) this is the activation before dropout is applied."""
        if x_needs_permutation:
            x = self.apply_permutation(x)
        h = self._leaf(x, detach_params=detach_params)
        below_product = h
        for layer in self._inner_layers[:layer_index]:
            if isinstance(layer, CrossProduct):
                below_product = h
            elif layer._is_input_cache_enabled:
                layer._input_cache_fused = below_product.detach()
            h = layer(h, detach_params=detach_params)
        if layer_index == self.max_layer_index:
            if self.root._is_input_cache_enabled:
                self.root._input_cache_fused = below_product.detach()
            h = h.transpose(-1, -2).flatten(-2, -1).unsqueeze(-1)
            h = self.root(h, detach_params=detach_params)
        return h

    # ------------------------------------------------------------------------------------------ sampling
    def mpe(self, evidence: th.Tensor = None, layer_index=None) -> th.Tensor:
        return self.sample(mode='index', evidence=evidence, layer_index=layer_index, is_mpe=True).sample

    def sample(self, mode: str = None, n: Union[int, Tuple] = 1, class_index=None, evidence: th.Tensor = None,
               is_mpe: bool = False, layer_index: int = None, do_sample_postprocessing: bool = True,
               post_processing_kwargs: dict = None, noise: Optional[List[th.Tensor]] = None,
               seed: Optional[int] = None, sample_offset: int = 0) -> Sample:
        """Ancestral sampling; arguments and the returned ``Sample`` follow rat_spn.py:348-498.

        ``seed`` / ``sample_offset`` (extension, index mode, batch-sharded sampling over several GPUs): the in-kernel
        Philox noise is keyed by (seed, global sample index), so a rank that draws samples [sample_offset, sample_offset
        + n) of a job with the job's common ``seed`` gets exactly the rows the single-GPU call with that seed returns.

        ``noise`` (extension, used by the parity tests): list of noise tensors consumed in the reference's draw
        order -- one Gumbel tensor [nodes, *n, w, d, ic, r] per Sum layer top-down, then the leaf's standard-normal
        tensor.  Without it, index mode draws in-kernel (Philox keyed by the sample index).
        """
        post_processing_kwargs = dict(post_processing_kwargs or {})
        assert mode is not None, "A sampling mode must be provided!"
        assert class_index is None or class_index.dim() == 1, "Class index must have shape [conditionals]"
        if class_index is not None and evidence is not None:
            assert class_index.shape[0] == evidence.shape[1], \
                "class_index must have same number of conditionals as evidence"
        assert evidence is None or evidence.dtype == self.dtype, \
            f"evidence has data type {evidence.dtype} but must have data type {self.dtype}"
        if layer_index is None:
            layer_index = self.max_layer_index
        n = (n,) if isinstance(n, int) else tuple(n)
        if evidence is not None:
            n = (*n, *evidence.shape[:-4])
        if noise is None:
            noise = self._test_noise  # parity-test hook: a shared queue consumed across nested sample() calls

        with provide_evidence(self, evidence, requires_grad=(mode == 'onehot')):
            ctx = Sample(n=n, is_mpe=is_mpe, sampling_mode=mode, needs_squashing=self.config.tanh_squash,
                         sampled_with_evidence=evidence is not None)
            fuse_post = do_sample_postprocessing and not post_processing_kwargs
            if mode == 'index':
                ctx, fused = self._sample_index(ctx, class_index, evidence, layer_index, fuse_post, noise, seed,
                                                sample_offset)
            elif mode == 'onehot':
                from .diffsample import sample_onehot
                ctx, fused = sample_onehot(self, ctx, class_index, evidence, layer_index, noise, fuse_post)
            else:
                raise ValueError(f"unknown sampling mode {mode}")
        if do_sample_postprocessing and not fused:
            ctx = self.sample_postprocessing(ctx, evidence, **post_processing_kwargs)
        return ctx

    def _num_weight_sets(self) -> int:
        return self._leaf.base_leaf.mean_param.shape[0]

    def _sample_index(self, ctx: Sample, class_index, evidence, layer_index, fuse_post, noise, user_seed=None,
                      sample_offset=0):
        cfg = self.config
        dev = self.device
        n = tuple(ctx.n)
        W = self._num_weight_sets()
        if evidence is not None:
            w = evidence.shape[-4]
            Qe = int(np.prod(evidence.shape[:-3]))
        else:
            w = class_index.shape[0] if class_index is not None else W
            Qe = 1
        is_mpe = ctx.is_mpe
        seed = ops.new_seed(dev, is_mpe or noise is not None)  # host integer; a device word under graph capture
        if user_seed is not None and not is_mpe and noise is None:
            seed = int(user_seed)
        q_off = 0
        if sample_offset:
            # items are enumerated (node, *n, conditional): a constant shift only when there is a single root node
            assert len(n) == 1 and (class_index is not None or cfg.C == 1) and layer_index == self.max_layer_index, \
                "sample_offset needs a one-dimensional sample count and a single root node (C = 1 or class_index)"
            q_off = int(sample_offset) * w

        def take(shape):
            if is_mpe or noise is None:
                return None
            t = noise.pop(0)
            assert tuple(t.shape) == tuple(shape), f"noise shape {tuple(t.shape)} != {tuple(shape)}"
            return t.to(device=dev, dtype=th.float32).contiguous()

        def column_major(lw):
            """Shared weights of a wide layer, many samples: in-channel-contiguous copy (ops.sampling_weights)."""
            if lw.shape[0] == 1 and lw.shape[2] >= 64 and Q >= 64:
                return ops.sampling_weights(lw), True
            return lw, False

        rep = None
        parent = None          # int32 [Q, d, Rs]: channel chosen at the CrossProduct currently above us
        parent_strides = None
        start = layer_index
        if layer_index == self.max_layer_index:
            ctx.scopes = 1
            lwr = self.root.weights
            K = lwr.shape[2]
            ic_top = K // cfg.R
            if class_index is not None:
                nodes = 1
                node = th.as_tensor(class_index, device=dev).to(th.int32).view(*([1] * (len(n) + 1)), w)
            else:
                nodes = cfg.C
                node = th.arange(nodes, device=dev, dtype=th.int32).view(nodes, *([1] * (len(n) + 1)))
            lead = (nodes, *n, w)
            Q = int(np.prod(lead))
            pa = node.expand(*lead).reshape(Q).contiguous()
            g = take((*lead, 1, K, 1))
            xc = self.root._input_cache_fused if self.root._is_input_cache_enabled else None
            lwt, tr = column_major(lwr)
            k0 = ops.sample_sum(lwt, cfg.R, Q, 1, 1, pa, (1, 0, 0), False, None, xc, Qe, g, is_mpe, seed, q_offset=q_off, transposed=tr)
            k0 = k0.view(Q)
            rep = th.div(k0, ic_top, rounding_mode='floor').to(th.int32)
            parent = (k0 - rep * ic_top).to(th.int32).view(Q, 1, 1)
            parent_strides = (1, 1, 1)
            Rs = 1
            start = layer_index - 1
            ctx.has_rep_dim = False
        else:
            Rs = cfg.R
            if layer_index == 0:
                return self._sample_leaf_root(ctx, w, noise), False
            layer = self.layer_index_to_obj(layer_index)
            nodes = layer.in_channels ** 2 if isinstance(layer, CrossProduct) else layer.out_channels
            lead = (nodes, *n, w)
            Q = int(np.prod(lead))
            node = th.arange(nodes, device=dev, dtype=th.int32).view(nodes, *([1] * (len(n) + 1)))
            pa = node.expand(*lead).reshape(Q).contiguous()
            if isinstance(layer, CrossProduct):
                ctx.scopes = layer.in_features // layer.cardinality
                parent, parent_strides = pa, (1, 0, 0)
                start = layer_index - 1  # the un-ravelling of this CrossProduct is fused into the next step
            else:
                ctx.scopes = layer.in_features
                d = layer.in_features
                g = take((*lead, d, layer.in_channels, cfg.R))
                xc = layer._input_cache_fused if layer._is_input_cache_enabled else None
                lwt, tr = column_major(layer.weights)
                parent = ops.sample_sum(lwt, 0, Q, d, Rs, pa, (1, 0, 0), False, None, xc, Qe, g, is_mpe, seed,
                                        q_offset=q_off, transposed=tr)
                parent_strides = (d * Rs, Rs, 1)
                start = layer_index - 1

        # remaining Sum layers, top -> bottom (CrossProducts are fused into the step below them)
        for li in range(start, 0, -1):
            if li % 2 == 1:
                continue
            layer = self._inner_layers[li - 1]
            d = layer.in_features
            g = take((*lead, d, layer.in_channels, Rs))
            xc = layer._input_cache_fused if layer._is_input_cache_enabled else None
            lwt, tr = column_major(layer.weights)
            parent = ops.sample_sum(lwt, 0, Q, d, Rs, parent, parent_strides, True, rep, xc, Qe, g, is_mpe,
                                    seed, q_offset=q_off, transposed=tr)
            parent_strides = (d * Rs, Rs, 1)

        base = self._leaf.base_leaf
        eps = take((*lead, cfg.F) if rep is not None else (*lead, cfg.F, Rs))
        inv32 = self._perm32()[1] if fuse_post else None
        ev, ev_strides = None, (0, 0, 0)
        if fuse_post and evidence is not None:
            assert nodes == 1, "sampling with evidence needs a single root node (C = 1 or class_index)"
            ev = evidence.to(th.float32).contiguous()
            r_ev = ev.shape[-1]
            ev_strides = (cfg.F * ev.shape[-2] * r_ev, ev.shape[-2] * r_ev, 1 if r_ev > 1 else 0)
        squash = bool(fuse_post and cfg.tanh_squash and ctx.needs_squashing)
        y, ch = ops.sample_leaf(base.means, base.stds, Q, self._leaf.cardinality, Rs, parent, parent_strides, rep, eps,
                                is_mpe, inv32, ev, ev_strides, Qe, squash, True, seed, q_offset=q_off)
        ctx.parent_indices = ch.view(*lead, cfg.F, Rs).long()
        if rep is not None:
            ctx.repetition_indices = rep.view(*lead).long()
        if fuse_post:
            ctx.sample = y.view(*lead, cfg.F, Rs)
            ctx.permutation_inverted = True
            ctx.has_rep_dim = True
            ctx.evidence_filled_in = evidence is not None
            if squash:
                ctx.needs_squashing = False
        else:
            ctx.sample = y.view(*lead, cfg.F) if rep is not None else y.view(*lead, cfg.F, Rs)
        return ctx, fuse_post

    def _sample_leaf_root(self, ctx: Sample, w: int, noise) -> Sample:
        """layer_index = 0: every leaf draws from itself, result [I, n, w, F, R]  (distributions.py:257-260)."""
        assert len(ctx.n) == 1
        base = self._leaf.base_leaf
        mu = base.means
        mu = mu.expand(*ctx.n, w, *mu.shape[1:])
        if ctx.is_mpe:
            y = mu
        else:
            eps = noise.pop(0).to(mu.device) if noise is not None else th.randn(mu.shape, device=mu.device)
            y = mu + base.stds.expand_as(mu) * eps
        ctx.scopes = self._leaf.in_features // self._leaf.cardinality
        ctx.sample = y.permute(3, 0, 1, 2, 4)
        return ctx

    def sample_postprocessing(self, ctx: Sample, evidence: th.Tensor = None, split_by_scope: bool = False,
                              invert_permutation: bool = True) -> Sample:
        """Optional scope splitting, inverse permutation, evidence fill-in and tanh squashing of raw samples
        (rat_spn.py:500-583).  The default index-mode path does this inside the leaf kernel already."""
        sample = ctx.sample
        F_ = self.config.F
        if split_by_scope:
            assert not ctx.is_split_by_scope, "Samples are already split by scope!"
            assert F_ % ctx.scopes == 0, "Size of entire scope is not divisible by the number of scopes"
            per = F_ // ctx.scopes
            sample = sample.unsqueeze(1).repeat(1, ctx.scopes, *([1] * (sample.dim() - 1)))
            fdim = -2 if ctx.has_rep_dim else -1
            scope_of_f = th.arange(F_, device=sample.device) // per
            keep = scope_of_f.view(1, F_) == th.arange(ctx.scopes, device=sample.device).view(-1, 1)  # [s, F]
            shape = [1] * sample.dim()
            shape[1], shape[fdim] = ctx.scopes, F_
            sample = th.where(keep.view(shape), sample, th.full_like(sample, float('nan')))
            ctx.is_split_by_scope = True
        if invert_permutation:
            assert not ctx.permutation_inverted, "Permutations are already inverted on the samples!"
            inv = self.inv_permutation
            if ctx.repetition_indices is not None:
                sel = inv.t()[ctx.repetition_indices]
            elif ctx.onehot_repetition is not None:
                sel = inv.t()[ctx.onehot_repetition]
            elif ctx.sampling_mode == 'onehot' and ctx.parent_indices is not None and not ctx.has_rep_dim:
                sel = (inv * ctx.parent_indices.detach().sum(-2).long()).sum(-1)
            else:
                # one selector per repetition for every sample; the scope dim of split samples is added below
                sel = inv.view(*([1] * (sample.dim() - 2 - int(ctx.is_split_by_scope))), *inv.shape)
            if ctx.is_split_by_scope:
                sel = sel.unsqueeze(1)
            sample = th.gather(sample, dim=-2 if ctx.has_rep_dim else -1, index=sel.expand_as(sample))
            ctx.permutation_inverted = True
            if not ctx.has_rep_dim:
                sample = sample.unsqueeze(-1)
                ctx.has_rep_dim = True
        if evidence is not None:
            ev = evidence.expand(*ctx.n, *evidence.shape[-4:]).movedim(-2, 0)
            assert ev.shape == sample.shape, f"evidence {tuple(ev.shape)} vs sample {tuple(sample.shape)}"
            sample = th.where(th.isnan(ev), sample, ev)
            ctx.evidence_filled_in = True
        if self.config.tanh_squash and ctx.needs_squashing:
            sample = sample.clamp(-6.0, 6.0).tanh()
            ctx.needs_squashing = False
        ctx.sample = sample
        return ctx

    def sample_index_style(self, **kwargs):
        return self.sample(mode='index', **kwargs)

    def sample_onehot_style(self, **kwargs):
        return self.sample(mode='onehot', **kwargs)

    # ------------------------------------------------------------------------------------------ entropy
    def weigh_tensors(self, layer_index, tensors: List = None, return_weight_ent=False) \
            -> Tuple[List[th.Tensor], Optional[th.Tensor]]:
        """sum_i w_i * t_i over the in-channels of a Sum layer for every tensor in ``tensors``; the root's weights are
        viewed per repetition and additionally summed over repetitions (rat_spn.py:591-623)."""
        layer = self.layer_index_to_obj(layer_index)
        assert isinstance(layer, Sum), "Responsibilities can only be computed for Sum layers!"
        log_w = layer.weights
        lin_w = log_w.exp()
        weight_entropy = -(lin_w * log_w).sum(dim=2) if return_weight_ent else None
        if layer_index < self.max_layer_index:
            return [(lin_w * t).sum(dim=2) for t in tensors], weight_entropy
        w, _, ic, oc, _ = lin_w.shape
        r = self.config.R
        per_rep = lin_w.view(w, 1, r, ic // r, oc).permute(0, 1, 3, 4, 2)  # [w, 1, ic/r, oc, r]
        if weight_entropy is not None:
            weight_entropy = weight_entropy.squeeze(-1)
        return [(per_rep * t).sum(dim=-3).sum(dim=-1) for t in tensors], weight_entropy

    def _permuted_marginal_mask(self, marginal_mask):
        if marginal_mask is None:
            return None
        shape_should = tuple(self._leaf.base_leaf.mean_param.shape[:2])
        assert tuple(marginal_mask.shape) == shape_should, \
            f"marginal_mask is of shape {tuple(marginal_mask.shape)} but needs to have shape {shape_should}"
        m = marginal_mask.clone().bool().unsqueeze(-1).unsqueeze(-1)
        return self.apply_permutation(m)  # [w, F, 1, R]

    def layer_entropy_approx(self, layer_index=0, child_entropies: th.Tensor = None, child_ll: th.Tensor = None,
                             sample_size=5, aux_with_grad=False, verbose=False, marginal_mask=None) \
            -> Tuple[th.Tensor, Optional[Dict]]:
        """Entropy estimate of every node of one layer from the entropies of its children (rat_spn.py:625-750):
        leaves score their own samples; entropies add across a CrossProduct; a Sum node adds its weight entropy, the
        weighted child entropies and the weighted log-responsibility of each child's own samples."""
        assert child_entropies is not None or layer_index == 0, \
            "When sampling from a layer other than the leaf layer, child entropies must be provided!"
        metrics = {}
        mask_p = self._permuted_marginal_mask(marginal_mask)
        sampling_mode = 'onehot' if th.is_grad_enabled() else 'index'
        nan = float('nan')

        if layer_index == 0:
            if child_ll is None:
                ctx = self.sample(mode=sampling_mode, n=sample_size, layer_index=0, is_mpe=False,
                                  do_sample_postprocessing=False)
                x = ctx.sample.permute(1, 2, 3, 0, 4)  # [n, w, F, I, R]
                if mask_p is not None:
                    x = th.where(mask_p.expand_as(x), th.full_like(x, nan), x)
                child_ll = self.forward(x=x, layer_index=0, x_needs_permutation=False)
            node_entropies = -child_ll.mean(dim=0)
        else:
            layer = self.layer_index_to_obj(layer_index)
            if isinstance(layer, CrossProduct):
                node_entropies = layer(child_entropies)
            else:
                if child_ll is None:
                    ctx = self.sample(mode=sampling_mode, n=sample_size, layer_index=layer_index - 1, is_mpe=False,
                                      do_sample_postprocessing=False)
                    xs = ctx.sample
                    if xs.dim() == 4:
                        xs = xs.unsqueeze(-1)
                    xs = xs.unsqueeze(-2)  # [K, n, w, F, 1, R]: all K child nodes of every repetition were sampled
                    if mask_p is not None:
                        xs = th.where(mask_p.expand_as(xs), th.full_like(xs, nan), xs)
                    child_ll = self.forward(x=xs, layer_index=layer_index - 1, x_needs_permutation=False,
                                            detach_params=not aux_with_grad).mean(dim=1)
                # child_ll [K, w, d, K, R]: LL of node k's samples under every node k' (same scope, repetition)
                if layer_index == self.max_layer_index:
                    log_weights = self.root_weights_split_by_rep
                    if not aux_with_grad:
                        log_weights = log_weights.detach()
                    ll = ops.sum_forward(child_ll, log_weights.contiguous(), "sum")
                else:
                    ll = layer.forward(child_ll, detach_params=not aux_with_grad)
                    log_weights = layer.weights if aux_with_grad else layer.weights.detach()
                ll = ll.permute(1, 2, 0, 3, 4)  # [w, d, K, oc, R]
                own_ll = th.diagonal(child_ll, dim1=0, dim2=3).permute(0, 1, 3, 2)  # [w, d, K, R]
                if not verbose:
                    # one fused kernel per direction (spn_entropy_combine_*): weight entropy + weighted child entropies +
                    # weighted log-responsibilities; `weights_g` carries the gradient of the mixture weights, the
                    # responsibility's own log-weight term only does so with aux_with_grad (it is detached otherwise)
                    is_root = layer_index == self.max_layer_index
                    weights_g = self.root_weights_split_by_rep if is_root else layer.weights
                    node_entropies = ops.entropy_combine(weights_g, child_entropies, own_ll, ll, reduce_r=is_root,
                                                         lwd_has_grad=aux_with_grad)
                else:
                    responsibility = log_weights + own_ll.unsqueeze(-2) - ll
                    [weighted_ch_ents, weighted_aux_resp], weight_entropy = self.weigh_tensors(
                        layer_index=layer_index, tensors=[child_entropies.unsqueeze(3), responsibility],
                        return_weight_ent=True)
                    node_entropies = weight_entropy + weighted_ch_ents + weighted_aux_resp
                    metrics.update({'weight_entropy': weight_entropy.detach(),
                                    'responsibility': responsibility.detach()})
        logging_dict = {}
        if verbose:
            metrics.update({'node_entropy': node_entropies.detach()})
            logging_dict = self.log_dict_from_metric(layer_index, metrics)
        return node_entropies, logging_dict

    def log_dict_from_metric(self, layer_index: int, metrics: Dict, rep_dim=-1, batch_dim: Optional[int] = 0):
        """{"lay{L}/rep{r}/{metric}/{min,max,mean,std}": float}; the statistics of all repetitions are reduced on the
        device and fetched with one transfer per metric (the reference issues 4*R .item() syncs, rat_spn.py:752-774)."""
        log_dict = {}
        # the number of repetitions reported is that of the FIRST metric, for every metric (rat_spn.py:754)
        n_rep = list(metrics.values())[0].size(rep_dim)
        for key, metric in metrics.items():
            if metric is None:
                continue
            m = metric.to(self.device).movedim(rep_dim, 0)[:n_rep]
            R = m.shape[0]
            flat = m.reshape(R, -1)
            stats = [flat.min(dim=1).values, flat.max(dim=1).values, flat.mean(dim=1)]
            if batch_dim is not None:
                bd = batch_dim if batch_dim >= 0 else batch_dim + metric.dim()
                stats.append(m.std(dim=bd + 1).reshape(R, -1).mean(dim=1))
            host = th.stack(stats).cpu().tolist()
            for rep in range(R):
                base = f"lay{layer_index}/rep{rep}/{key}"
                log_dict[f"{base}/min"], log_dict[f"{base}/max"], log_dict[f"{base}/mean"] = \
                    host[0][rep], host[1][rep], host[2][rep]
                if batch_dim is not None and not np.isnan(host[3][rep]):
                    log_dict[f"{base}/std"] = host[3][rep]
        return log_dict

    def recursive_entropy_approx(self, sample_size=10, aux_with_grad: bool = False, verbose=False,
                                 marginal_mask: th.Tensor = None) -> Tuple[th.Tensor, Optional[Dict]]:
        """Bottom-up recursive entropy approximation of the root nodes, flattened to [w * C]  (rat_spn.py:776-797)."""
        logging_dict = {}
        child_entropies = None
        for layer_index in range(self.num_layers):
            child_entropies, layer_log = self.layer_entropy_approx(
                layer_index=layer_index, child_entropies=child_entropies, sample_size=sample_size,
                aux_with_grad=aux_with_grad, verbose=verbose, marginal_mask=marginal_mask)
            logging_dict.update(layer_log)
        return child_entropies.flatten(), logging_dict

    def naive_entropy_approx(self, sample_size=100, layer_index: int = None, sample_with_grad=False,
                             marginal_mask: th.Tensor = None):
        """Monte-Carlo entropy -E[log p(x)] from the model's own samples (rat_spn.py:799-823).  Like the reference,
        the raw (not inverse-permuted) samples are passed through ``forward`` with its default permutation."""
        if layer_index is None:
            layer_index = self.max_layer_index
        if marginal_mask is not None:
            shape_should = tuple(self._leaf.base_leaf.mean_param.shape[:2])
            assert tuple(marginal_mask.shape) == shape_should
            marginal_mask = marginal_mask.clone().bool()
        kw = dict(n=sample_size, layer_index=layer_index, do_sample_postprocessing=False)
        if sample_with_grad:
            samples = self.sample(mode='onehot', **kw).sample
        else:
            with th.no_grad():
                samples = self.sample(mode='index', **kw).sample
        if marginal_mask is not None:
            samples = th.where(marginal_mask.expand_as(samples), th.full_like(samples, float('nan')), samples)
        samples = samples.movedim(0, -1).unsqueeze(-1)  # 'o...d -> ...do' then a repetition dim
        return -self.forward(samples, layer_index=layer_index).mean()

    def huber_entropy_lb(self, layer_index: int = None, verbose=True, detach_weights: bool = False,
                         add_sub_weight_ent: bool = False, marginal_mask: th.Tensor = None):
        """Entropy lower bound for mixtures of Huber et al. (2008), evaluated layer by layer (rat_spn.py:825-951):
        H(sum_k w_k p_k) >= -sum_k w_k log sum_j w_j z_kj with the pairwise overlaps z_kj = int p_k p_j dx.

        The pass carries the log-overlaps L[w, scope, a, b, ra, rb] between node a of repetition ra and node b of
        repetition rb.  Leaves: int N(x; m1, v1) N(x; m2, v2) dx = N(m1; m2, v1 + v2); overlaps multiply across the
        scopes joined by a Product / CrossProduct (logs add) and mix linearly through Sum weights.  Returns
        (bound for the nodes of ``layer_index`` -- [w, d, oc, R], or [w, 1, C] at the root -- , logging dict).
        Like the reference, the leaf Product requires F to be a multiple of its cardinality and the pairing of
        repetitions uses the permutation of the second repetition index.
        """
        logging = {}
        entropy_lb = None
        if layer_index is None:
            layer_index = self.max_layer_index
        assert 1 <= layer_index <= self.max_layer_index, "layer_index must address an inner layer or the root"
        base = self._leaf.base_leaf
        if marginal_mask is not None:
            shape_should = tuple(base.mean_param.shape[:2])
            assert tuple(marginal_mask.shape) == shape_should, \
                f"marginal_mask is of shape {tuple(marginal_mask.shape)} but needs to have shape {shape_should}"
            marginal_mask = marginal_mask.clone().bool()

        # leaf overlaps in the original feature order, [w, F, a, b, ra, rb]
        mu = self.apply_inv_permutation(self.means)
        var = self.apply_inv_permutation(self.var)
        v = var[:, :, :, None, :, None] + var[:, :, None, :, None, :]
        diff = mu[:, :, :, None, :, None] - mu[:, :, None, :, None, :]
        L = -0.5 * diff * diff / v - 0.5 * v.log() - 0.5 * np.log(2.0 * np.pi)
        if marginal_mask is not None:
            L = L.masked_fill(marginal_mask.view(*marginal_mask.shape, 1, 1, 1, 1), 0.0)
        L = th.gather(L, dim=1, index=self.permutation.view(1, -1, 1, 1, 1, self.config.R).expand_as(L))
        w, F_, I_, _, R, _ = L.shape
        card = self._leaf.prod.cardinality
        assert F_ % card == 0, "huber_entropy_lb needs F to be a multiple of the leaf cardinality"
        L = L.view(w, F_ // card, card, I_, I_, R, R).sum(2)

        for i in range(1, layer_index + 1):
            layer = self.layer_index_to_obj(i)
            if isinstance(layer, CrossProduct):
                left, right = layer.split_shuffled_scopes(L, 1)  # each [w, d, o, o, R, R]
                w, d, o = left.shape[:3]
                # pair channel (a', a) of the new node = (right child a', left child a), likewise (b', b)
                L = (right[:, :, :, None, :, None] + left[:, :, None, :, None, :]).reshape(w, d, o * o, o * o, R, R)
                continue
            log_w = layer.weights
            weight_entropy = -(log_w * log_w.exp()).sum(dim=2)
            is_root = i == self.max_layer_index
            weights = self.root_weights_split_by_rep if is_root else log_w
            if detach_weights:
                weights = weights.detach()
            # M[w, d, a, o, ra, rb] = log sum_b w[b, o, rb] z[a, b, ra, rb]
            M = (L.unsqueeze(-3) + weights.unsqueeze(2).unsqueeze(-2)).logsumexp(3)
            if add_sub_weight_ent:
                we = weight_entropy.unsqueeze(2).unsqueeze(-1)
                M = M - we + we.detach()
            if is_root:
                M = M.logsumexp(-1)  # the root mixes over repetitions as well
            if verbose or i == layer_index:
                own = M if is_root else th.diagonal(M, dim1=-2, dim2=-1)
                entropy_lb = -(weights.exp() * own).sum(dim=-3)
                if is_root:
                    entropy_lb = entropy_lb.sum(-1)
            if i < layer_index:
                # L'[w, d, o1, o2, ra, rb] = log sum_a w[a, o2, ra] exp(M[a, o1, ra, rb])
                L = (M.unsqueeze(-3) + weights.unsqueeze(-3).unsqueeze(-1)).logsumexp(2)
            if verbose:
                metrics = {'weight_entropy': weight_entropy.detach(), 'huber_entropy_LB': entropy_lb.detach()}
                logging.update(self.log_dict_from_metric(i, metrics, batch_dim=None))
        return entropy_lb, logging

    # ------------------------------------------------------------------------------------------ misc API
    def debug__set_weights_dirac(self, layer_index):
        self._debug_set_weights(layer_index, dirac=True)

    def debug__set_root_weights_dirac(self):
        self.debug__set_weights_dirac(self.max_layer_index)

    def debug__set_weights_uniform(self, layer_index):
        self._debug_set_weights(layer_index, dirac=False)

    def _debug_set_weights(self, layer_index, dirac: bool):
        layer = self.layer_index_to_obj(layer_index)
        assert isinstance(layer, Sum), "Given layer index is not a Sum layer!"
        with th.no_grad():
            weights = th.full_like(layer.weights, -100.0)
            if dirac:
                weights[0, 0, 0, :, 0] = -1e-3
            weights = weights.log_softmax(dim=2)
        is_param = isinstance(layer.weight_param, nn.Parameter)
        del layer.weight_param
        layer.weight_param = nn.Parameter(weights) if is_param else weights

    def shape_like_crossprod_input_mapping(self, t: th.Tensor, ic_dim=-3):
        """View an I^2 channel axis as (left channel, right channel)   (rat_spn.py:1023-1044)."""
        ic_dim = ic_dim % t.dim()
        c = int(round(np.sqrt(t.size(ic_dim))))
        return t.view(*t.shape[:ic_dim], c, c, *t.shape[ic_dim + 1:])

    def save(self, *args, **kwargs):
        th.save(self, *args, **kwargs)
