"""Import the UNMODIFIED reference from /root/reference (build container only) for golden-vector generation.

TEST INFRASTRUCTURE.  Never imported by the product package, by ``-m gpu`` tests, ``smoke()`` or ``bench.py``:
``/root/reference`` does not exist on the GPU box.  Only ``oracle/gen_golden.py`` uses it.

The reference does not import on a current stack; the non-invasive shim is:
  1. flat imports -> put the reference root on ``sys.path``          (rat_spn.py:7-10)
  2. ``np.int`` / ``np.NINF`` were removed from numpy                 (layers.py:391, layers.py:108)
  3. ``Type, Dict, Optional`` are used but never imported             (rat_spn.py:51-52, :592, :628; utils.py:4)
Plus one oracle-only patch for a reference defect: ``IndependentMultivariate.sample`` strips the leaf padding on
dim -2, which is the feature dim in ``index`` mode but the channel dim in ``onehot`` mode
(distributions.py:384-393).  The patch slices the feature dim in both modes (deliberate deviation, documented in
DESIGN.md); index-mode behaviour is unchanged.
"""
import contextlib
import sys
import typing

import numpy as np
import torch

REFERENCE_ROOT = "/root/reference"


def load_reference():
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "NINF"):
        np.NINF = -np.inf
    import utils as ref_utils
    ref_utils.Type, ref_utils.Dict, ref_utils.Optional = typing.Type, typing.Dict, typing.Optional
    import cspn as ref_cspn
    import distributions as ref_dist
    import layers as ref_layers
    import rat_spn as ref_rat_spn

    def patched_sample(self, mode=None, ctx=None):
        if not ctx.is_root:
            ctx = self.prod.sample(ctx=ctx)
            if self._pad:
                if mode == "index":
                    ctx.parent_indices = ctx.parent_indices[..., :-self._pad, :]
                else:
                    ctx.parent_indices = ctx.parent_indices[..., :-self._pad, :, :]
        return self.base_leaf.sample(ctx=ctx, mode=mode)

    ref_dist.IndependentMultivariate.sample = patched_sample
    return ref_rat_spn, ref_cspn, ref_layers, ref_dist, ref_utils


@contextlib.contextmanager
def recorded_noise(ref_layers, ref_dist, record: list):
    """Make the reference's Gumbel (layers.py:227) and normal (distributions.py:300-301) draws observable.

    The replacements draw from torch's global generator exactly like the originals, and append every noise tensor to
    ``record`` in draw order, so the oracle (and the CUDA kernels) can be fed the same noise.
    """
    F = ref_layers.F
    orig_gs = F.gumbel_softmax
    Normal = ref_dist.dist.Normal
    orig_rs = Normal.rsample

    def gumbel_softmax(logits, tau=1, hard=False, eps=1e-10, dim=-1):
        g = -torch.empty_like(logits).exponential_().log()
        record.append(g.detach().clone())
        y = (logits + g) / tau
        y_soft = y.softmax(dim)
        if not hard:
            return y_soft
        index = y_soft.max(dim, keepdim=True)[1]
        y_hard = torch.zeros_like(logits).scatter_(dim, index, 1.0)
        return y_hard - y_soft.detach() + y_soft

    def rsample(self, sample_shape=torch.Size()):
        shape = self._extended_shape(sample_shape)
        eps = torch.randn(shape, dtype=self.loc.dtype, device=self.loc.device)
        record.append(eps.detach().clone())
        return self.loc + eps * self.scale

    F.gumbel_softmax = gumbel_softmax
    Normal.rsample = rsample
    try:
        yield
    finally:
        F.gumbel_softmax = orig_gs
        Normal.rsample = orig_rs
