"""Sampling context and evidence context manager (API of the reference's utils.py:52-145)."""
import os
from contextlib import contextmanager, nullcontext
from dataclasses import dataclass, fields
from typing import List, Tuple, Union

import torch as th
from torch import nn


@contextmanager
def provide_evidence(spn: nn.Module, evidence: th.Tensor, requires_grad=False):
    """Run ``spn(evidence)`` once with the layer input caches switched on, so that a sampling pass inside the
    ``with`` block draws from the posterior given the evidence (utils.py:52-86).  The caches are dropped on exit.
    Gradients are only recorded when ``requires_grad`` (differentiable sampling)."""
    cached = [m for m in spn.modules() if hasattr(m, "_enable_input_cache")]
    with (nullcontext() if requires_grad else th.no_grad()):
        for m in cached:
            m._enable_input_cache()
        try:
            if evidence is not None:
                spn(evidence)
            yield
        finally:
            for m in cached:
                m._disable_input_cache()


@dataclass
class Sample:
    """State carried through a top-down sampling pass (utils.py:89-145)."""
    n: Union[th.Size, Tuple[int]] = None          # sample shape per conditional
    scopes: int = None                            # number of scopes of the layer sampling started at
    parent_indices: th.Tensor = None              # channel choice handed to the next (lower) layer
    repetition_indices: th.Tensor = None          # chosen repetition per sample (root sampling, index mode)
    onehot_repetition: th.Tensor = None           # onehot mode: index of the repetition whose one-hots are non-zero
    is_mpe: bool = False
    sample: th.Tensor = None
    has_rep_dim: bool = True
    sampled_with_evidence: bool = None
    sampling_mode: str = None
    evidence_filled_in: bool = False
    is_split_by_scope: bool = False
    needs_squashing: bool = None
    permutation_inverted: bool = False            # False: features are still in per-repetition leaf order

    def __setattr__(self, key, value):
        if key not in {f.name for f in fields(self)}:
            raise AttributeError(f"SamplingContext object has no attribute {key}")
        super().__setattr__(key, value)

    @property
    def is_root(self):
        return self.parent_indices is None and self.repetition_indices is None

    def assert_correct_indices(self):
        """Structural invariants of the context (utils.py:129-145)."""
        if self.sampling_mode == 'onehot':
            assert self.repetition_indices is None, "Onehot sampling: Repetition indices cannot exist!"
            if self.parent_indices is not None:
                tot = self.parent_indices.detach().sum(-2)
                if not self.has_rep_dim:
                    tot = tot.sum(-1)
                assert (tot == 1.0).all(), "Onehot sampling: every feature must select exactly one channel."
        else:
            k = len(self.n) + 1
            assert self.parent_indices is None or tuple(self.parent_indices.shape[1:k]) == tuple(self.n)
            assert self.repetition_indices is None or tuple(self.repetition_indices.shape[1:k]) == tuple(self.n)


def non_existing_folder_name(base_dir: str, folder_name: str, create_path: bool = True) -> str:
    """``folder_name``, or ``folder_name__<k>`` with the smallest k >= 1 that does not exist under ``base_dir`` yet
    (utils.py:12-25); optionally creates it."""
    candidate, k = folder_name, 0
    while os.path.exists(os.path.join(base_dir, candidate)):
        k += 1
        candidate = f"{folder_name}__{k}"
    if create_path:
        os.makedirs(os.path.join(base_dir, candidate), exist_ok=False)
    return candidate


def flat_index_to_tensor_index(index: int, tensor_shape) -> List[int]:
    """Row-major flat offset -> per-dimension indices (utils.py:28-39)."""
    if isinstance(index, th.Tensor):
        index = index.item()
    out = []
    for size in reversed(tuple(tensor_shape)[1:]):
        index, rem = divmod(int(index), int(size))
        out.append(rem)
    out.append(int(index))
    return out[::-1]


def tensor_index_to_flat_index(index: List[int], tensor_shape) -> int:
    """Per-dimension indices -> row-major flat offset (utils.py:42-49)."""
    flat = 0
    for ind, size in zip(index, tuple(tensor_shape)):
        flat = flat * int(size) + int(ind)
    return flat
