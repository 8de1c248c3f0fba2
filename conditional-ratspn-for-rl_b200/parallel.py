"""Batch-sharded data parallelism: one process per GPU, parameters replicated, the batch split contiguously across
ranks, and ONE flat all-reduce of the gradients per step (NCCL over NVLink/NVSwitch; gloo in the CPU tests).

The reference is single-process; this is the multi-GPU plumbing BASELINE.json asks for.  Nothing else is partitioned:
forward, sampling and entropy need no collective because every sample / conditional is independent.
"""
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``n_items`` owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, dim: int = 0, world_size: int = None, rank: int = None) -> torch.Tensor:
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_bounds(x.shape[dim], world_size, rank)
    return x.narrow(dim, lo, hi - lo)


class FlatGradAllReducer:
    """Keeps every trainable parameter's ``.grad`` as a view into one flat fp32 buffer.

    ``overlap=False``: the gradient exchange is a single ``all_reduce`` on that buffer after the backward pass.
    ``overlap=True``: every parameter's slice is all-reduced asynchronously as soon as autograd has finished accumulating
    it (post-accumulate-grad hook), so the collectives of the upper layers run on NCCL's stream while the backward kernels
    of the lower layers are still executing; ``all_reduce()`` then only waits for the outstanding handles.  The largest
    bucket of a RAT-SPN (the bottom Sum layer, ~half of all bytes) is produced before the leaf backward, which hides it."""

    def __init__(self, params: Iterable[torch.nn.Parameter], average: bool = True, overlap: bool = False):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev, dt = self.params[0].device, self.params[0].dtype
        assert all(p.device == dev and p.dtype == dt for p in self.params)
        # every slice starts on a 256-byte boundary (the kernels use 16-byte vector / TMA accesses on whole tensors)
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 63) // 64 * 64
        self.numel = off
        self.flat = torch.zeros(self.numel, device=dev, dtype=dt)
        self.average = average
        self.overlap = overlap
        self._handles = []
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)
        if overlap:
            # registered unconditionally: the process group may be initialised after this object is built
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._hook)

    @staticmethod
    def _distributed() -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _hook(self, p: torch.nn.Parameter):
        if p.grad is None or not self._distributed():
            return
        if self.average:
            p.grad.mul_(1.0 / dist.get_world_size())
        self._handles.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, async_op=True))

    def zero_(self):
        self.flat.zero_()

    def rebind(self):
        """Re-attach the views if something replaced ``.grad`` (e.g. ``zero_grad(set_to_none=True)``)."""
        for p, off in zip(self.params, self.offsets):
            view = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view

    def all_reduce(self, async_op: bool = False):
        if not self._distributed():
            return None
        if self.overlap and self._handles:
            for h in self._handles:
                h.wait()
            self._handles.clear()
            return None
        # no per-parameter collectives were queued (overlap off, or the hooks did not fire): one flat all-reduce
        if self.average:
            self.flat.mul_(1.0 / dist.get_world_size())
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)


class FlatAdam:
    """Adam (``torch.optim.Adam`` semantics: no amsgrad, no weight decay) as ONE kernel over flat buffers
    (``spn_adam_step``): the parameters are re-homed as views into one flat fp32 buffer next to the reducer's flat
    gradient buffer, so the update of the whole model is a single 28-byte-per-parameter HBM stream.

    The kernel writes the parameters behind autograd's version counters, so every step bumps ``ops.PARAM_EPOCH``, which
    is part of the key of all derived-buffer caches (bf16 operand images of the tensor-core path, int32 permutations):
    the next forward rebuilds them.  ``module`` is accepted for backward compatibility and additionally frees the stale
    images eagerly.  The step count (bias correction) lives on the host, so a step must not be captured into a CUDA
    graph; use ``torch.optim.Adam(capturable=True)`` there."""

    def __init__(self, reducer: FlatGradAllReducer, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 module: torch.nn.Module = None):
        self.reducer = reducer
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.flat_p = torch.zeros_like(reducer.flat)
        for p, off in zip(reducer.params, reducer.offsets):
            n = p.numel()
            self.flat_p[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + n].view_as(p)
        self.exp_avg = torch.zeros_like(reducer.flat)
        self.exp_avg_sq = torch.zeros_like(reducer.flat)
        self.t = 0
        self._cached = [m for m in module.modules() if hasattr(m, "invalidate_weight_cache")] if module is not None else []

    def step(self):
        from . import _abi
        self.t += 1
        b1, b2 = self.betas
        bc1, bc2 = 1.0 - b1 ** self.t, 1.0 - b2 ** self.t
        _abi.call("spn_adam_step", _abi.ptr(self.flat_p), _abi.ptr(self.reducer.flat), _abi.ptr(self.exp_avg),
                  _abi.ptr(self.exp_avg_sq), self.flat_p.numel(), self.lr / bc1, b1, b2, self.eps, 1.0 / bc2 ** 0.5,
                  _abi.stream_ptr(self.flat_p.device))
        from . import ops
        ops.bump_param_epoch()
        for m in self._cached:
            m.invalidate_weight_cache()


def broadcast_parameters(module: torch.nn.Module, src: int = 0):
    """Make every rank start from rank ``src``'s parameters and buffers (incl. the feature permutation)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)
    from . import ops
    ops.bump_param_epoch()  # .data writes bypass the version counters the derived-buffer caches are keyed on
    for m in module.modules():
        if hasattr(m, "invalidate_weight_cache"):
            m.invalidate_weight_cache()
        if hasattr(m, "_perm_cache"):
            m._perm_cache = None
