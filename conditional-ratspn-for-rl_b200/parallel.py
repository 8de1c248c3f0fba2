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
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=dt)
        self.average = average
        self.overlap = overlap
        self._handles = []
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        if overlap and self._distributed():
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._hook)

    @staticmethod
    def _distributed() -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _hook(self, p: torch.nn.Parameter):
        if p.grad is None:
            return
        if self.average:
            p.grad.mul_(1.0 / dist.get_world_size())
        self._handles.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, async_op=True))

    def zero_(self):
        self.flat.zero_()

    def rebind(self):
        """Re-attach the views if something replaced ``.grad`` (e.g. ``zero_grad(set_to_none=True)``)."""
        off = 0
        for p in self.params:
            view = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view
            off += p.numel()

    def all_reduce(self, async_op: bool = False):
        if not self._distributed():
            return None
        if self.overlap:
            for h in self._handles:
                h.wait()
            self._handles.clear()
            return None
        if self.average:
            self.flat.mul_(1.0 / dist.get_world_size())
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)


def broadcast_parameters(module: torch.nn.Module, src: int = 0):
    """Make every rank start from rank ``src``'s parameters and buffers (incl. the feature permutation)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)
    for m in module.modules():  # .data writes bypass the version counter the operand-image cache is keyed on
        if hasattr(m, "invalidate_weight_cache"):
            m.invalidate_weight_cache()
