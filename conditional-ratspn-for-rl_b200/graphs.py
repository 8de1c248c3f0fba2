"""CUDA-graph capture of launch-bound call patterns (a CSPN policy update is several hundred small
kernels -- the gating MLP, the evidence pass, top-down sampling, the recursive entropy and their backward -- and is
bound by launch overhead, not by the device).  Everything on the hot path is capture-safe: the C ABI never allocates or
synchronises, kernels take the caller's stream, and the in-kernel noise of the sampling kernels is keyed by a device
word refreshed inside the graph (ops.new_seed), so every replay draws fresh samples.
"""
import gc
from typing import Callable

import torch


class GraphedCall:
    """``g = GraphedCall(fn)`` runs ``fn()`` a few times on a side stream, captures one call into a CUDA graph and
    replays it on ``g()``; the return value of the captured call (static tensors, overwritten by each replay) is
    ``g.outputs``.

    Rules (those of torch.cuda.graph): ``fn`` must read its inputs from tensors that stay allocated -- copy each new batch
    into them with ``copy_`` before ``g()``; shapes are fixed; no host synchronisation (``.item()``, ``.cpu()``, printing a
    tensor) inside ``fn``; gradients must either exist before capture and be zeroed in place inside ``fn`` or be freed
    (``set_to_none``) before construction; and no autograd graph built on the default stream may still be alive.
    An optimiser step may be part of ``fn`` when the optimiser is capturable (``torch.optim.Adam(..., capturable=True)``).
    """

    def __init__(self, fn: Callable, warmup: int = 3):
        gc.collect()  # autograd graphs kept alive only by reference cycles would pin default-stream AccumulateGrad nodes
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn()

    def __call__(self):
        self.graph.replay()
        return self.outputs
