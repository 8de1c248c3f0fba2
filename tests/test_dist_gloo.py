"""CPU, world_size = 2 over gloo: the data-parallel plumbing (batch sharding, parameter broadcast, one flat gradient
all-reduce).  The arithmetic kernels are CUDA-only, so a stand-in differentiable function is used for the gradients;
what is tested here is exactly the host logic that bench.py / a trainer uses around the kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from ratspn_b200.parallel import FlatGradAllReducer, broadcast_parameters, shard_batch, shard_bounds
    torch.manual_seed(100 + rank)  # ranks start from different parameters on purpose
    lin = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    broadcast_parameters(lin)
    red = FlatGradAllReducer(lin.parameters())
    full = torch.arange(7 * 6, dtype=torch.float32).view(7, 6) / 10.0  # 7 rows: uneven split 4 + 3
    mine = shard_batch(full, 0)
    lo, hi = shard_bounds(7, world, rank)
    assert mine.shape[0] == hi - lo and torch.equal(mine, full[lo:hi])
    red.zero_()
    (lin(mine).sum() / full.shape[0]).backward()  # global-mean loss: per-rank partial sums
    red.average = False
    red.all_reduce()
    # reference: single process on the full batch
    torch.manual_seed(100)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    ref.load_state_dict(lin.state_dict())
    (ref(full).sum() / full.shape[0]).backward()
    ok = all(torch.allclose(p.grad, q.grad, atol=1e-6) for p, q in zip(lin.parameters(), ref.parameters()))
    # grads are views into the flat buffer (single collective, no copies)
    views = all(p.grad.data_ptr() >= red.flat.data_ptr() for p in lin.parameters())
    # zero_grad(set_to_none=True) followed by rebind keeps the single-buffer property
    lin.zero_grad(set_to_none=True)
    red.rebind()
    rebound = all(p.grad is not None for p in lin.parameters())
    # overlapped mode: per-parameter async all-reduce from post-accumulate-grad hooks, waited for in all_reduce()
    torch.manual_seed(100)
    lin2 = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    lin2.load_state_dict(ref.state_dict())
    red2 = FlatGradAllReducer(lin2.parameters(), average=False, overlap=True)
    red2.zero_()
    (lin2(mine).sum() / full.shape[0]).backward()
    red2.all_reduce()
    overlap_ok = all(torch.allclose(p.grad, q.grad, atol=1e-6) for p, q in zip(lin2.parameters(), ref.parameters()))
    ret[rank] = bool(ok and views and rebound and overlap_ok and red.numel == sum((p.numel() + 63) // 64 * 64 for p in lin.parameters()))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_grad_allreduce_world_size_2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_shard_bounds_cover_everything_once():
    from ratspn_b200.parallel import shard_bounds
    for n in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
