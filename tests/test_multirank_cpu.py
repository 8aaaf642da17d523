"""N>1 host logic on CPU (gloo, world_size 2): batch sharding + the single all-gather of alpha reproduce the single-rank
result exactly (samples are independent, so the shards concatenate bit for bit)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_alpha(lo, hi, R):
    # stands in for the engine: a deterministic per-sample function of the GLOBAL sample index
    out = torch.empty(hi - lo, R, R, dtype=torch.float16)
    for i, g in enumerate(range(lo, hi)):
        gen = torch.Generator().manual_seed(g)
        out[i] = torch.rand(R, R, generator=gen).half()
    return out


def _worker(rank, world, port, total, R, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bench import shard_range

    lo, hi = shard_range(total, rank, world)
    mine = _fake_alpha(lo, hi, R)
    gathered = torch.empty(total, R, R, dtype=torch.float16)
    dist.all_gather_into_tensor(gathered, mine)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max-over-ranks timing reduction used by bench.py
    if rank == 0:
        q.put((gathered, t.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_allgather_world2():
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    total, R, world = 8, 16, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, R, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(gathered, _fake_alpha(0, total, R))
    assert tmax == 2.0


def test_shard_range_partitions():
    from bench import shard_range

    for world in (1, 2, 4, 8):
        spans = [shard_range(64, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 64
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
