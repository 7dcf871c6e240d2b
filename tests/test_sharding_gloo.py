"""Multi-rank host logic on CPU: world_size-2 gloo run of the chunk sharding + final gather (the N > 1 path of bench.py)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from whisper_apr_b200 import sharding


@pytest.mark.parametrize("n,world", [(256, 8), (7, 2), (5, 8), (1, 1), (0, 4), (33, 4)])
def test_shard_range_partitions_in_order(n, world):
    cover = []
    sizes = []
    for r in range(world):
        s, e = sharding.shard_range(n, world, r)
        cover += list(range(s, e))
        sizes.append(e - s)
    assert cover == list(range(n))
    assert max(sizes) - min(sizes) <= 1
    for c in range(n):
        s, e = sharding.shard_range(n, world, sharding.owner_of(c, n, world))
        assert s <= c < e


def _worker(rank, world, port, n_chunks, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = sharding.shard_range(n_chunks, world, rank)
    # each rank "encodes" its own chunks: state of chunk c is a deterministic function of c only
    local = torch.stack([torch.full((3, 4), float(c)) + torch.arange(4.0) for c in range(s, e)]) if e > s else torch.zeros((0, 3, 4))
    full = sharding.gather_states(local, n_chunks)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)           # the max-over-ranks timing reduction bench.py uses
    q.put((rank, full.numpy(), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_chunks", [5, 4])
def test_two_rank_gloo_shard_and_gather(n_chunks):
    world, port = 2, 29500 + os.getpid() % 1000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_chunks, q)) for r in range(world)]
    [p.start() for p in procs]
    results = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    expect = np.stack([np.full((3, 4), float(c)) + np.arange(4.0) for c in range(n_chunks)]).astype(np.float32)
    for rank, full, mx in results:
        assert np.array_equal(full, expect)            # every rank holds all states, in chunk order
        assert mx == 2.0
