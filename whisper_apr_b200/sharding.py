"""Multi-GPU chunk sharding: the B200 successor of src/parallel.rs (rayon `parallel_map` over independent items).

Audio chunks are independent units (no cross-chunk state in mel or encoder; the max-8 clamp is per chunk), so the
path shards by batch with replicated weights and NO collective inside the layer stack.  One process per GPU; chunk i of a
global batch of n goes to rank floor(i * world / n) (contiguous blocks, earlier ranks take the remainder).  The only
communication is the optional final gather of encoder states (`gather_states`), an all_gather over NCCL/NVLink (gloo on CPU
in the tests).
"""
from __future__ import annotations


def shard_range(n_chunks: int, world: int, rank: int) -> tuple[int, int]:
    """[start, end) of the chunks rank `rank` owns; sizes differ by at most one, order preserved (parallel_map keeps order,
    src/parallel.rs:176-188)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n_chunks, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def owner_of(chunk: int, n_chunks: int, world: int) -> int:
    for r in range(world):
        s, e = shard_range(n_chunks, world, r)
        if s <= chunk < e:
            return r
    raise ValueError("chunk out of range")


def gather_states(local, n_chunks: int, group=None):
    """All-gather per-rank encoder states [n_local][S][d] (torch tensor) into [n_chunks][S][d] on every rank, in chunk order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if n_chunks % world == 0 and local.shape[0] == n_chunks // world and local.is_contiguous():
        # equal shards: one collective straight into the result, no staging copies
        out = torch.empty((n_chunks,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    sizes = [shard_range(n_chunks, world, r) for r in range(world)]
    mx = max(e - s for s, e in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: e - s] for r, (s, e) in enumerate(sizes)], dim=0)
