"""Batch driver: a set of stereo pairs sharded over GPUs, one process per GPU, no collective
on the data path (every pair's pipeline touches only its own tensors; the
reference's own multi-GPU use is exactly this -- independent processes with `-gpu k`,
main.lua:16,342; rgs.py:9-14,85).

torch.distributed is plumbing only: rendezvous, the barrier around timed regions, the
max-over-ranks of the timings and (optionally) gathering the disparity maps on rank 0.
The compute callable is injected, so the host logic is testable on CPU with gloo.
"""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world):
    """Round-robin ownership: rank r processes items r, r + world, r + 2*world, ..."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, n_items, world))


def owner_of(index, world):
    return index % world


def max_over_ranks(value, device="cpu"):
    """max of a python float over all ranks (device-timed durations are reported as the max)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_sharded(n_items, compute, rank=None, world=None, gather=True, device="cpu"):
    """Run ``compute(i) -> tensor`` for the items this rank owns.

    Returns, on rank 0 (when ``gather``), the list of all results in item order; on other
    ranks (or without gather) the dict {item index: result} of the local shard.  Results of
    one item may have any shape but all ranks must agree on shapes per item index parity
    (equal shapes in practice: disparity maps of equal-size pairs).
    """
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    mine = shard_indices(n_items, rank, world)
    local = {i: compute(i) for i in mine}
    if not gather or world == 1:
        return [local[i] for i in range(n_items)] if world == 1 else local
    # gather round by round: in round k, rank r contributes item r + k*world (or nothing)
    out = [None] * n_items
    rounds = (n_items + world - 1) // world
    for k in range(rounds):
        idx = rank + k * world
        have = idx < n_items
        # fixed-length shape record [ndim, d0 .. d6]: every rank sends the same number of elements even when it owns no item
        # in this round (n_items < world) or items of different rank
        rec = [0] * 8
        if have:
            shp0 = list(local[idx].shape)
            assert len(shp0) <= 7, "run_sharded gathers tensors of at most 7 dimensions"
            rec[0] = len(shp0)
            rec[1:1 + len(shp0)] = shp0
        shape = torch.tensor(rec, dtype=torch.int64, device=device)
        shapes = [torch.zeros_like(shape) for _ in range(world)]
        dist.all_gather(shapes, shape)
        shapes = [s[1:1 + int(s[0])] for s in shapes]
        numel = max([int(torch.prod(s).item()) if s.numel() else 0 for s in shapes] + [1])
        buf = torch.zeros(numel, dtype=torch.float32, device=device)
        if have:
            buf[: local[idx].numel()] = local[idx].reshape(-1).to(device)
        bufs = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, bufs, dst=0)
        if rank == 0:
            for r in range(world):
                j = r + k * world
                if j < n_items:
                    shp = [int(v) for v in shapes[r].tolist()]
                    n = 1
                    for v in shp:
                        n *= v
                    out[j] = bufs[r][:n].reshape(shp).clone()
    return out if rank == 0 else local
