"""Multi-GPU layer: objects are independent, so the batch is cut into contiguous ranges, one per rank
(one process per GPU), each rank solves its range with the CUDA kernel, and ONE all-gather of the fixed-width
result rows ([N_local, 24] fp32, 96 B per object) over NCCL/NVLink makes every pose available everywhere.
The reference has no multi-GPU inference at all (test.py:74-75: "multi-gpu testing is not yet supported").
"""
import torch
import torch.distributed as dist

from .pnp import RESULT_STRIDE


def shard_range(n_total, rank, world_size):
    """Contiguous range [start, stop) of rank `rank`; the first n_total % world_size ranks get one extra object."""
    base, rem = divmod(int(n_total), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_counts(n_total, world_size):
    return [shard_range(n_total, r, world_size)[1] - shard_range(n_total, r, world_size)[0]
            for r in range(world_size)]


def all_gather_rows(local_rows, n_total, group=None):
    """All-gather of per-rank result rows into the global [n_total, C] tensor (rank order == object order).

    Uses a single ``all_gather_into_tensor`` on the current stream; ranks whose shard is one row short are
    padded to the common size and the padding is dropped after the collective.
    """
    world = dist.get_world_size(group)
    if world == 1:
        return local_rows
    counts = shard_counts(n_total, world)
    width = local_rows.shape[1]
    m = max(counts)
    send = local_rows
    if local_rows.shape[0] != m:
        send = local_rows.new_zeros((m, width))
        send[:local_rows.shape[0]] = local_rows
    out = local_rows.new_empty((world * m, width))
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    if all(c == m for c in counts):
        return out
    return torch.cat([out[r * m:r * m + c] for r, c in enumerate(counts)], dim=0)


def solve_sharded(solve_fn, n_total, group=None):
    """``solve_fn(start, stop) -> [stop-start, 24]`` result rows for this rank's range; returns all rows."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    start, stop = shard_range(n_total, rank, world)
    local = solve_fn(start, stop)
    assert local.shape == (stop - start, RESULT_STRIDE)
    return all_gather_rows(local, n_total, group)
