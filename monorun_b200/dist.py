"""Multi-GPU layer: objects are independent, so the batch is cut into contiguous ranges, one per rank
(one process per GPU), each rank solves its range with the CUDA kernel, and the fixed-width result rows
([N_local, 24] fp32, 96 B per object) are made available everywhere -- either by ONE NCCL all-gather
(``all_gather_rows``) or, fused into the solver, by peer-to-peer stores from the kernel's epilogue into every rank's
symmetric buffer followed by a cross-rank barrier (``FusedGather``; 7 us instead of ~20 us on NVSwitch).
The reference has no multi-GPU inference at all (test.py:74-75: "multi-gpu testing is not yet supported").
"""
import torch
import torch.distributed as dist

from .pnp import RESULT_STRIDE

MAX_PEERS = 8   # MRPNP_MAX_PEERS


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


class FusedGather:
    """All-gather of the result rows fused into the solver kernel.

    Every rank owns a symmetric-memory buffer ``rows [n_total, 24]`` (torch.distributed._symmetric_memory: the same
    allocation mapped into every peer over NVLink / NVSwitch).  ``solve_batched(..., **fg.solve_kwargs())`` makes the
    kernel store the row of local object i into ALL ranks' buffers at row ``start + i`` -- 96-byte peer-to-peer stores
    from the epilogue of each object, no separate collective launch, no send / receive staging.

    Completion is signalled by the solver itself (``signal='flags'``, the default): the last thread block of each
    rank's launch raises that rank's slot in every rank's flag array (a second, tiny symmetric allocation), and
    ``fg.finish()`` is a one-warp kernel on the caller's stream that waits until all ranks' slots have reached the
    current solve number -- no barrier, nothing that makes a rank's NEXT solve wait for the slowest rank's current one;
    call it when the rows are needed, e.g. a few solves later when several FusedGather instances rotate.
    ``fg.release()`` (or ``finish(release=True)``) tells the peers that this rank has read the rows: the next solve of
    any rank into this buffer waits for every rank's release first (write-after-read), normally for zero time.
    ``signal='barrier'`` is the round-1 protocol: ``finish()`` runs the symmetric-memory barrier.
    """

    def __init__(self, n_total, device, group=None, signal='flags'):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        self.n_total = int(n_total)
        self.start, self.stop = shard_range(n_total, self.rank, self.world)
        self.rows = symm.empty((self.n_total, RESULT_STRIDE), dtype=torch.float32, device=self.device)
        self.rows.zero_()
        self.handle = symm.rendezvous(self.rows, group.group_name)
        self.peers = [int(p) for p in self.handle.buffer_ptrs]
        self.signal = signal
        self.epoch = 0        # solves issued into this buffer
        self.waited = 0       # last solve number finish() waited for
        if signal == 'flags':
            # [0, world): completion flags (slot r raised by rank r's launch); [world, 2 world): releases of the readers
            self.sig = symm.empty((2 * MAX_PEERS,), dtype=torch.int32, device=self.device)
            self.sig.zero_()
            self.sig_handle = symm.rendezvous(self.sig, group.group_name)
            self.sig_ptrs = [int(p) for p in self.sig_handle.buffer_ptrs]
            torch.cuda.synchronize(self.device)
            self.sig_handle.barrier()   # every rank's arrays are zero before anybody raises a flag
            torch.cuda.synchronize(self.device)

    def solve_kwargs(self):
        """Keyword arguments of ONE ``solve_batched`` call into this buffer (advances the solve number)."""
        kw = dict(peers=self.peers, row_offset=self.start)
        if self.signal == 'flags':
            self.epoch += 1
            kw.update(peer_flags=self.sig_ptrs, flag_slot=self.rank, flag_value=self.epoch,
                      acks=self.sig_ptrs[self.rank] + 4 * MAX_PEERS, ack_value=self.epoch - 1)
        return kw

    def finish(self, release=True):
        """On the current stream: afterwards every rank's rows of the latest solve are in ``rows``.  ``release=True``
        also tells the peers that the rows may be overwritten by the next solve -- pass False and call ``release()``
        after the consumer when something reads ``rows`` on this stream."""
        if self.signal != 'flags':
            self.handle.barrier()
            return self.rows
        from . import pnp
        acks = [p + 4 * MAX_PEERS for p in self.sig_ptrs] if release else None
        pnp.gather_wait(self.device, self.sig_ptrs[self.rank], self.world, self.epoch, acks, self.rank, self.epoch)
        self.waited = self.epoch
        return self.rows

    def release(self):
        from . import pnp
        if self.signal == 'flags':
            pnp.gather_wait(self.device, None, self.world, 0, [p + 4 * MAX_PEERS for p in self.sig_ptrs], self.rank, self.waited)


def solve_sharded(solve_fn, n_total, group=None):
    """``solve_fn(start, stop) -> [stop-start, 24]`` result rows for this rank's range; returns all rows."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    start, stop = shard_range(n_total, rank, world)
    local = solve_fn(start, stop)
    assert local.shape == (stop - start, RESULT_STRIDE)
    return all_gather_rows(local, n_total, group)
