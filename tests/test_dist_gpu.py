"""Multi-GPU path on real devices (-m gpu, needs >= 2 GPUs; skipped on a single-GPU box): the fused all-gather
(peer-to-peer stores from the solver's epilogue into every rank's symmetric buffer, completion signalled by flags the
solver raises in every rank's memory -- or by the symmetric-memory barrier) must return exactly the rows of the NCCL
all-gather, also when the buffers are re-used by solves of different problems.  The host-side sharding logic is covered on CPU by tests/test_host.py (gloo, world 2)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n_total, out_dir):
    import torch.distributed as dist
    from monorun_b200 import dist as mdist, pnp, synth
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', device_id=dev)
    start, stop = mdist.shard_range(n_total, rank, world)
    b = synth.make_batch(n_total, config=3, weights='full', mode='S1')   # same seed on every rank; each takes its shard
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[start:stop])).to(dev)
    ih, iw = b['img_shape']
    args = (t(b['coords_3d']), t(b['coords_2d']), t(b['w_full']), torch.from_numpy(b['cam_mat'][None]).to(dev),
            torch.tensor([[-200., iw + 200., -200., ih + 200.]], device=dev))
    kw = dict(init_pose=t(b['init_pose']), layout='planar', weight_mode='full', return_inlier_mask=False)
    rows, _, _ = pnp.solve_batched(*args, **kw)
    ref = mdist.all_gather_rows(rows, n_total)
    kw_b = dict(kw, init_pose=kw['init_pose'] * 1.01)           # a second problem: other rows in the same buffers
    ref_b = mdist.all_gather_rows(pnp.solve_batched(*args, **kw_b)[0], n_total)
    assert not torch.equal(ref, ref_b)
    # round-1 protocol: symmetric-memory barrier after the launch
    fg = mdist.FusedGather(n_total, dev, signal='barrier')
    for _ in range(3):   # re-use of the same buffer across solves
        res, _, _ = pnp.solve_batched(*args, **kw, **fg.solve_kwargs())
        assert res is None
        got = fg.finish().clone()
        fg.finish()      # nobody starts the next solve before everyone has copied the rows out
    ok_barrier = torch.equal(got, ref)
    # completion flags raised by the solver's last thread block; the consumer waits, reads, releases
    fg = mdist.FusedGather(n_total, dev)
    ok_flags = True
    for i in range(6):
        pnp.solve_batched(*args, **(kw_b if i % 2 else kw), **fg.solve_kwargs())
        got = fg.finish(release=False).clone()
        fg.release()
        ok_flags = ok_flags and torch.equal(got, ref_b if i % 2 else ref)
    # two buffers in rotation, the wait lagging one solve behind the launch (no rank waits for another's current solve)
    ring = [mdist.FusedGather(n_total, dev) for _ in range(2)]
    outs = []
    for i in range(7):
        pnp.solve_batched(*args, **(kw_b if i % 2 else kw), **ring[i % 2].solve_kwargs())
        if i >= 1:
            outs.append((i - 1, ring[(i - 1) % 2].finish(release=False).clone()))
            ring[(i - 1) % 2].release()
    outs.append((6, ring[0].finish().clone()))
    ok_ring = all(torch.equal(g, ref_b if i % 2 else ref) for i, g in outs) and len(outs) == 7
    ok_flags = ok_flags and ok_ring and pnp.gather_timeouts(dev) == 0 and ok_barrier
    got = outs[-1][1]
    torch.cuda.synchronize()
    ok = ok_flags and torch.equal(got, ref) and bool((got[:, 20] == 1).all())
    open(os.path.join(out_dir, f'ok{rank}'), 'w').write('1' if ok else '0')
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_fused_gather_equals_nccl_all_gather(cuda_lib, tmp_path):
    import torch.multiprocessing as mp
    world, n_total = 2, 1001    # uneven shards: 501 + 500
    mp.spawn(_worker, args=(world, 29533, n_total, str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f'ok{r}').read() for r in range(world)] == ['1'] * world
