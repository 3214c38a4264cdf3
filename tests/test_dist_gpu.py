"""Multi-GPU path on real devices (-m gpu, needs >= 2 GPUs; skipped on a single-GPU box): the fused all-gather
(peer-to-peer stores from the solver's epilogue into every rank's symmetric buffer + barrier) must return exactly the
rows of the NCCL all-gather.  The host-side sharding logic is covered on CPU by tests/test_host.py (gloo, world 2)."""
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
    fg = mdist.FusedGather(n_total, dev)
    for _ in range(3):   # re-use of the same buffer across solves
        res, _, _ = pnp.solve_batched(*args, **kw, **fg.solve_kwargs())
        assert res is None
        got = fg.finish().clone()
        fg.finish()      # nobody starts the next solve before everyone has copied the rows out
    torch.cuda.synchronize()
    ok = torch.equal(got, ref) and bool((got[:, 20] == 1).all())
    open(os.path.join(out_dir, f'ok{rank}'), 'w').write('1' if ok else '0')
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_fused_gather_equals_nccl_all_gather(cuda_lib, tmp_path):
    import torch.multiprocessing as mp
    world, n_total = 2, 1001    # uneven shards: 501 + 500
    mp.spawn(_worker, args=(world, 29533, n_total, str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f'ok{r}').read() for r in range(world)] == ['1'] * world
