"""GPU diagnostic: for each decision-band setting, over several seeded data sets of both bench workloads: objects handed
to the exact routine, objects whose evaluation count differs from the fp64 kernel's (which reproduces the oracle), objects
outside the north-star tolerance, and the serialised kernel time.
    python tools/band_sweep.py [n] "b1:b2:b3[:ratio:min],..." [ranks]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monorun_b200 import synth, pnp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
bands = [tuple(float(v) for v in b.split(':')) for b in sys.argv[2].split(',')]
ranks = [int(r) for r in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0, 1, 2, 3]
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
tot = {b: dict(handed=0, diff=0, off=0, us=[]) for b in bands}
for cfg, weights in ((3, 'full'), (2, 'diag')):
    for rank in ranks:
        b = synth.make_batch(n, config=cfg, rank=rank, weights=weights, mode='S1', classes=(0, 1, 2) if weights == 'full' else (0,))
        full = weights == 'full'
        ih, iw = b['img_shape']
        rng = torch.tensor([[-200., iw + 200., -200., ih + 200.]], device='cuda')
        args = (t(b['coords_3d']), t(b['coords_2d']), t(b['w_full'] if full else b['logstd']), t(b['cam_mat'][None]), rng)
        kw = dict(init_pose=t(b['init_pose']), layout='planar', weight_mode='full' if full else 'logstd')
        _, _, ref = pnp.solve_batched(*args, precision='fp64', return_fp64=True, return_inlier_mask=False, **kw)
        ref = ref.cpu().numpy()
        for bd in bands:
            hb0 = pnp.handed_back_count()
            _, _, r = pnp.solve_batched(*args, precision='fast', decision_bands=bd, return_fp64=True, return_inlier_mask=False, **kw)
            hb = pnp.handed_back_count() - hb0
            r = r.cpu().numpy()
            terr = np.linalg.norm(r[:, 1:4] - ref[:, 1:4], axis=1) / np.linalg.norm(ref[:, 1:4], axis=1)
            yerr = np.abs((r[:, 0] - ref[:, 0] + np.pi) % (2 * np.pi) - np.pi)
            diff = int((r[:, 6] != ref[:, 6]).sum()); off = int(((terr >= 1e-4) | (yerr >= 1e-3)).sum())
            for _ in range(3):
                pnp.solve_batched(*args, precision='fast', decision_bands=bd, return_inlier_mask=False, **kw)
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(4)]
            for a_, b_ in ev:
                a_.record(); pnp.solve_batched(*args, precision='fast', decision_bands=bd, return_inlier_mask=False, **kw); b_.record()
            torch.cuda.synchronize()
            us = float(np.mean([1e3 * a_.elapsed_time(b_) for a_, b_ in ev]))
            print(json.dumps(dict(workload=weights, rank=rank, bands=bd, handed=hb, diff_evals=diff, off_tolerance=off, us=round(us, 1))), flush=True)
            d = tot[bd]; d['handed'] += hb; d['diff'] += diff; d['off'] += off; d['us'].append(us)
for bd, d in tot.items():
    print('TOTAL', bd, 'handed', d['handed'], 'diff_evals', d['diff'], 'off_tolerance', d['off'], 'mean us', round(float(np.mean(d['us'])), 1),
          'max us', round(float(np.max(d['us'])), 1), flush=True)
