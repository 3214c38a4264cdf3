"""GPU diagnostic: serialized kernel time and hand-back count of the fast kernel as a function of the decision bands."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monorun_b200 import synth, pnp


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    bands = [tuple(float(v) for v in b.split(':')) for b in (sys.argv[2] if len(sys.argv) > 2 else '0:0:0,8e-6:4e-3:2e-6').split(',')]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for cfg, weights in ((3, 'full'), (2, 'diag')):
        sets = []
        for rank in (0, 16):
            b = synth.make_batch(n, config=cfg, rank=rank, weights=weights, mode='S1', classes=(0, 1, 2) if weights == 'full' else (0,))
            ih, iw = b['img_shape']
            sets.append((t(b['coords_3d']), t(b['coords_2d']), t(b['w_full'] if weights == 'full' else b['logstd']), t(b['cam_mat'][None]),
                         torch.tensor([[-200., iw + 200., -200., ih + 200.]], device='cuda'), t(b['init_pose'])))
        others = [] if 'fastonly' in sys.argv else [('mixed', None), ('fp64', None)]
        for prec, bd in [('fast', b) for b in bands] + others:
            kw = dict(layout='planar', weight_mode='full' if weights == 'full' else 'logstd', precision=prec, return_inlier_mask=False,
                      decision_bands=bd)
            for i in range(4):
                s = sets[i % 2]
                pnp.solve_batched(*s[:5], init_pose=s[5], **kw)
            torch.cuda.synchronize()
            try:
                hb0 = pnp.handed_back_count()
            except Exception:
                hb0 = None
            reps = 20 if prec == 'fast' else 4
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
            for i in range(reps):
                s = sets[i % 2]
                ev[i][0].record()
                pnp.solve_batched(*s[:5], init_pose=s[5], **kw)
                ev[i][1].record()
            torch.cuda.synchronize()
            ms = [a.elapsed_time(b) for a, b in ev]
            hb = (pnp.handed_back_count() - hb0) if hb0 is not None else 0
            print(json.dumps({'workload': weights, 'precision': prec, 'bands': bd, 'us_mean': 1e3 * float(np.mean(ms)), 'us_min': 1e3 * float(np.min(ms)),
                              'us_max': 1e3 * float(np.max(ms)), 'handed_back_per_launch': hb / reps}), flush=True)


if __name__ == '__main__':
    main()
