"""GPU diagnostic: which objects does the fp32 path hand to the exact routine, why, after how many evaluations, and how
many evaluations does their exact solve take?   python tools/hand_back_report.py [n] [ranks] [bands]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monorun_b200 import synth, pnp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ranks = [int(r) for r in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 1, 2, 3]
bands = tuple(float(v) for v in sys.argv[3].split(':')) if len(sys.argv) > 3 else None
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
why = {1: 'clip', 2: 'first', 3: 'ftol', 4: 'accept'}
for cfg, weights in ((3, 'full'), (2, 'diag')):
    for rank in ranks:
        b = synth.make_batch(n, config=cfg, rank=rank, weights=weights, mode='S1', classes=(0, 1, 2) if weights == 'full' else (0,))
        full = weights == 'full'
        ih, iw = b['img_shape']
        rng = torch.tensor([[-200., iw + 200., -200., ih + 200.]], device='cuda')
        args = (t(b['coords_3d']), t(b['coords_2d']), t(b['w_full'] if full else b['logstd']), t(b['cam_mat'][None]), rng)
        log = torch.zeros(n, dtype=torch.int32, device='cuda')
        _, _, r = pnp.solve_batched(*args, init_pose=t(b['init_pose']), layout='planar', weight_mode='full' if full else 'logstd',
                                    precision='fast', return_fp64=True, return_inlier_mask=False, hand_back_log=log, decision_bands=bands)
        log = log.cpu().numpy(); r = r.cpu().numpy()
        idx = np.nonzero(log)[0]
        items = sorted(((why[int(log[i]) & 255], int(log[i]) >> 8, int(r[i, 6])) for i in idx), key=lambda x: -x[2])
        print(weights, rank, 'handed', len(idx), '(reason, evaluations at hand-back, evaluations of the exact solve):', items, flush=True)
