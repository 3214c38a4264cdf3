"""Scratch: a few launches of the solver for ncu (config via argv: n precision weights mode)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monorun_b200 import synth, pnp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
prec = sys.argv[2] if len(sys.argv) > 2 else 'fp64'
weights = sys.argv[3] if len(sys.argv) > 3 else 'diag'
mode = sys.argv[4] if len(sys.argv) > 4 else 'S1'
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
bands = tuple(float(v) for v in sys.argv[6].split(':')) if len(sys.argv) > 6 else None
b = synth.make_batch(n, config=3 if weights == 'full' else 2, weights=weights, mode=mode)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
full = weights == 'full'
c3, c2 = t(b['coords_3d']), t(b['coords_2d'])
w = t(b['w_full']) if full else t(b['logstd'])
cam = t(b['cam_mat'][None]); ih, iw = b['img_shape']
uvr = torch.tensor([[-200., iw + 200., -200., ih + 200.]], device='cuda')
init = t(b['init_pose'])
for _ in range(reps):
    res, _, _ = pnp.solve_batched(c3, c2, w, cam, uvr, init_pose=init, layout='planar',
                                  weight_mode='full' if full else 'logstd', precision=prec, return_inlier_mask=False, decision_bands=bands)
torch.cuda.synchronize()
print('valid', res[:, 20].mean().item(), 'iters', res[:, 21].mean().item())
