"""Scratch: per-object phase timeline from the MRPNP_TRACE build (clock64 ticks of lane 0 of the object's warp).
usage: trace_run.py [n] [diag|full] [mixed|fast]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monorun_b200 import _native
_native.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libmonorun_pnp_trace.so')
from monorun_b200 import synth, pnp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
weights = sys.argv[2] if len(sys.argv) > 2 else 'diag'
prec = {'fp64': 0, 'mixed': 1, 'fast': 2}[sys.argv[3] if len(sys.argv) > 3 else 'fast']
b = synth.make_batch(n, config=3 if weights == 'full' else 2, weights=weights, mode='S1')
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
full = weights == 'full'
c3, c2, w = t(b['coords_3d']), t(b['coords_2d']), t(b['w_full'] if full else b['logstd'])
cam = t(b['cam_mat'][None]); ih, iw = b['img_shape']
uvr = torch.tensor([[-200., iw + 200., -200., ih + 200.]], device='cuda')
init = t(b['init_pose'])
ctx = pnp.get_ctx('cuda')
for rep in range(3):
    result = torch.empty((n, 24), device='cuda'); tr = torch.zeros((n, 32), dtype=torch.float64, device='cuda')
    p = pnp.make_params(n, 784, weight_mode=2 if full else 0, precision=prec)
    f = lambda x, ty='float*': _native.ffi.cast(ty, x.data_ptr())
    _native.check(_native.lib().mrpnp_solve(ctx.ptr, p, f(c3), f(c2), f(w), f(cam), f(uvr), f(init), _native.ffi.NULL,
                                            f(result), _native.ffi.NULL, f(tr, 'double*'), _native.ffi.NULL))
    torch.cuda.synchronize()
tr = tr.cpu().numpy(); res = result.cpu().numpy()
names = ['0 staging wait', '1 weights+mask+compact', '2 initialiser', '3 first evaluation', '4 candidate evals',
         '5 scalar TR algebra', '6 roll-backs', '7 covariance+stores']
ev = tr[:, 8]
print(f'objects {n} {weights} prec {prec}: mean LM iters {res[:, 21].mean():.2f}, mean evals {ev.mean():.2f}, mean inliers {tr[:, 10].mean():.0f}')
tot = tr[:, 9]
for i, nm in enumerate(names):
    print(f'{nm:24s} mean {tr[:, i].mean():9.0f} ticks ({100 * tr[:, i].sum() / tot.sum():5.1f} %)   p50 {np.median(tr[:, i]):9.0f}')
print(f'total per object         mean {tot.mean():9.0f}   p50 {np.median(tot):9.0f}   p99 {np.quantile(tot, 0.99):9.0f}  max {tot.max():9.0f}')
print(f'per candidate evaluation mean {tr[:, 4].sum() / np.maximum(ev - 1, 0).sum():9.0f}   scalar per evaluation {tr[:, 5].sum() / ev.sum():9.0f}')
# kernel span and tail: per warp busy time vs the span of the whole launch
t0, t1 = tr[:, 11].min(), tr[:, 12].max()
key = tr[:, 13] * 64 + tr[:, 14]
busy = np.array([tot[key == k].sum() for k in np.unique(key)])
print(f'kernel span {t1 - t0:.0f} ticks; per-warp busy mean {busy.mean():.0f} min {busy.min():.0f} max {busy.max():.0f}; warps {len(busy)}')
# regression: candidate-evaluation time = (evals - 1) * (a + b * rows); first evaluation = a1 + b1 * rows; prologue
rows = np.ceil(tr[:, 10] / 32.0)
ne = np.maximum(ev - 1, 0)
ok = ne > 0
A = np.stack([ne[ok], ne[ok] * rows[ok]], 1)
coef, *_ = np.linalg.lstsq(A, tr[ok, 4], rcond=None)
print(f'candidate eval: fixed {coef[0]:.0f} ticks + {coef[1]:.0f} ticks per row of 32 points')
A = np.stack([np.ones(n), rows], 1)
coef, *_ = np.linalg.lstsq(A, tr[:, 3], rcond=None)
print(f'first eval:     fixed {coef[0]:.0f} ticks + {coef[1]:.0f} ticks per row')
coef, *_ = np.linalg.lstsq(A, tr[:, 1], rcond=None)
print(f'prologue:       fixed {coef[0]:.0f} ticks + {coef[1]:.0f} ticks per inlier row')
A = np.stack([ev], 1)
coef, *_ = np.linalg.lstsq(A, tr[:, 5], rcond=None)
print(f'scalar:         {coef[0]:.0f} ticks per evaluation')
