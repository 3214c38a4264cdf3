"""Scratch: per-object phase timeline from the MRPNP_TRACE build (clock64 ticks of warp A, lane 0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monorun_b200 import _native
_native.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libmonorun_pnp_trace.so')
from monorun_b200 import synth, pnp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
weights = sys.argv[2] if len(sys.argv) > 2 else 'diag'
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
    p = pnp.make_params(n, 784, weight_mode=2 if full else 0, precision=1)
    f = lambda x, ty='float*': _native.ffi.cast(ty, x.data_ptr())
    _native.check(_native.lib().mrpnp_solve(ctx.ptr, p, f(c3), f(c2), f(w), f(cam), f(uvr), f(init), _native.ffi.NULL,
                                            f(result), _native.ffi.NULL, f(tr, 'double*'), _native.ffi.NULL))
    torch.cuda.synchronize()
tr = tr.cpu().numpy(); res = result.cpu().numpy()
evals = None
names = ['load+cam', 'sweepA(S1)', 'masks(S2)', 'compact(S3)']
tr[:, 3] = np.where(tr[:, 3] > 0, tr[:, 3], tr[:, 2])
d = np.diff(tr[:, 0:5], axis=1)
print('objects', n, 'mean LM iters', res[:, 21].mean())
for i, nm in enumerate(names):
    print(f'{nm:14s} mean {d[:, i].mean():9.0f} ticks  p50 {np.median(d[:, i]):9.0f}')
tot = tr[:, 30] - tr[:, 0]
print(f'total/object   mean {tot.mean():9.0f} p50 {np.median(tot):9.0f}   LM part {np.mean(tr[:,29]-tr[:,4]):9.0f}  epilogue {np.mean(tr[:,30]-tr[:,29]):9.0f}')
for e in range(4):
    ok = (tr[:, 7 + 4 * e] > 0) | (tr[:, 6 + 4 * e] > 0)
    if ok.sum() == 0: break
    a = tr[ok]
    if a[:, 4 + 4 * e].min() == 0 or a[:, 7 + 4 * e].min() == 0:  # single-warp kernel: only 5+4e (pass start), 6+4e (pass end)
        ok = tr[:, 6 + 4 * e] > 0; a = tr[ok]
        pa = a[:, 6 + 4 * e] - a[:, 5 + 4 * e]
        nxt = np.where(a[:, 9 + 4 * e] > 0, a[:, 9 + 4 * e], a[:, 29]) - a[:, 6 + 4 * e]
        print(f'eval {e}: n={ok.sum():5d} pass {pa.mean():7.0f}  scalar-to-next-pass {nxt.mean():7.0f}')
        continue
    pub = a[:, 5 + 4 * e] - a[:, 4 + 4 * e]
    pa = a[:, 6 + 4 * e] - a[:, 5 + 4 * e]
    wb = a[:, 7 + 4 * e] - a[:, 6 + 4 * e]
    nxt = np.where(a[:, 8 + 4 * e] > 0, a[:, 8 + 4 * e], a[:, 29]) - a[:, 7 + 4 * e]
    print(f'eval {e}: n={ok.sum():5d} publish+BAR1 {pub.mean():7.0f}  passA {pa.mean():7.0f}  wait-B(BAR2) {wb.mean():7.0f}  scalar-to-next {nxt.mean():7.0f}')
span = tr[:, 30].max() - tr[:, 0].min()
print('kernel span ticks (max end - min start, per-SM clocks differ):', span)
