"""oracle/sixdof_driver.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Binds ``oracle/libpnp_6dof_oracle.so``: the CPU statement of the 6-DoF extension of the uncertainty PnP (unknowns
[rvec(3), t(3)], ceres::AngleAxisRotatePoint, otherwise the reference's 4-DoF op).  The reference has no 6-DoF code
(``use_6dof`` is never read, pnp_uncert.py:11,98,122,142), so PARITY IS UNPINNED by construction -- see the header of
pnp_6dof_oracle.cpp.  Only tests/ may import this.
"""
import os
import subprocess

import numpy as np
from cffi import FFI

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libpnp_6dof_oracle.so')

ffi = FFI()
ffi.cdef("""
void pnp_6dof_batch(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K, const double* init,
                    int* result_val, double* result_pose, double* result_cov, const int* pn, const long long* off,
                    const double* clips, int nb, int full_w, int* stats, double* cost, int threads);
void pnp_6dof_eval(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K, const double* pose,
                   int pn, const double* clips, int full_w, double* cost, double* grad, double* JtJ);
""")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, 'pnp_6dof_oracle.cpp')
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, 'ceres_lm.h'))):
        subprocess.check_call(['make', '-C', _HERE, '-s', 'libpnp_6dof_oracle.so'] + (['-B'] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ffi.dlopen(_LIB_PATH)
    return _lib


def _dp(a):
    return ffi.cast('double*', a.ctypes.data)


def _c64(a):
    return np.ascontiguousarray(a, np.float64)


def solve_batch(coords_2d, coords_3d, wgt, cam_mats, init_pose, clips, inlier_mask=None, full_w=False,
                with_cov=True, threads=1):
    """coords_2d (N,P,2), coords_3d (N,P,3), wgt (N,P,2|3), cam_mats (N|1,3,3), init_pose (N,6) [rvec, t],
    clips (N|1,5), inlier_mask (N,P) bool or None.  Returns dict(val, pose[N,6], cov[N,6,6]|None, stats[N,4], cost)."""
    n, p = coords_2d.shape[:2]
    wc = 3 if full_w else 2
    if inlier_mask is None:
        inlier_mask = np.ones((n, p), bool)
    inlier_mask = np.asarray(inlier_mask, bool)
    pn = inlier_mask.sum(1).astype(np.int32)
    off = np.zeros(n, np.int64)
    off[1:] = np.cumsum(pn[:-1])
    flat = inlier_mask.reshape(-1)
    p2 = _c64(np.asarray(coords_2d).reshape(-1, 2)[flat])
    p3 = _c64(np.asarray(coords_3d).reshape(-1, 3)[flat])
    w = _c64(np.asarray(wgt).reshape(-1, wc)[flat])
    k = _c64(np.broadcast_to(np.asarray(cam_mats, np.float64).reshape(-1, 9), (n, 9)))
    cl = _c64(np.broadcast_to(np.asarray(clips, np.float64).reshape(-1, 5), (n, 5)))
    init = _c64(init_pose)
    val = np.zeros(n, np.int32)
    pose = np.zeros((n, 6))
    cov = np.tile(np.eye(6), (n, 1, 1)) if with_cov else None
    stats = np.zeros((n, 4), np.int32)
    cost = np.zeros(n)
    lib().pnp_6dof_batch(_dp(p2), _dp(p3), _dp(w), _dp(k), _dp(init), ffi.cast('int*', val.ctypes.data), _dp(pose),
                         _dp(cov) if with_cov else ffi.NULL, ffi.cast('int*', pn.ctypes.data),
                         ffi.cast('long long*', off.ctypes.data), _dp(cl), n, int(full_w),
                         ffi.cast('int*', stats.ctypes.data), _dp(cost), int(threads))
    return dict(val=val > 0, pose=pose, cov=cov, stats=stats, cost=cost)


def eval_cost_grad_hess(coord_2d, coord_3d, wgt, cam_mat, pose, clips, full_w=False):
    coord_2d, coord_3d, wgt = _c64(coord_2d), _c64(coord_3d), _c64(wgt)
    cam_mat, pose, clips = _c64(cam_mat), _c64(pose), _c64(clips)
    cost, grad, jtj = np.zeros(1), np.zeros(6), np.zeros((6, 6))
    lib().pnp_6dof_eval(_dp(coord_2d), _dp(coord_3d), _dp(wgt), _dp(cam_mat), _dp(pose), coord_2d.shape[0],
                        _dp(clips), int(full_w), _dp(cost), _dp(grad), _dp(jtj))
    return cost[0], grad, jtj
