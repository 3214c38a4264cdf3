"""oracle/noc_driver.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Binds ``oracle/libpnp_noc_oracle.so``: the CPU restatement of the reference's 7-parameter solvers
``pnp_noc_uncert`` / ``pnp_noc_cov_uncert`` (monorun/ops/least_squares/src/ext.h:15-43,
pnp_uncert_cpu.cpp:294-377).  The reference has no Python caller for them, so there is no driver to
restate; the functions below only marshal numpy buffers the way pnp_uncert_cpu.py:70-117 does for the
4-parameter op.  PARITY UNPINNED (see the header of pnp_noc_oracle.cpp).  Only tests/ may import this.
"""
import os
import subprocess

import numpy as np
from cffi import FFI

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libpnp_noc_oracle.so')

ffi = FFI()
ffi.cdef("""
void pnp_noc_uncert(double* pts2d, double* pts3d, double* wgt2d, double* logdim, double* logdim_wgt, double* K,
                    double* init_dimpose, int* result_val, double* result_dimpose, int pn, double* clips,
                    double delta);
void pnp_noc_cov_uncert(double* pts2d, double* pts3d, double* wgt2d, double* logdim, double* logdim_wgt,
                        double* K, double* init_dimpose, int* result_val, double* result_dimpose, int pn,
                        double* clips, double delta);
void pnp_noc_batch(const double* pts2d, const double* pts3d, const double* wgt2d, const double* logdim,
                   const double* logdim_wgt, const double* K, const double* init_dimpose, int* result_val,
                   double* result_dimpose, const int* pn, const long long* off, const double* clips, double delta,
                   int nb, int full_w, int* stats, double* cost, int threads);
void pnp_noc_eval(const double* pts2d, const double* pts3d, const double* wgt2d, const double* logdim,
                  const double* logdim_wgt, const double* K, const double* dimpose, int pn, const double* clips,
                  double delta, int full_w, double* cost, double* grad, double* JtJ);
""")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, 'pnp_noc_oracle.cpp')
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, 'ceres_lm.h'))):
        subprocess.check_call(['make', '-C', _HERE, '-s', 'libpnp_noc_oracle.so'] + (['-B'] if force else []))
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


def noc_single(coord_2d, coord_3d, wgt, logdim, logdim_wgt, cam_mat, init_dimpose, clips, delta, full_w=False):
    """One native call with the reference's argument list (ext.h:15-43).  Returns (val, dimpose[7])."""
    coord_2d, coord_3d, wgt = _c64(coord_2d), _c64(coord_3d), _c64(wgt)
    logdim, logdim_wgt, cam_mat = _c64(logdim), _c64(logdim_wgt), _c64(cam_mat)
    init_dimpose, clips = _c64(init_dimpose), _c64(clips)
    val = np.zeros(1, np.int32)
    out = np.zeros(7, np.float64)
    fn = lib().pnp_noc_cov_uncert if full_w else lib().pnp_noc_uncert
    fn(_dp(coord_2d), _dp(coord_3d), _dp(wgt), _dp(logdim), _dp(logdim_wgt), _dp(cam_mat), _dp(init_dimpose),
       ffi.cast('int*', val.ctypes.data), _dp(out), coord_2d.shape[0], _dp(clips), float(delta))
    return val[0] > 0, out


def noc_batch(coords_2d, coords_3d, wgt, logdim, logdim_wgt, cam_mats, init_dimpose, clips, delta,
              inlier_mask=None, full_w=False, threads=1):
    """coords_2d (N,P,2), coords_3d (N,P,3), wgt (N,P,2|3), logdim / logdim_wgt (N,3), cam_mats (N|1,3,3),
    init_dimpose (N,7), clips (N|1,5), inlier_mask (N,P) bool or None (points are compacted in order).
    Returns dict(val, dimpose, stats[N,4]=(iterations, cost evals, jacobian evals, termination), cost)."""
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
    ld, lw, init = _c64(logdim), _c64(logdim_wgt), _c64(init_dimpose)
    val = np.zeros(n, np.int32)
    out = np.zeros((n, 7), np.float64)
    stats = np.zeros((n, 4), np.int32)
    cost = np.zeros(n, np.float64)
    lib().pnp_noc_batch(_dp(p2), _dp(p3), _dp(w), _dp(ld), _dp(lw), _dp(k), _dp(init),
                        ffi.cast('int*', val.ctypes.data), _dp(out), ffi.cast('int*', pn.ctypes.data),
                        ffi.cast('long long*', off.ctypes.data), _dp(cl), float(delta), n, int(full_w),
                        ffi.cast('int*', stats.ctypes.data), _dp(cost), int(threads))
    return dict(val=val > 0, dimpose=out, stats=stats, cost=cost)


def noc_eval(coord_2d, coord_3d, wgt, logdim, logdim_wgt, cam_mat, dimpose, clips, delta, full_w=False):
    """Robustified cost, gradient [7] and J^T J [7,7] at ``dimpose`` (what the LM loop sees)."""
    coord_2d, coord_3d, wgt = _c64(coord_2d), _c64(coord_3d), _c64(wgt)
    logdim, logdim_wgt, cam_mat = _c64(logdim), _c64(logdim_wgt), _c64(cam_mat)
    dimpose, clips = _c64(dimpose), _c64(clips)
    cost, grad, jtj = np.zeros(1), np.zeros(7), np.zeros((7, 7))
    lib().pnp_noc_eval(_dp(coord_2d), _dp(coord_3d), _dp(wgt), _dp(logdim), _dp(logdim_wgt), _dp(cam_mat),
                       _dp(dimpose), coord_2d.shape[0], _dp(clips), float(delta), int(full_w), _dp(cost),
                       _dp(grad), _dp(jtj))
    return cost[0], grad, jtj
