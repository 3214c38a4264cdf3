"""oracle/pnp_driver.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python side of the CPU oracle: binds ``oracle/libpnp_oracle.so`` (cffi ABI mode) and restates the
reference's per-object driver so parity tests read like the reference's own call chain:

* ``u2d_pnp_cpu``         <- monorun/ops/least_squares/pnp_uncert_cpu.py:128-209
* ``u2d_pnp_cpu_single``  <- monorun/ops/least_squares/pnp_uncert_cpu.py:11-125
* ``pnp_uncert_ref``      <- monorun/ops/least_squares/pnp_uncert.py:7-87 (numpy fp32 in/out, covariance
                             through the restated approx_hessian, hessian.py:67-87)

No output of the reference binary exists to pin against; the minimiser is pinned to Ceres' own published tutorial
tables instead (see the header of pnp_oracle.cpp).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import this module.
"""
import os
import subprocess

import numpy as np
from cffi import FFI

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libpnp_oracle.so')

_CDEF = """
void pnp_uncert(double* pts2d, double* pts3d, double* wgt2d, double* K, double* init_pose,
                int* result_val, double* result_pose, double* result_cov, double* result_tr,
                int pn, double* clips);
void pnp_uncert_fullw(double* pts2d, double* pts3d, double* wgt2d, double* K, double* init_pose,
                int* result_val, double* result_pose, double* result_cov, double* result_tr,
                int pn, double* clips);
void pnp_uncert_batch(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K,
                const double* init_pose, int* result_val, double* result_pose, double* result_cov,
                double* result_tr, const int* pn, const long long* off, const double* clips,
                int nb, int full_w, int* stats, double* cost, int threads);
void pnp_eval(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K,
              const double* pose, int pn, const double* clips, int full_w, double* cost, double* grad,
              double* JtJ);
void pnp_approx_hessian(const double* pts2d, const double* pts3d, const double* istd, const double* K,
                        const double* pose, const unsigned char* inlier, int pn, const double* clips,
                        double* H);
int pnp_spd_inverse4(const double* H, double* inv);
int ceres_kat_hello_world(double* trace, int max_rows, double* x_final, double* summary);
int ceres_kat_powell(double* trace, int max_rows, double* x_final, double* summary);
void ceres_lm_expfit(const double* t, const double* y, int m, double* x, double* summary);
void pnp_oracle_set_adopt_candidate_on_ftol(int v);
int pnp_oracle_num_threads(void);
"""

ffi = FFI()
ffi.cdef(_CDEF)
_lib = None


def build(force=False):
    """Compile the oracle with its Makefile (g++ -O2 -fopenmp, the reference's own flags)."""
    src = os.path.join(_HERE, 'pnp_oracle.cpp')
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, 'ceres_lm.h'))):
        subprocess.check_call(['make', '-C', _HERE, '-s'] + (['-B'] if force else []))
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


# --------------------------------------------------------------------------------------
# per-object native call: the cffi sequence of pnp_uncert_cpu.py:70-117
# --------------------------------------------------------------------------------------
def lm_single(coord_2d, coord_3d, wgt, cam_mat, init_pose, clips, with_pose_cov=True, full_w=False):
    """One native solve.  coord_2d (n,2), coord_3d (n,3), wgt (n,2|3), cam_mat (3,3), init_pose (4,),
    clips (5,) = z_min,u_min,u_max,v_min,v_max.  Returns (val, pose[4], cov[4,4]|None, tr_radius)."""
    coord_2d, coord_3d, wgt = _c64(coord_2d), _c64(coord_3d), _c64(wgt)
    cam_mat, init_pose, clips = _c64(cam_mat), _c64(init_pose), _c64(clips)
    result_val = np.zeros([1], np.int32)
    result_pose = np.zeros([4], np.float64)
    result_cov = np.eye(4, dtype=np.float64) if with_pose_cov else None
    result_tr = np.zeros([1], np.float64)
    fn = lib().pnp_uncert_fullw if full_w else lib().pnp_uncert
    fn(_dp(coord_2d), _dp(coord_3d), _dp(wgt), _dp(cam_mat), _dp(init_pose),
       ffi.cast('int*', result_val.ctypes.data), _dp(result_pose),
       _dp(result_cov) if with_pose_cov else ffi.NULL, _dp(result_tr),
       coord_2d.shape[0], _dp(clips))
    return result_val[0] > 0, result_pose, result_cov, result_tr[0]


def pack_lm_batch(coords_2d, coords_3d, wgt, cam_mats, init_pose, clips, inlier_mask=None, full_w=False):
    """The fp64 buffers ``pnp_uncert_batch`` reads: inliers compacted per object exactly like
    pnp_uncert_cpu.py:24-27,62-66 (boolean-mask indexing keeps point order), K and clips broadcast.  Separate from
    the solve so that a timing loop (bench.py's CPU legs) can pack once and time the native calls only."""
    n, p = coords_2d.shape[:2]
    wc = 3 if full_w else 2
    if inlier_mask is None:
        inlier_mask = np.ones((n, p), bool)
    inlier_mask = np.asarray(inlier_mask, bool)
    pn = inlier_mask.sum(1).astype(np.int32)
    off = np.zeros(n, np.int64)
    off[1:] = np.cumsum(pn[:-1])
    flat = inlier_mask.reshape(-1)
    return dict(
        n=n, full_w=bool(full_w), pn=pn, off=off,
        p2=_c64(np.asarray(coords_2d).reshape(-1, 2)[flat]), p3=_c64(np.asarray(coords_3d).reshape(-1, 3)[flat]),
        w=_c64(np.asarray(wgt).reshape(-1, wc)[flat]),
        k=_c64(np.broadcast_to(np.asarray(cam_mats, np.float64).reshape(-1, 9), (n, 9))),
        cl=_c64(np.broadcast_to(np.asarray(clips, np.float64).reshape(-1, 5), (n, 5))), init=_c64(init_pose))


def lm_batch_packed(pk, with_pose_cov=False, threads=1):
    """``pnp_uncert_batch`` on the buffers of :func:`pack_lm_batch`.  Returns dict(val, pose, cov, tr,
    stats[N,4]=(iters, cost evals, jac evals, termination), cost)."""
    n = pk['n']
    val = np.zeros(n, np.int32)
    pose = np.zeros((n, 4), np.float64)
    cov = np.tile(np.eye(4), (n, 1, 1)) if with_pose_cov else None
    tr = np.zeros(n, np.float64)
    stats = np.zeros((n, 4), np.int32)
    cost = np.zeros(n, np.float64)
    lib().pnp_uncert_batch(
        _dp(pk['p2']), _dp(pk['p3']), _dp(pk['w']), _dp(pk['k']), _dp(pk['init']), ffi.cast('int*', val.ctypes.data),
        _dp(pose), _dp(cov) if with_pose_cov else ffi.NULL, _dp(tr), ffi.cast('int*', pk['pn'].ctypes.data),
        ffi.cast('long long*', pk['off'].ctypes.data), _dp(pk['cl']), n, int(pk['full_w']),
        ffi.cast('int*', stats.ctypes.data), _dp(cost), int(threads))
    return dict(val=val > 0, pose=pose, cov=cov, tr=tr, stats=stats, cost=cost)


def lm_batch(coords_2d, coords_3d, wgt, cam_mats, init_pose, clips, inlier_mask=None, full_w=False,
             with_pose_cov=False, threads=1):
    """Batched LM from a given init and a given inlier mask (the LM-parity contract).

    coords_2d (N,P,2), coords_3d (N,P,3), wgt (N,P,2|3), cam_mats (N|1,3,3), init_pose (N,4),
    clips (N|1,5), inlier_mask (N,P) bool or None.
    Returns dict(val, pose, cov, tr, stats[N,4]=(iters, cost evals, jac evals, termination), cost).
    """
    return lm_batch_packed(pack_lm_batch(coords_2d, coords_3d, wgt, cam_mats, init_pose, clips, inlier_mask, full_w),
                           with_pose_cov=with_pose_cov, threads=threads)


def ceres_tutorial_trace(problem):
    """Runs the oracle's minimiser (the same template that solves the PnP problems) on a Ceres tutorial problem:
    'hello_world' (examples/helloworld.cc) or 'powell' (examples/powell.cc).  Returns (rows [k,7] = iteration, cost,
    cost_change, |gradient|, |step|, tr_ratio, tr_radius; x_final; (termination, iterations, final |gradient|))."""
    fn, n = {'hello_world': (lib().ceres_kat_hello_world, 1), 'powell': (lib().ceres_kat_powell, 4)}[problem]
    trace, x, summary = np.zeros((64, 7)), np.zeros(n), np.zeros(3)
    k = fn(_dp(trace), 64, _dp(x), _dp(summary))
    return trace[:k], x, summary


def expfit(t, y, x0):
    """The oracle's minimiser on r_i = x0 exp(x1 t_i) + x2 + x3 t_i - y_i.  Returns (x[4], (termination, iterations,
    cost evaluations, final cost))."""
    t, y, x = _c64(t), _c64(y), _c64(x0).copy()
    summary = np.zeros(4)
    lib().ceres_lm_expfit(_dp(t), _dp(y), t.shape[0], _dp(x), _dp(summary))
    return x, summary


def eval_cost_grad_hess(coord_2d, coord_3d, wgt, cam_mat, pose, clips, full_w=False):
    coord_2d, coord_3d, wgt = _c64(coord_2d), _c64(coord_3d), _c64(wgt)
    cam_mat, pose, clips = _c64(cam_mat), _c64(pose), _c64(clips)
    cost = np.zeros(1)
    grad = np.zeros(4)
    jtj = np.zeros((4, 4))
    lib().pnp_eval(_dp(coord_2d), _dp(coord_3d), _dp(wgt), _dp(cam_mat), _dp(pose), coord_2d.shape[0],
                   _dp(clips), int(full_w), _dp(cost), _dp(grad), _dp(jtj))
    return cost[0], grad, jtj


def approx_hessian(coords_2d, coords_2d_istd, coords_3d, cam_mats, u_range, v_range, z_min, yaw, t_vec,
                   inlier_mask):
    """fp64 restatement of hessian.py:67-87 (same argument order). Returns H (N,4,4)."""
    n = coords_2d.shape[0]
    cam = np.broadcast_to(np.asarray(cam_mats, np.float64), (n, 3, 3))
    ur = np.broadcast_to(np.asarray(u_range, np.float64), (n, 2))
    vr = np.broadcast_to(np.asarray(v_range, np.float64), (n, 2))
    h = np.zeros((n, 4, 4))
    for b in range(n):
        p2, p3, w = _c64(coords_2d[b]), _c64(coords_3d[b]), _c64(coords_2d_istd[b])
        k = _c64(cam[b])
        pose = _c64(np.concatenate([np.ravel(yaw[b]), np.ravel(t_vec[b])]))
        clips = _c64([z_min, ur[b, 0], ur[b, 1], vr[b, 0], vr[b, 1]])
        hb = np.zeros((4, 4))
        if inlier_mask is not None:
            m = np.ascontiguousarray(inlier_mask[b], np.uint8)
            mp = ffi.cast('unsigned char*', m.ctypes.data)
        else:
            mp = ffi.NULL
        lib().pnp_approx_hessian(_dp(p2), _dp(p3), _dp(w), _dp(k), _dp(pose), mp, p2.shape[0], _dp(clips),
                                 _dp(hb))
        h[b] = hb
    return h


def exact_hessian(coords_2d, coords_2d_istd, coords_3d, cam_mats, u_range, v_range, z_min, yaw, t_vec,
                  inlier_mask):
    """fp64 restatement of hessian.py:5-64 (same argument order). Returns H (N,4,4).

    The reference differentiates g = J^T e with autograd, J and e from jacobian.py:4-98 & :157-184.  Rows in
    ``zero_mask`` (z-clipped point, own coordinate clipped, outlier; jacobian.py:52-59) have a constant zero
    Jacobian and drop out of g; on every other row J is the true derivative of e (nothing it depends on is
    clamped), so H = sum_rows J^T J + e * Hess(e) with the unclipped projection's second derivatives:
    for f = x'/z' (or y'/z'):  f_pq = (x'_pq - f z'_pq - f_q z'_p - f_p z'_q) / z'.
    Pinned to the reference's own autograd result by tests/golden/exact_hessian_ref.npz."""
    c2, w, c3 = _c64(coords_2d), _c64(coords_2d_istd), _c64(coords_3d)
    n = c2.shape[0]
    cam = np.broadcast_to(np.asarray(cam_mats, np.float64), (n, 3, 3))
    ur = np.broadcast_to(np.asarray(u_range, np.float64), (n, 2))[:, None, :]
    vr = np.broadcast_to(np.asarray(v_range, np.float64), (n, 2))[:, None, :]
    yaw = np.asarray(yaw, np.float64).reshape(n, 1)
    t = np.asarray(t_vec, np.float64).reshape(n, 1, 3)
    fx, fy, cx, cy = (cam[:, i, j][:, None] for i, j in ((0, 0), (1, 1), (0, 2), (1, 2)))
    sn, cs = np.sin(yaw), np.cos(yaw)
    qx = cs * c3[..., 0] + sn * c3[..., 2]
    qz = -sn * c3[..., 0] + cs * c3[..., 2]
    xc, yc, zc = qx + t[..., 0], c3[..., 1] + t[..., 1], qz + t[..., 2]
    z_free = ~(zc < z_min)                                            # jacobian.py:28
    iz = 1.0 / np.where(z_free, zc, z_min)
    zero = np.zeros_like(iz)
    h = np.zeros((n, 4, 4))
    inl = np.ones_like(z_free) if inlier_mask is None else np.asarray(inlier_mask, bool)
    for f, num_p, focal, centre, rng_, obs, wgt in (
            (xc * iz, (qz, 1.0 + zero, zero, zero), fx, cx, ur, c2[..., 0], w[..., 0]),
            (yc * iz, (zero, zero, 1.0 + zero, zero), fy, cy, vr, c2[..., 1], w[..., 1])):
        proj = focal * f + centre
        free = z_free & inl & ~((proj < rng_[..., 0]) | (proj > rng_[..., 1]))   # jacobian.py:38-40, :52-59
        e = wgt * (proj - obs)                                        # jacobian.py:181
        zp = (-qx, zero, zero, 1.0 + zero)                            # d z' / d(yaw, tx, ty, tz)
        fp = [(num_p[k] - f * zp[k]) * iz for k in range(4)]
        for a in range(4):
            for b in range(4):
                num_pq = -qx if (a == 0 and b == 0 and num_p[0] is qz) else zero   # d2 x'/dyaw2 = -qx; y' is linear
                z_pq = -qz if (a == 0 and b == 0) else zero                       # d2 z'/dyaw2 = -qz
                f_pq = (num_pq - f * z_pq - fp[b] * zp[a] - fp[a] * zp[b]) * iz
                term = (wgt * focal) ** 2 * fp[a] * fp[b] + e * wgt * focal * f_pq
                h[:, a, b] += np.where(free, term, 0.0).sum(1)
    return h


# --------------------------------------------------------------------------------------
# restated reference driver
# --------------------------------------------------------------------------------------
def u2d_pnp_cpu_single(coord_2d, coord_2d_istd, coord_3d, istd_inlier_mask, cam_mat, u_range, v_range,
                       epnp_ransac_thres, inlier_opt_only=False, z_min=0.5, dist_coeffs=None,
                       with_pose_cov=True, init_pose=None):
    """pnp_uncert_cpu.py:11-125.  ``init_pose`` (4,) is an extension used by the LM-parity tests: when
    given, the OpenCV EPnP(+RANSAC) initialisation of :34-58 is skipped and that pose seeds LM."""
    import cv2
    istd_inlier_mask = istd_inlier_mask.copy()
    istd_inlier_count = np.count_nonzero(istd_inlier_mask)
    if istd_inlier_count > 4:                                   # :23-27
        coord_3d_inlier = coord_3d[istd_inlier_mask]
        coord_2d_inlier = coord_2d[istd_inlier_mask]
        coord_2d_istd_inlier = coord_2d_istd[istd_inlier_mask]
    else:                                                       # :28-32
        coord_3d_inlier, coord_2d_inlier, coord_2d_istd_inlier = coord_3d, coord_2d, coord_2d_istd
        istd_inlier_mask[:] = True

    if init_pose is not None:
        ret_val, r_vec, t_vec = True, np.array([[0.], [init_pose[0]], [0.]]), np.asarray(init_pose[1:]).reshape(3, 1)
    elif epnp_ransac_thres is not None:                         # :34-51
        ret_val, r_vec, t_vec, ransac_inlier_ind = cv2.solvePnPRansac(
            coord_3d_inlier, coord_2d_inlier, cam_mat, dist_coeffs,
            reprojectionError=float(epnp_ransac_thres), iterationsCount=30, flags=cv2.SOLVEPNP_EPNP)
        if ransac_inlier_ind is not None and len(ransac_inlier_ind) > 4:
            ransac_inlier_ind = ransac_inlier_ind.squeeze(1)
            ransac_inlier_mask = np.zeros(coord_3d_inlier.shape[0], dtype=bool)
            ransac_inlier_mask[ransac_inlier_ind] = True
            coord_3d_inlier = coord_3d_inlier[ransac_inlier_ind]
            coord_2d_inlier = coord_2d_inlier[ransac_inlier_ind]
            coord_2d_istd_inlier = coord_2d_istd_inlier[ransac_inlier_ind]
            istd_inlier_mask[istd_inlier_mask] = ransac_inlier_mask
    else:                                                       # :53-58
        ret_val, r_vec, t_vec = cv2.solvePnP(coord_3d_inlier, coord_2d_inlier, cam_mat, dist_coeffs,
                                             flags=cv2.SOLVEPNP_EPNP)
    inlier_mask = istd_inlier_mask

    if ret_val:
        if inlier_opt_only:                                     # :62-66
            coord_3d, coord_2d, coord_2d_istd = coord_3d_inlier, coord_2d_inlier, coord_2d_istd_inlier
        yaw = np.asarray(r_vec).reshape(3)[1:2]                 # :68
        clips = np.array([z_min, u_range[0], u_range[1], v_range[0], v_range[1]], np.float64)
        init = np.concatenate([yaw, np.asarray(t_vec).reshape(3)], axis=0)
        val, pose, cov, tr = lm_single(coord_2d, coord_3d, coord_2d_istd, cam_mat, init, clips,
                                       with_pose_cov=with_pose_cov)
        return (val, pose[0:1].astype(np.float32), pose[1:].astype(np.float32),   # :108-117
                cov.astype(np.float32) if cov is not None else None,
                np.array([tr], np.float32), inlier_mask)
    return (False, np.zeros(1, np.float32), np.zeros(3, np.float32),              # :119-125
            np.eye(4, dtype=np.float32) if with_pose_cov else None, np.zeros(1, np.float32), inlier_mask)


def istd_inlier_masks(coords_2d_istd, epnp_istd_thres):
    """pnp_uncert_cpu.py:164-168 -- numpy fp32 mean over points, both axes must pass."""
    mean = np.mean(coords_2d_istd, axis=1, keepdims=True)
    return np.min(coords_2d_istd >= epnp_istd_thres * mean, axis=2)


def u2d_pnp_cpu(coords_2d, coords_2d_istd, coords_3d, cam_mats, u_range, v_range, z_min=0.5,
                epnp_istd_thres=1.0, epnp_ransac_thres=None, inlier_opt_only=False, with_pose_cov=True,
                init_pose=None):
    """pnp_uncert_cpu.py:128-209 (multi_apply == map + transpose)."""
    bn, pn = coords_2d.shape[0], coords_2d.shape[1]
    if bn == 0:                                                 # :201-207
        return (np.zeros((0,), bool), np.zeros((0, 1), np.float32), np.zeros((0, 3), np.float32),
                np.zeros((0, 4, 4), np.float32), np.zeros((0, 1), np.float32), np.zeros((0, pn), bool))
    assert coords_2d_istd.shape[1] == coords_3d.shape[1] == pn >= 4
    masks = istd_inlier_masks(coords_2d_istd, epnp_istd_thres)
    cam_mats = np.broadcast_to(cam_mats, (bn, 3, 3))
    u_range = np.broadcast_to(u_range, (bn, 2))
    v_range = np.broadcast_to(v_range, (bn, 2))
    thres = [None] * bn if epnp_ransac_thres is None else epnp_ransac_thres
    dist_coeffs = np.zeros((8, 1), dtype=np.float32)
    outs = [u2d_pnp_cpu_single(coords_2d[b], coords_2d_istd[b], coords_3d[b], masks[b], cam_mats[b],
                               u_range[b], v_range[b], thres[b], inlier_opt_only=inlier_opt_only,
                               z_min=z_min, dist_coeffs=dist_coeffs, with_pose_cov=with_pose_cov,
                               init_pose=None if init_pose is None else init_pose[b])
            for b in range(bn)]
    ret_val, yaw, t_vec, pose_cov, tr_radius, inlier_mask = zip(*outs)
    return (np.array(ret_val, dtype=bool), np.stack(yaw, 0), np.stack(t_vec, 0),
            np.stack(pose_cov, 0) if with_pose_cov else None, np.stack(tr_radius, 0), np.stack(inlier_mask, 0))


def pnp_uncert_ref(coords_2d, coords_2d_istd, coords_3d, cam_mats, u_range, v_range, z_min=0.5,
                   epnp_istd_thres=1.0, epnp_ransac_thres=None, inlier_opt_only=False, init_pose=None):
    """pnp_uncert.py:7-87 on numpy arrays: solve, then pose_cov = inverse(approx_hessian) with the
    eigenvalue fallback of :79-85.  Returns (ret_val, r_vec, t_vec, pose_cov, inlier_mask)."""
    ret_val, r_vec, t_vec, _, _, inlier_mask = u2d_pnp_cpu(
        coords_2d, coords_2d_istd, coords_3d, cam_mats, u_range, v_range, z_min=z_min,
        epnp_istd_thres=epnp_istd_thres, epnp_ransac_thres=epnp_ransac_thres,
        inlier_opt_only=inlier_opt_only, with_pose_cov=False, init_pose=init_pose)
    if ret_val.shape[0] == 0:
        return ret_val, r_vec, t_vec, np.zeros((0, 4, 4), np.float32), inlier_mask
    h = approx_hessian(coords_2d, coords_2d_istd, coords_3d, cam_mats, u_range, v_range, z_min, r_vec, t_vec,
                       inlier_mask)
    ret_val = ret_val.copy()
    try:
        pose_cov = np.linalg.inv(h)
    except np.linalg.LinAlgError:
        eigval = np.linalg.eigvalsh(h)
        valid = eigval[:, 0] > np.clip(1e-6 * eigval[:, 3], 0, None)
        ret_val &= valid
        h[~ret_val] = np.eye(4)
        pose_cov = np.linalg.inv(h)
    return ret_val, r_vec, t_vec, pose_cov.astype(np.float32), inlier_mask
