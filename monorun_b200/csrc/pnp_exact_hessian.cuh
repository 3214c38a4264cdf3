// pnp_exact_hessian.cuh -- PnPUncert(forward_exact_hessian=True): the second-order pose covariance of
// monorun/ops/least_squares/hessian.py:5-64, which the reference obtains by differentiating g = J^T e with
// autograd (J, e from jacobian.py:4-98, :157-184) and inverts (pnp_uncert.py:63-85).
//
// Rows in jacobian.py's zero_mask (z-clipped point, own coordinate clipped, outlier; :52-59) have a constant zero
// Jacobian and drop out of g.  On every other row nothing J or e depends on is clamped, so the autograd result is
//     H = sum_rows  J^T J + e * Hess(e),     e_u = w_u (fx x'/z' + cx - u),  e_v = w_v (fy y'/z' + cy - v)
// with, for f = x'/z' or y'/z' and parameters (yaw, tx, ty, tz):
//     f_p  = (num_p - f z'_p) / z'
//     f_pq = (num_pq - f z'_pq - f_q z'_p - f_p z'_q) / z'
//     x'_p = (qz, 1, 0, 0), y'_p = (0, 0, 1, 0), z'_p = (-qx, 0, 0, 1), x'_yaw,yaw = -qx, z'_yaw,yaw = -qz.
// One warp per object, fp64, one pass over the points at the final pose; the 4x4 inverse (Gauss-Jordan with
// partial pivoting: the exact Hessian need not be positive definite) is written into the result row.
// The point functor and the inverse are __host__ __device__ (tests/harness/ runs them under g++).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define MRXH_HD __host__ __device__ __forceinline__
#else
#define MRXH_HD inline
#endif

namespace mrxh {

struct Camera { double fx, fy, cx, cy, z_min, u_min, u_max, v_min, v_max; };

// acc[10]: upper triangle of H, row-major (00 01 02 03 11 12 13 22 23 33).
MRXH_HD void add_row(const double* fp, const double* zp, double f, double num_yy, double z_yy, double iz,
                     double wf /* weight * focal */, double e, double* acc) {
    int k = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a; b < 4; ++b) {
            const double num_pq = (a == 0 && b == 0) ? num_yy : 0.0;
            const double z_pq = (a == 0 && b == 0) ? z_yy : 0.0;
            const double f_pq = (num_pq - f * z_pq - fp[b] * zp[a] - fp[a] * zp[b]) * iz;
            acc[k++] += wf * wf * fp[a] * fp[b] + e * wf * f_pq;
        }
}

MRXH_HD void add_point(const Camera& cam, double sn, double cs, const double* t, double X, double Y, double Z,
                       double u_obs, double v_obs, double wu, double wv, double* acc) {
    const double qx = cs * X + sn * Z, qz = -sn * X + cs * Z;
    const double xc = qx + t[0], yc = Y + t[1], zc = qz + t[2];
    if (zc < cam.z_min) return;                                   // jacobian.py:28, :52-59: both rows masked
    const double iz = 1.0 / zc;
    const double zp[4] = {-qx, 0.0, 0.0, 1.0};
    {
        const double f = xc * iz, u = cam.fx * f + cam.cx;
        if (!(u < cam.u_min || u > cam.u_max)) {                  // jacobian.py:38-40
            const double fp[4] = {(qz + f * qx) * iz, iz, 0.0, -f * iz};
            add_row(fp, zp, f, -qx, -qz, iz, wu * cam.fx, wu * (u - u_obs), acc);
        }
    }
    {
        const double f = yc * iz, v = cam.fy * f + cam.cy;
        if (!(v < cam.v_min || v > cam.v_max)) {
            const double fp[4] = {f * qx * iz, 0.0, iz, -f * iz};
            add_row(fp, zp, f, 0.0, -qz, iz, wv * cam.fy, wv * (v - v_obs), acc);
        }
    }
}

// inv = H^-1 for a symmetric 4x4 given by its upper triangle; false if a pivot vanishes or is not finite
// (torch.inverse raising in pnp_uncert.py:78-79).
MRXH_HD bool invert4(const double* tri, double* inv) {
    double a[4][8];
    int k = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j) { a[i][j] = tri[k]; a[j][i] = tri[k]; ++k; }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) a[i][4 + j] = (i == j) ? 1.0 : 0.0;
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r) if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
        const double d = a[piv][c];
        if (!(fabs(d) > 0.0) || !isfinite(d)) return false;
        for (int j = 0; j < 8; ++j) { const double tmp = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = tmp; }
        const double id = 1.0 / d;
        for (int j = 0; j < 8; ++j) a[c][j] *= id;
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            const double m = a[r][c];
            for (int j = 0; j < 8; ++j) a[r][j] -= m * a[c][j];
        }
    }
    bool ok = true;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { inv[i * 4 + j] = a[i][4 + j]; ok = ok && isfinite(a[i][4 + j]); }
    return ok;
}

struct KParams {
    const float *coords_3d, *coords_2d, *weights, *cam_mats, *uv_range, *pose;
    const uint32_t* inlier;
    float* hessian;  // [N,16] or NULL
    float* rows;     // [N,24] or NULL
    int n_obj, n_pts, planar, logstd, cam_stride, range_stride, pose_stride;
    double z_min, std_scale;
};

#ifdef __CUDACC__

constexpr int kWarpsPerCta = 4;

__global__ void __launch_bounds__(kWarpsPerCta * 32) exact_hessian_kernel(const KParams kp) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * kWarpsPerCta;
    for (int obj = warp; obj < kp.n_obj; obj += n_warps) {
        const float* K = kp.cam_mats + (size_t)obj * kp.cam_stride;
        const float* rg = kp.uv_range + (size_t)obj * kp.range_stride;
        Camera cam;
        cam.fx = K[0]; cam.fy = K[4]; cam.cx = K[2]; cam.cy = K[5];
        cam.z_min = kp.z_min;
        cam.u_min = rg[0]; cam.u_max = rg[1]; cam.v_min = rg[2]; cam.v_max = rg[3];
        const float* ps = kp.pose + (size_t)obj * kp.pose_stride;
        const double yaw = ps[0], t[3] = {ps[1], ps[2], ps[3]};
        const double sn = sin(yaw), cs = cos(yaw);
        const float* c3 = kp.coords_3d + (size_t)obj * 3 * kp.n_pts;
        const float* c2 = kp.coords_2d + (size_t)obj * 2 * kp.n_pts;
        const float* cw = kp.weights + (size_t)obj * 2 * kp.n_pts;
        const uint32_t* mask = kp.inlier ? kp.inlier + (size_t)obj * ((kp.n_pts + 31) >> 5) : nullptr;
        auto at = [&](const float* base, int c, int nc, int p) {
            return kp.planar ? __ldg(base + (size_t)c * kp.n_pts + p) : __ldg(base + (size_t)p * nc + c);
        };
        double acc[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) acc[i] = 0.0;
        for (int base = 0; base < kp.n_pts; base += 32) {
            const int p = base + lane;
            bool on = p < kp.n_pts;
            if (mask) on = on && ((__ldg(mask + (base >> 5)) >> lane) & 1u);
            if (on) {
                double wu = at(cw, 0, 2, p), wv = at(cw, 1, 2, p);
                if (kp.logstd) { wu = exp(-wu) / kp.std_scale; wv = exp(-wv) / kp.std_scale; }  // uncert_prop_pnp_optimizer.py:73
                add_point(cam, sn, cs, t, at(c3, 0, 3, p), at(c3, 1, 3, p), at(c3, 2, 3, p), at(c2, 0, 2, p),
                          at(c2, 1, 2, p), wu, wv, acc);
            }
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            double v = acc[i];
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
            acc[i] = v;
        }
        if (lane == 0) {
            if (kp.hessian) {
                float* h = kp.hessian + (size_t)obj * 16;
                int k = 0;
                for (int i = 0; i < 4; ++i)
                    for (int j = i; j < 4; ++j) { h[i * 4 + j] = (float)acc[k]; h[j * 4 + i] = (float)acc[k]; ++k; }
            }
            if (kp.rows) {
                float* row = kp.rows + (size_t)obj * 24;
                double inv[16];
                const bool ok = invert4(acc, inv);
                for (int i = 0; i < 16; ++i) row[4 + i] = ok ? (float)inv[i] : ((i % 5 == 0) ? 1.f : 0.f);
                if (!ok) row[20] = 0.f;  // pnp_uncert.py:80-85: not invertible -> invalid, identity covariance
            }
        }
    }
}

#endif  // __CUDACC__

}  // namespace mrxh
