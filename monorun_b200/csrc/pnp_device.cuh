// pnp_device.cuh -- device-side building blocks of the batched uncertainty-PnP solver (sm_100a).
//
// One warp owns one object: its correspondence slab ([3|2|2or3] x P fp32, 22-25 KB at P=784) is staged
// into the warp's shared-memory slot with 1-D bulk TMA copies (cp.async.bulk + mbarrier), the istd
// inlier test compacts it in place, and every Levenberg-Marquardt pass streams the compacted points
// from shared memory, accumulating cost, J^T r and the upper triangle of J^T J in registers, followed by
// a warp-shuffle butterfly.  The 4x4 damped solve and the trust-region bookkeeping run redundantly on all
// lanes in fp64 (no divergence, no broadcast).
//
// Reference semantics being reproduced (see DESIGN.md section 2 for the full table):
//   residual + clips   monorun/ops/least_squares/src/pnp_uncert_cpu.cpp:24-51, :189-217
//   LM control flow    Ceres 1.14 TrustRegionMinimizer / LevenbergMarquardtStrategy defaults
//   inlier test        monorun/ops/least_squares/pnp_uncert_cpu.py:164-168, :23-32
//   covariance masks   monorun/ops/least_squares/jacobian.py:48-98
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "monorun_pnp.h"

namespace mrpnp {

constexpr int kMaxWarpsPerCta = 10;        // 10 x 21,952 B slots fill the 227 KB of one SM at P = 784
constexpr int kMaxThreads = kMaxWarpsPerCta * 32;
constexpr int kBarrierBytes = 128;          // one 8-byte mbarrier per warp, padded
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNumAcc = 15;                 // |r|^2, g[4], upper-tri(J^T J)[10]

struct KParams {
    const float* c3d;
    const float* c2d;
    const float* wgt;
    const float* cam;
    const float* range;
    const float* init;
    const uint8_t* inl_in;
    float* result;
    uint8_t* inl_out;
    double* result64;
    int* counters;  // [0] next object, [1] finished CTAs (self-resetting)
    int n_obj, n_pts, cam_stride, range_stride;
    int cov_mode, init_mode, inlier_opt_only, max_iter, adopt_ftol;
    int use_tma, slot_floats;
    float z_min, std_scale, istd_thres;
};

// ------------------------------------------------------------------ PTX wrappers (TMA bulk copy + mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// ------------------------------------------------------------------ shared-memory accessors
// Planar slot: channel c of point p at base[c * P + p]; interleaved slot: base[p * C + c].
template <int LAYOUT, int C>
__device__ __forceinline__ int sidx(int p, int c, int P) {
    return LAYOUT == MRPNP_LAYOUT_PLANAR ? c * P + p : p * C + c;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// fp64 reciprocal: MUFU.RCP64H seed (2^-23) + two Newton steps on the fp64 pipe (no slow-path branch).
__device__ __forceinline__ double fast_rcp(double z) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(z));
    double e = fma(-z, x, 1.0);
    x = fma(x, e, x);
    e = fma(-z, x, 1.0);
    x = fma(x, e, x);
    return x;
}

template <typename T>
struct Camera {
    T fx, fy, cx, cy;
    T z_min, u_min, u_max, v_min, v_max;
};

__device__ __forceinline__ float fast_rcp(float z) {
    float x;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(z));
    return x;
}
__device__ __forceinline__ void sincos_t(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ void sincos_t(double a, float* s, float* c) {
    double sd, cd;  // once per pass: keep the rotation exact, round once
    sincos(a, &sd, &cd);
    *s = (float)sd;
    *c = (float)cd;
}

// ------------------------------------------------------------------ fused residual / Jacobian / normal-equation pass
// Accumulates, over this lane's share of the n active points,
//   acc[0]     sum |r|^2
//   acc[1..4]  J^T r        (order yaw, tx, ty, tz)
//   acc[5..14] J^T J upper triangle  (00 01 02 03 11 12 13 22 23 33)
// and butterfly-reduces them over the warp (every lane ends with the full sums).
// T = double reproduces the fp64 reference arithmetic; T = float is the fast path.
// CLIPSEM 0: Ceres-Jet semantics (pnp_uncert_cpu.cpp:36-42): z clip drops only d/dz', u/v clamp drops that row.
// CLIPSEM 1: jacobian.py:52-59 semantics: a z-clipped point loses both rows (used for the pipeline covariance).
// USE_BITS : skip points whose bit in `bits` (bit k <-> point 32k+lane) is 0 (outliers when not compacted).
// Returns in `clip` whether any point processed by this lane hit a clip.
template <typename T, int WMODE, int LAYOUT, int CLIPSEM, bool USE_BITS>
__device__ __forceinline__ void eval_pass(const float* __restrict__ s3, const float* __restrict__ s2,
                                          const float* __restrict__ sw, int P, int n, int lane, uint32_t bits,
                                          const double x[4], const Camera<T>& cam, T acc[kNumAcc], bool& clip) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    T sn, cs;
    sincos_t(x[0], &sn, &cs);
    const T tx = (T)x[1], ty = (T)x[2], tz = (T)x[3];
    const T zero = (T)0, one = (T)1;
#pragma unroll
    for (int i = 0; i < kNumAcc; ++i) acc[i] = zero;
    bool any = false;
#pragma unroll 2
    for (int p = lane, k = 0; p < n; p += 32, ++k) {
        if (USE_BITS && !((bits >> k) & 1u)) continue;
        const T X = (T)s3[sidx<LAYOUT, 3>(p, 0, P)];
        const T Y = (T)s3[sidx<LAYOUT, 3>(p, 1, P)];
        const T Z = (T)s3[sidx<LAYOUT, 3>(p, 2, P)];
        const T uo = (T)s2[sidx<LAYOUT, 2>(p, 0, P)];
        const T vo = (T)s2[sidx<LAYOUT, 2>(p, 1, P)];
        const T qx = fma(cs, X, sn * Z);
        const T qz = fma(cs, Z, -sn * X);
        const T xc = qx + tx, yc = Y + ty, zc = qz + tz;
        const bool zfree = !(zc < cam.z_min);
        const T z = zfree ? zc : cam.z_min;
        const T iz = fast_rcp(z);
        const T xn = xc * iz, yn = yc * iz;
        T pu = fma(cam.fx, xn, cam.cx);
        T pv = fma(cam.fy, yn, cam.cy);
        bool ufree = true, vfree = true;
        if (pu < cam.u_min) { pu = cam.u_min; ufree = false; } else if (pu > cam.u_max) { pu = cam.u_max; ufree = false; }
        if (pv < cam.v_min) { pv = cam.v_min; vfree = false; } else if (pv > cam.v_max) { pv = cam.v_max; vfree = false; }
        any = any || !zfree || !ufree || !vfree;
        const T du = pu - uo, dv = pv - vo;
        const T mz = zfree ? one : zero;
        if (CLIPSEM == 1 && !zfree) { ufree = false; vfree = false; }
        // unweighted projection Jacobian rows: Ju = (ju0, a_u, 0, b_u), Jv = (jv0, 0, a_v, b_v)
        const T au = ufree ? cam.fx * iz : zero;
        const T av = vfree ? cam.fy * iz : zero;
        const T bu = -au * xn * mz, bv = -av * yn * mz;
        const T ju0 = fma(au, qz, -bu * qx);
        const T jv0 = -bv * qx;
        if (WMODE != MRPNP_W_FULL) {
            const T wu = (T)sw[sidx<LAYOUT, WC>(p, 0, P)];
            const T wv = (T)sw[sidx<LAYOUT, WC>(p, 1, P)];
            const T ru = wu * du, rv = wv * dv;
            const T a0 = wu * ju0, a1 = wu * au, a3 = wu * bu;   // row u: (a0, a1, 0, a3)
            const T b0 = wv * jv0, b2 = wv * av, b3 = wv * bv;   // row v: (b0, 0, b2, b3)
            acc[0] = fma(ru, ru, fma(rv, rv, acc[0]));
            acc[1] = fma(a0, ru, fma(b0, rv, acc[1]));
            acc[2] = fma(a1, ru, acc[2]);
            acc[3] = fma(b2, rv, acc[3]);
            acc[4] = fma(a3, ru, fma(b3, rv, acc[4]));
            acc[5] = fma(a0, a0, fma(b0, b0, acc[5]));
            acc[6] = fma(a0, a1, acc[6]);
            acc[7] = fma(b0, b2, acc[7]);
            acc[8] = fma(a0, a3, fma(b0, b3, acc[8]));
            acc[9] = fma(a1, a1, acc[9]);
            // acc[10] (tx,ty) is identically zero for diagonal weights
            acc[11] = fma(a1, a3, acc[11]);
            acc[12] = fma(b2, b2, acc[12]);
            acc[13] = fma(b2, b3, acc[13]);
            acc[14] = fma(a3, a3, fma(b3, b3, acc[14]));
        } else {
            const T wxx = (T)sw[sidx<LAYOUT, WC>(p, 0, P)];
            const T wxy = (T)sw[sidx<LAYOUT, WC>(p, 1, P)];
            const T wyy = (T)sw[sidx<LAYOUT, WC>(p, 2, P)];
            const T r0 = fma(wxx, du, wxy * dv), r1 = fma(wxy, du, wyy * dv);
            // whitened rows: r0-row = wxx*Ju + wxy*Jv, r1-row = wxy*Ju + wyy*Jv
            const T a0 = fma(wxx, ju0, wxy * jv0), a1 = wxx * au, a2 = wxy * av, a3 = fma(wxx, bu, wxy * bv);
            const T b0 = fma(wxy, ju0, wyy * jv0), b1 = wxy * au, b2 = wyy * av, b3 = fma(wxy, bu, wyy * bv);
            acc[0] = fma(r0, r0, fma(r1, r1, acc[0]));
            acc[1] = fma(a0, r0, fma(b0, r1, acc[1]));
            acc[2] = fma(a1, r0, fma(b1, r1, acc[2]));
            acc[3] = fma(a2, r0, fma(b2, r1, acc[3]));
            acc[4] = fma(a3, r0, fma(b3, r1, acc[4]));
            acc[5] = fma(a0, a0, fma(b0, b0, acc[5]));
            acc[6] = fma(a0, a1, fma(b0, b1, acc[6]));
            acc[7] = fma(a0, a2, fma(b0, b2, acc[7]));
            acc[8] = fma(a0, a3, fma(b0, b3, acc[8]));
            acc[9] = fma(a1, a1, fma(b1, b1, acc[9]));
            acc[10] = fma(a1, a2, fma(b1, b2, acc[10]));
            acc[11] = fma(a1, a3, fma(b1, b3, acc[11]));
            acc[12] = fma(a2, a2, fma(b2, b2, acc[12]));
            acc[13] = fma(a2, a3, fma(b2, b3, acc[13]));
            acc[14] = fma(a3, a3, fma(b3, b3, acc[14]));
        }
    }
    clip = any;
#pragma unroll
    for (int i = 0; i < kNumAcc; ++i) acc[i] = warp_sum(acc[i]);
}

// ------------------------------------------------------------------ 4x4 SPD helpers (fp64, fully unrolled)
// Upper-triangle packing index: (0,0)=0 (0,1)=1 (0,2)=2 (0,3)=3 (1,1)=4 (1,2)=5 (1,3)=6 (2,2)=7 (2,3)=8 (3,3)=9.
__device__ __forceinline__ constexpr int tri(int i, int j) {
    return i <= j ? (i * 4 - (i * (i - 1)) / 2 + (j - i)) : (j * 4 - (j * (j - 1)) / 2 + (i - j));
}

// Cholesky A = L L^T of the packed symmetric A; L packed by tri(j,i) (i >= j).  false if not SPD.
__device__ __forceinline__ bool chol4(const double A[10], double L[10]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double d = A[tri(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) d = fma(-L[tri(k, j)], L[tri(k, j)], d);
        ok = ok && (d > 0.0) && (d < 1.7e308);
        const double ld = sqrt(d);
        const double inv = 1.0 / ld;
        L[tri(j, j)] = ld;
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            double s = A[tri(j, i)];
#pragma unroll
            for (int k = 0; k < j; ++k) s = fma(-L[tri(k, i)], L[tri(k, j)], s);
            L[tri(j, i)] = s * inv;  // element L[i][j] stored at tri(j,i)
        }
    }
    return ok;
}

// Solve L L^T y = b.
__device__ __forceinline__ void chol4_solve(const double L[10], const double b[4], double y[4]) {
    double w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s = fma(-L[tri(k, i)], w[k], s);
        w[i] = s / L[tri(i, i)];
    }
#pragma unroll
    for (int i = 3; i >= 0; --i) {
        double s = w[i];
#pragma unroll
        for (int k = i + 1; k < 4; ++k) s = fma(-L[tri(i, k)], y[k], s);
        y[i] = s / L[tri(i, i)];
    }
}

// Full inverse of the packed SPD matrix H into row-major inv[16]; false if not SPD.
__device__ __forceinline__ bool spd_inverse4(const double H[10], double inv[16]) {
    double L[10];
    if (!chol4(H, L)) return false;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double e[4] = {0.0, 0.0, 0.0, 0.0}, y[4];
        e[c] = 1.0;
        chol4_solve(L, e, y);
#pragma unroll
        for (int r = 0; r < 4; ++r) inv[r * 4 + c] = y[r];
    }
    return true;
}


// ------------------------------------------------------------------ on-device linear initialiser
// Replaces cv2.solvePnP(EPNP) of pnp_uncert_cpu.py:53-58 as the LM starting point.  With the rotation
// restricted to yaw (pnp_uncert_cpu.cpp:28) the projection equations are linear in (cos, sin, tx, ty, tz):
//     c (X - un Z) + s (Z + un X) + tx - un tz = 0          un = (u - cx) / fx
//     c (   - vn Z) + s (    vn X) + ty - vn tz = -Y         vn = (v - cy) / fy
// Stage A solves the weighted 5x5 normal equations; stage B fixes yaw = atan2(s, c) and re-solves the
// 3x3 system for t.  Rows are weighted by w*f so they approximate depth-scaled pixel residuals.
template <int N>
__device__ __forceinline__ bool chol_solve_dense(double A[N][N], const double b[N], double x[N]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double d = A[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d = fma(-A[j][k], A[j][k], d);
        ok = ok && (d > 0.0) && (d < 1.7e308);
        const double ld = sqrt(d), inv = 1.0 / ld;
        A[j][j] = ld;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            double v = A[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) v = fma(-A[i][k], A[j][k], v);
            A[i][j] = v * inv;
        }
    }
    double w[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double v = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v = fma(-A[i][k], w[k], v);
        w[i] = v / A[i][i];
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
        double v = w[i];
#pragma unroll
        for (int k = i + 1; k < N; ++k) v = fma(-A[k][i], x[k], v);
        x[i] = v / A[i][i];
    }
    return ok;
}

template <typename T, int WMODE, int LAYOUT>
__device__ __forceinline__ bool linear_init(const float* __restrict__ s3, const float* __restrict__ s2,
                                            const float* __restrict__ sw, int P, int n, int lane,
                                            const Camera<T>& cam, double x[4]) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    const T ifx = (T)1 / cam.fx, ify = (T)1 / cam.fy;
    // ---- stage A: 5 unknowns ----
    T m[15], r[5];
#pragma unroll
    for (int i = 0; i < 15; ++i) m[i] = (T)0;
#pragma unroll
    for (int i = 0; i < 5; ++i) r[i] = (T)0;
    for (int p = lane; p < n; p += 32) {
        const T X = (T)s3[sidx<LAYOUT, 3>(p, 0, P)], Y = (T)s3[sidx<LAYOUT, 3>(p, 1, P)], Z = (T)s3[sidx<LAYOUT, 3>(p, 2, P)];
        const T un = ((T)s2[sidx<LAYOUT, 2>(p, 0, P)] - cam.cx) * ifx;
        const T vn = ((T)s2[sidx<LAYOUT, 2>(p, 1, P)] - cam.cy) * ify;
        const T wu = (T)sw[sidx<LAYOUT, WC>(p, 0, P)] * cam.fx, wv = (T)sw[sidx<LAYOUT, WC>(p, WC - 1, P)] * cam.fy;
        const T a[5] = {wu * (X - un * Z), wu * (Z + un * X), wu, (T)0, -wu * un};
        const T b[5] = {-wv * vn * Z, wv * vn * X, (T)0, wv, -wv * vn};
        const T rb = -wv * Y;
        int q = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
            for (int j = i; j < 5; ++j, ++q) m[q] = fma(a[i], a[j], fma(b[i], b[j], m[q]));
            r[i] = fma(b[i], rb, r[i]);
        }
    }
    double A[5][5], rhs[5], sol[5];
    {
        int q = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
            for (int j = i; j < 5; ++j, ++q) {
                const double v = (double)warp_sum(m[q]);
                A[i][j] = v;
                A[j][i] = v;
            }
            rhs[i] = (double)warp_sum(r[i]);
        }
    }
    if (!chol_solve_dense<5>(A, rhs, sol)) return false;
    const double yaw = atan2(sol[1], sol[0]);
    // ---- stage B: translation with yaw fixed ----
    T sn, cs;
    sincos_t(yaw, &sn, &cs);
    T m3[6], r3[3];
#pragma unroll
    for (int i = 0; i < 6; ++i) m3[i] = (T)0;
#pragma unroll
    for (int i = 0; i < 3; ++i) r3[i] = (T)0;
    for (int p = lane; p < n; p += 32) {
        const T X = (T)s3[sidx<LAYOUT, 3>(p, 0, P)], Y = (T)s3[sidx<LAYOUT, 3>(p, 1, P)], Z = (T)s3[sidx<LAYOUT, 3>(p, 2, P)];
        const T un = ((T)s2[sidx<LAYOUT, 2>(p, 0, P)] - cam.cx) * ifx;
        const T vn = ((T)s2[sidx<LAYOUT, 2>(p, 1, P)] - cam.cy) * ify;
        const T wu = (T)sw[sidx<LAYOUT, WC>(p, 0, P)] * cam.fx, wv = (T)sw[sidx<LAYOUT, WC>(p, WC - 1, P)] * cam.fy;
        const T qx = fma(cs, X, sn * Z), qz = fma(cs, Z, -sn * X);
        // rows: wu [1 0 -un] t = -wu (qx - un qz);  wv [0 1 -vn] t = -wv (Y - vn qz)
        const T ra = -wu * (qx - un * qz), rb = -wv * (Y - vn * qz);
        const T a2 = -wu * un, b2 = -wv * vn;
        m3[0] = fma(wu, wu, m3[0]);                 // (0,0)
        m3[1] = fma(wu, a2, m3[1]);                 // (0,2)
        m3[2] = fma(wv, wv, m3[2]);                 // (1,1)
        m3[3] = fma(wv, b2, m3[3]);                 // (1,2)
        m3[4] = fma(a2, a2, fma(b2, b2, m3[4]));    // (2,2)
        r3[0] = fma(wu, ra, r3[0]);
        r3[1] = fma(wv, rb, r3[1]);
        r3[2] = fma(a2, ra, fma(b2, rb, r3[2]));
    }
    double B[3][3], rh3[3], t[3];
    B[0][0] = (double)warp_sum(m3[0]);
    B[0][1] = B[1][0] = 0.0;
    B[0][2] = B[2][0] = (double)warp_sum(m3[1]);
    B[1][1] = (double)warp_sum(m3[2]);
    B[1][2] = B[2][1] = (double)warp_sum(m3[3]);
    B[2][2] = (double)warp_sum(m3[4]);
#pragma unroll
    for (int i = 0; i < 3; ++i) rh3[i] = (double)warp_sum(r3[i]);
    if (!chol_solve_dense<3>(B, rh3, t)) return false;
    x[0] = yaw; x[1] = t[0]; x[2] = t[1]; x[3] = t[2];
    return (fabs(t[0]) + fabs(t[1]) + fabs(t[2])) < 1.7e308;
}

}  // namespace mrpnp
