// pnp_noc.cuh -- the reference's two 7-parameter solvers on the GPU:
//   pnp_noc_uncert      monorun/ops/least_squares/src/pnp_uncert_cpu.cpp:294-334 (ext.h:15-28)
//   pnp_noc_cov_uncert  monorun/ops/least_squares/src/pnp_uncert_cpu.cpp:336-377 (ext.h:30-43)
// Unknowns [log l, log h, log w, yaw, tx, ty, tz]; residual blocks: one 2-vector per point
// (normalised object coordinates scaled by exp(log dims), rotated, translated, projected, clipped,
// whitened -- :122-148 diag, :189-217 full 2x2) and one 3-vector dimension prior (:77-104); every
// block goes through ceres::HuberLoss(delta).
//
// One warp per object, everything in fp64.  A pass over the points keeps, per lane, the cost, the
// gradient and the upper triangle of J^T J of the *robustified* problem (36 accumulators), which a
// butterfly reduction leaves bit-identical in all 32 lanes; the trust-region controller then runs
// redundantly (and warp-uniformly) in every lane on those 36 numbers, so it is plain scalar code.
// That code -- the point functor, the Huber corrector and the controller -- is __host__ __device__:
// tests/harness/noc_host_harness.cpp compiles it with g++ and drives it with an emulated warp, which
// lets the CPU test suite check the solver logic against the oracle without a GPU.  The harness is
// test code; the library has no CPU path.
//
// Ceres solves each Levenberg-Marquardt step by QR of [J S; sqrt(D/radius)]; here the same step
// comes from the 7x7 normal equations (S J^T J S + D/radius) y = S J^T r by Cholesky in fp64.
#pragma once
#include <math.h>
#include <stdint.h>

#include "lm_dense.cuh"

#ifdef __CUDACC__
#define MRNOC_HD __host__ __device__ __forceinline__
#define MRNOC_HD_NOINLINE __host__ __device__ __noinline__
#else
#define MRNOC_HD inline
#define MRNOC_HD_NOINLINE inline
#endif

namespace mrnoc {

constexpr int kNP = 7;  // parameters
constexpr int kNAcc = mrlm::Layout<kNP>::kNAcc;  // cost | gradient | upper triangle of J^T J  (36)
constexpr int kAccG = mrlm::Layout<kNP>::kAccG, kAccH = mrlm::Layout<kNP>::kAccH;
using mrlm::LMOptions;
using mrlm::LMResult;
using mrlm::default_options;
using mrlm::kConvergence;
using mrlm::kNoConvergence;
using mrlm::kFailure;

MRNOC_HD int tri(int a, int b) { return mrlm::tri<kNP>(a, b); }

struct Camera { double fx, fy, cx, cy, z_min, u_min, u_max, v_min, v_max; };

// What a pass needs from the parameter vector.
struct DimPose {
    double e[3];   // exp(log dims)                          cpp:125-127
    double sn, cs; // sin / cos yaw                          cpp:128-131 (AngleAxisRotatePoint about y)
    double t[3];
};

MRNOC_HD DimPose make_dimpose(const double* x) {
    DimPose d;
    d.e[0] = exp(x[0]); d.e[1] = exp(x[1]); d.e[2] = exp(x[2]);
    d.sn = sin(x[3]); d.cs = cos(x[3]);
    d.t[0] = x[4]; d.t[1] = x[5]; d.t[2] = x[6];
    return d;
}

// ceres::HuberLoss::Evaluate for s = |r_block|^2 (a = delta, b = delta^2): rho and rho'.  rho'' <= 0
// everywhere, so ceres::Corrector always scales residuals and Jacobian rows by sqrt(rho') (alpha = 0).
MRNOC_HD void huber(double a, double s, double* rho0, double* rho1) {
    const double b = a * a;
    if (s > b) {
        const double r = sqrt(s);
        *rho0 = 2.0 * a * r - b;
        *rho1 = fmax(2.2250738585072014e-308, a / r);
    } else {
        *rho0 = s;
        *rho1 = 1.0;
    }
}

// Adds one robustified block with residuals r[nr] and Jacobian rows j[nr][7] to acc:
// cost += rho/2, gradient += rho' J^T r, J^T J += rho' J^T J.
template <int NR, bool JAC>
MRNOC_HD void add_block(double delta, const double* r, const double (*j)[kNP], double* acc) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NR; ++i) s += r[i] * r[i];
    double rho0, rho1;
    huber(delta, s, &rho0, &rho1);
    acc[0] += 0.5 * rho0;
    if (JAC) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
#pragma unroll
            for (int a = 0; a < kNP; ++a) {
                const double ja = rho1 * j[i][a];
                acc[kAccG + a] += ja * r[i];
#pragma unroll
                for (int b = a; b < kNP; ++b) acc[kAccH + tri(a, b)] += ja * j[i][b];
            }
        }
    }
}

// One reprojection block, derivative semantics of ceres::Jet through max() and the clamping
// ternaries (a clipped depth keeps d/dx' but loses d/dz'; a clamped u or v has zero derivative).
template <bool FULLW, bool JAC>
MRNOC_HD void add_point(const Camera& cam, const DimPose& d, double delta, double X, double Y, double Z,
                        double u, double v, double w0, double w1, double w2, double* acc) {
    const double Sx = X * d.e[0], Sy = Y * d.e[1], Sz = Z * d.e[2];
    const double qx = d.cs * Sx + d.sn * Sz, qz = -d.sn * Sx + d.cs * Sz;
    const double xc = qx + d.t[0], yc = Sy + d.t[1], zc = qz + d.t[2];
    const bool z_free = !(zc < cam.z_min);                                  // cpp:136
    const double z = z_free ? zc : cam.z_min, iz = 1.0 / z;
    double pu = cam.fx * xc * iz + cam.cx, pv = cam.fy * yc * iz + cam.cy;  // cpp:138-139
    bool u_free = true, v_free = true;
    if (pu < cam.u_min) { pu = cam.u_min; u_free = false; } else if (pu > cam.u_max) { pu = cam.u_max; u_free = false; }
    if (pv < cam.v_min) { pv = cam.v_min; v_free = false; } else if (pv > cam.v_max) { pv = cam.v_max; v_free = false; }
    const double du = pu - u, dv = pv - v;                                  // cpp:144-145
    const double w00 = w0, w01 = FULLW ? w1 : 0.0, w11 = FULLW ? w2 : w1;   // cpp:147-148 / :214-215
    double r[2] = {w00 * du + w01 * dv, w01 * du + w11 * dv};
    double j[2][kNP];
    if (JAC) {
        const double mz = z_free ? 1.0 : 0.0;
        const double au = u_free ? cam.fx * iz : 0.0, bu = u_free ? -cam.fx * xc * iz * iz * mz : 0.0;
        const double av = v_free ? cam.fy * iz : 0.0, bv = v_free ? -cam.fy * yc * iz * iz * mz : 0.0;
        // d(x', y', z') / d(log l, log h, log w, yaw, tx, ty, tz)
        const double dx[kNP] = {d.cs * Sx, 0.0, d.sn * Sz, qz, 1.0, 0.0, 0.0};
        const double dy[kNP] = {0.0, Sy, 0.0, 0.0, 0.0, 1.0, 0.0};
        const double dz[kNP] = {-d.sn * Sx, 0.0, d.cs * Sz, -qx, 0.0, 0.0, 1.0};
#pragma unroll
        for (int k = 0; k < kNP; ++k) {
            const double ju = au * dx[k] + bu * dz[k], jv = av * dy[k] + bv * dz[k];
            j[0][k] = w00 * ju + w01 * jv;
            j[1][k] = w01 * ju + w11 * jv;
        }
    }
    add_block<2, JAC>(delta, r, j, acc);
}

// DimErrorArray (cpp:77-104): r_k = wgt_k (x_k - logdim_k), one 3-residual block.
template <bool JAC>
MRNOC_HD void add_dim_prior(const double* x, const double* logdim, const double* logdim_wgt, double delta,
                            double* acc) {
    double r[3], j[3][kNP];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r[k] = logdim_wgt[k] * (x[k] - logdim[k]);
#pragma unroll
        for (int c = 0; c < kNP; ++c) j[k][c] = (c == k) ? logdim_wgt[k] : 0.0;
    }
    add_block<3, JAC>(delta, r, j, acc);
}

// Kernel arguments (device pointers).  Layout codes as MRPNP_LAYOUT_* of monorun_pnp.h.
struct KParams {
    const float* coords_3d;   // normalised object coordinates
    const float* coords_2d;
    const float* weights;     // C = 2 (istd per axis) or 3 (wxx, wxy, wyy)
    const float* logdim;      // [N,3]
    const float* logdim_wgt;  // [N,3]
    const float* cam_mats;    // [N|1,9]
    const float* uv_range;    // [N|1,4]
    const float* init;        // [N,7]
    const uint32_t* inlier;   // [N, ceil(P/32)] or NULL
    double* result;           // [N,12]
    int n_obj, n_pts, planar, cam_stride, range_stride, max_iterations;
    double z_min, delta;
};

constexpr int kResultStride = 12;  // dimpose[7], valid, iterations, final_cost, cost_evals, termination

#ifdef __CUDACC__

constexpr int kWarpsPerCta = 4;
#ifndef MRNOC_MIN_BLOCKS
#define MRNOC_MIN_BLOCKS 3  // CTAs per SM the register allocation aims at (168 registers; 2 measured 8-12 % slower)
#endif

__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
    return __shfl_xor_sync(0xffffffffu, v, m);
}

// One pass of a warp over its object's points.
template <bool FULLW>
struct WarpPass {
    const KParams& kp;
    Camera cam;
    const float *c3, *c2, *cw;
    const uint32_t* mask;
    double logdim[3], logdim_wgt[3];
    int lane;

    __device__ __forceinline__ float at(const float* base, int c, int nc, int p) const {
        return kp.planar ? __ldg(base + (size_t)c * kp.n_pts + p) : __ldg(base + (size_t)p * nc + c);
    }

    template <bool JAC>
    __device__ __forceinline__ void run(const double* x, double* acc) const {
        const DimPose d = make_dimpose(x);
        constexpr int wc = FULLW ? 3 : 2;
        for (int base = 0; base < kp.n_pts; base += 32) {
            const int p = base + lane;
            bool on = p < kp.n_pts;
            if (mask) on = on && ((__ldg(mask + (base >> 5)) >> lane) & 1u);
            if (on) {
                const double w2 = FULLW ? (double)at(cw, 2, wc, p) : 0.0;
                add_point<FULLW, JAC>(cam, d, kp.delta, at(c3, 0, 3, p), at(c3, 1, 3, p), at(c3, 2, 3, p),
                                      at(c2, 0, 2, p), at(c2, 1, 2, p), at(cw, 0, wc, p), at(cw, 1, wc, p), w2,
                                      acc);
            }
        }
        constexpr int n = JAC ? kNAcc : 1;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            double v = acc[i];
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) v += shfl_xor_f64(v, m);
            acc[i] = v;
        }
        add_dim_prior<JAC>(x, logdim, logdim_wgt, kp.delta, acc);  // cpp:321-323: the prior block comes last
    }

    __device__ void operator()(const double* x, bool jac, double* acc) const {
        if (jac) run<true>(x, acc); else run<false>(x, acc);
    }
};

template <bool FULLW>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MRNOC_MIN_BLOCKS) pnp_noc_kernel(const KParams kp) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * kWarpsPerCta;
    constexpr int wc = FULLW ? 3 : 2;
    for (int obj = warp; obj < kp.n_obj; obj += n_warps) {
        const float* K = kp.cam_mats + (size_t)obj * kp.cam_stride;
        const float* rg = kp.uv_range + (size_t)obj * kp.range_stride;
        WarpPass<FULLW> pass{kp};
        pass.cam.fx = K[0]; pass.cam.fy = K[4]; pass.cam.cx = K[2]; pass.cam.cy = K[5];  // cpp:312
        pass.cam.z_min = kp.z_min;
        pass.cam.u_min = rg[0]; pass.cam.u_max = rg[1]; pass.cam.v_min = rg[2]; pass.cam.v_max = rg[3];
        pass.c3 = kp.coords_3d + (size_t)obj * 3 * kp.n_pts;
        pass.c2 = kp.coords_2d + (size_t)obj * 2 * kp.n_pts;
        pass.cw = kp.weights + (size_t)obj * wc * kp.n_pts;
        pass.mask = kp.inlier ? kp.inlier + (size_t)obj * ((kp.n_pts + 31) >> 5) : nullptr;
        pass.lane = lane;
        double x[kNP];
        for (int k = 0; k < 3; ++k) {
            pass.logdim[k] = kp.logdim[(size_t)obj * 3 + k];
            pass.logdim_wgt[k] = kp.logdim_wgt[(size_t)obj * 3 + k];
        }
        for (int k = 0; k < kNP; ++k) x[k] = kp.init[(size_t)obj * kNP + k];
        LMOptions opt = default_options();
        if (kp.max_iterations > 0) opt.max_num_iterations = kp.max_iterations;
        const LMResult r = mrlm::minimize<kNP>(pass, x, opt);
        if (lane == 0) {
            double* out = kp.result + (size_t)obj * kResultStride;
            for (int k = 0; k < kNP; ++k) out[k] = x[k];
            out[7] = (r.term == kConvergence || r.term == kNoConvergence) ? 1.0 : 0.0;  // IsSolutionUsable, cpp:333
            out[8] = r.iterations;
            out[9] = r.final_cost;
            out[10] = r.cost_evals;
            out[11] = r.term;
        }
    }
}

#endif  // __CUDACC__

}  // namespace mrnoc
