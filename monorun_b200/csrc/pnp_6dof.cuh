// pnp_6dof.cuh -- the 6-DoF extension of the uncertainty PnP (BASELINE.json north star): unknowns
// [rx, ry, rz, tx, ty, tz] with the rotation of ceres::AngleAxisRotatePoint; projection, depth / image clips,
// per-axis or full 2x2 whitening, Ceres 1.14 trust-region LM and the covariance (J^T J)^-1 exactly as the reference's
// 4-DoF op (monorun/ops/least_squares/src/pnp_uncert_cpu.cpp:9-74, :245-292).  The reference itself has no 6-DoF code
// (its rotation vector is hard-wired to (0, yaw, 0), :28; `use_6dof` is never read, pnp_uncert.py:11,98,122,142), so
// this variant is checked against its own fp64 CPU statement (oracle/pnp_6dof_oracle.cpp, dual-number Jacobians).
//
// Same structure as pnp_noc.cuh: one warp per object, fp64, 28 accumulators per lane (cost, gradient, upper triangle
// of J^T J), butterfly reduction, the shared controller of lm_dense.cuh run redundantly by every lane; one more
// Jacobian pass at the returned pose gives the 6x6 covariance.  The rotation derivative is closed-form:
//   d(R p)/dw_k = (dR/dw_k) p,   dR/dw_k = ( w_k [w]x + [w x (I - R) e_k]x ) R / theta^2     (theta^2 > epsilon)
//                                dR/dw_k = [e_k]x                                           (first-order branch)
// which is what automatic differentiation of AngleAxisRotatePoint yields in either branch.
#pragma once
#include <math.h>
#include <stdint.h>

#include "lm_dense.cuh"

namespace mr6 {

constexpr int kNP = 6;
constexpr int kNAcc = mrlm::Layout<kNP>::kNAcc;  // 28
constexpr int kAccG = mrlm::Layout<kNP>::kAccG, kAccH = mrlm::Layout<kNP>::kAccH;
constexpr int kResultStride = 48;  // rvec(3), t(3) | cov 6x6 | valid, iterations, final_cost, cost_evals, termination, pad

struct Camera { double fx, fy, cx, cy, z_min, u_min, u_max, v_min, v_max; };

struct Pose6 {
    double R[9];      // rotation (row-major); the first-order branch uses I + [w]x like the Ceres code
    double dR[3][9];  // dR/dw_k
    double t[3];
};

MRLM_HD void skew(const double* v, double* m) {
    m[0] = 0.0;   m[1] = -v[2]; m[2] = v[1];
    m[3] = v[2];  m[4] = 0.0;   m[5] = -v[0];
    m[6] = -v[1]; m[7] = v[0];  m[8] = 0.0;
}

MRLM_HD void mat3_mul(const double* a, const double* b, double* c) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

MRLM_HD_NOINLINE void make_pose(const double* x, Pose6* p) {
    const double w[3] = {x[0], x[1], x[2]};
    const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    p->t[0] = x[3]; p->t[1] = x[4]; p->t[2] = x[5];
    double W[9];
    skew(w, W);
    if (theta2 > 2.220446049250313e-16) {  // ceres rotation.h: theta2 > std::numeric_limits<double>::epsilon()
        const double theta = sqrt(theta2), c = cos(theta), s = sin(theta), inv = 1.0 / theta;
        const double k[3] = {w[0] * inv, w[1] * inv, w[2] * inv};
        double Kx[9];
        skew(k, Kx);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                p->R[i * 3 + j] = (i == j ? c : 0.0) + s * Kx[i * 3 + j] + (1.0 - c) * k[i] * k[j];
        for (int a = 0; a < 3; ++a) {
            // u = w x (I - R) e_a
            const double col[3] = {(a == 0 ? 1.0 : 0.0) - p->R[a], (a == 1 ? 1.0 : 0.0) - p->R[3 + a],
                                   (a == 2 ? 1.0 : 0.0) - p->R[6 + a]};
            const double u[3] = {w[1] * col[2] - w[2] * col[1], w[2] * col[0] - w[0] * col[2],
                                 w[0] * col[1] - w[1] * col[0]};
            double U[9], M[9];
            skew(u, U);
            for (int i = 0; i < 9; ++i) M[i] = (w[a] * W[i] + U[i]) / theta2;
            mat3_mul(M, p->R, p->dR[a]);
        }
    } else {
        for (int i = 0; i < 9; ++i) p->R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + W[i];
        for (int a = 0; a < 3; ++a) {
            const double e[3] = {a == 0 ? 1.0 : 0.0, a == 1 ? 1.0 : 0.0, a == 2 ? 1.0 : 0.0};
            skew(e, p->dR[a]);
        }
    }
}

// One reprojection block (pnp_uncert_cpu.cpp:24-51 / :189-217 with r_vec = x[0..3)), ceres::Jet branch semantics.
template <bool FULLW, bool JAC>
MRLM_HD void add_point(const Camera& cam, const Pose6& ps, double X, double Y, double Z, double u, double v,
                       double w0, double w1, double w2, double* acc) {
    const double* R = ps.R;
    const double xc = R[0] * X + R[1] * Y + R[2] * Z + ps.t[0];
    const double yc = R[3] * X + R[4] * Y + R[5] * Z + ps.t[1];
    const double zc = R[6] * X + R[7] * Y + R[8] * Z + ps.t[2];
    const bool z_free = !(zc < cam.z_min);
    const double z = z_free ? zc : cam.z_min, iz = 1.0 / z;
    double pu = cam.fx * xc * iz + cam.cx, pv = cam.fy * yc * iz + cam.cy;
    bool u_free = true, v_free = true;
    if (pu < cam.u_min) { pu = cam.u_min; u_free = false; } else if (pu > cam.u_max) { pu = cam.u_max; u_free = false; }
    if (pv < cam.v_min) { pv = cam.v_min; v_free = false; } else if (pv > cam.v_max) { pv = cam.v_max; v_free = false; }
    const double du = pu - u, dv = pv - v;
    const double w00 = w0, w01 = FULLW ? w1 : 0.0, w11 = FULLW ? w2 : w1;
    const double r0 = w00 * du + w01 * dv, r1 = w01 * du + w11 * dv;
    acc[0] += 0.5 * (r0 * r0 + r1 * r1);
    if (JAC) {
        const double mz = z_free ? 1.0 : 0.0;
        const double au = u_free ? cam.fx * iz : 0.0, bu = u_free ? -cam.fx * xc * iz * iz * mz : 0.0;
        const double av = v_free ? cam.fy * iz : 0.0, bv = v_free ? -cam.fy * yc * iz * iz * mz : 0.0;
        double j0[kNP], j1[kNP];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double* D = ps.dR[k];
            const double dx = D[0] * X + D[1] * Y + D[2] * Z, dy = D[3] * X + D[4] * Y + D[5] * Z,
                         dz = D[6] * X + D[7] * Y + D[8] * Z;
            const double ju = au * dx + bu * dz, jv = av * dy + bv * dz;
            j0[k] = w00 * ju + w01 * jv;
            j1[k] = w01 * ju + w11 * jv;
        }
        {   // translation columns: d(x',y',z')/dt = I
            const double ju[3] = {au, 0.0, bu}, jv[3] = {0.0, av, bv};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                j0[3 + k] = w00 * ju[k] + w01 * jv[k];
                j1[3 + k] = w01 * ju[k] + w11 * jv[k];
            }
        }
#pragma unroll
        for (int a = 0; a < kNP; ++a) {
            acc[kAccG + a] += j0[a] * r0 + j1[a] * r1;
#pragma unroll
            for (int b = a; b < kNP; ++b) acc[kAccH + mrlm::tri<kNP>(a, b)] += j0[a] * j0[b] + j1[a] * j1[b];
        }
    }
}

// (J^T J)^-1 from the upper triangle in acc (ceres::Covariance, pnp_uncert_cpu.cpp:279-291); false if not SPD.
MRLM_HD_NOINLINE bool covariance(const double* acc, double* cov) {
    double H[kNP * kNP];
    for (int a = 0; a < kNP; ++a)
        for (int b = a; b < kNP; ++b) {
            H[a * kNP + b] = acc[kAccH + mrlm::tri<kNP>(a, b)];
            H[b * kNP + a] = H[a * kNP + b];
        }
    bool ok = true;
    for (int c = 0; c < kNP; ++c) {
        double e[kNP], y[kNP];
        for (int i = 0; i < kNP; ++i) e[i] = (i == c) ? 1.0 : 0.0;
        ok = mrlm::cholesky_solve<kNP>(H, e, y) && ok;
        for (int i = 0; i < kNP; ++i) cov[i * kNP + c] = y[i];
    }
    return ok;
}

struct KParams {
    const float *coords_3d, *coords_2d, *weights, *cam_mats, *uv_range, *init;
    const uint32_t* inlier;
    double* result;  // [N,48]
    int n_obj, n_pts, planar, wmode /* 0 logstd, 1 istd, 2 full */, cam_stride, range_stride, max_iterations;
    double z_min, std_scale;
};

#ifdef __CUDACC__

constexpr int kWarpsPerCta = 4;
#ifndef MR6_MIN_BLOCKS
#define MR6_MIN_BLOCKS 2  // CTAs per SM the register allocation aims at (3 measured 10 % slower: Pose6 spills)
#endif

template <bool FULLW>
struct WarpPass {
    const KParams& kp;
    Camera cam;
    const float *c3, *c2, *cw;
    const uint16_t* idx;   // shared memory: the indices of the object's inliers, in point order (built once per object), so that
    int n_idx;             // every pass runs with full lanes instead of skipping ~half of them (round 2: 1.7x)
    int lane;

    __device__ __forceinline__ float at(const float* base, int c, int nc, int p) const {
        return kp.planar ? __ldg(base + (size_t)c * kp.n_pts + p) : __ldg(base + (size_t)p * nc + c);
    }

    template <bool JAC>
    __device__ __forceinline__ void run(const double* x, double* acc) const {
        Pose6 ps;
        make_pose(x, &ps);
        constexpr int wc = FULLW ? 3 : 2;
        for (int j = lane; j < n_idx; j += 32) {
            const int p = idx[j];
            {
                double w0 = at(cw, 0, wc, p), w1 = at(cw, 1, wc, p);
                const double w2 = FULLW ? (double)at(cw, 2, wc, p) : 0.0;
                if (!FULLW && kp.wmode == 0) { w0 = exp(-w0) / kp.std_scale; w1 = exp(-w1) / kp.std_scale; }
                add_point<FULLW, JAC>(cam, ps, at(c3, 0, 3, p), at(c3, 1, 3, p), at(c3, 2, 3, p), at(c2, 0, 2, p),
                                      at(c2, 1, 2, p), w0, w1, w2, acc);
            }
        }
        constexpr int n = JAC ? kNAcc : 1;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            double v = acc[i];
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
            acc[i] = v;
        }
    }

    __device__ void operator()(const double* x, bool jac, double* acc) const {
        if (jac) run<true>(x, acc); else run<false>(x, acc);
    }
};

template <bool FULLW>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MR6_MIN_BLOCKS) pnp_6dof_kernel(const KParams kp) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * kWarpsPerCta;
    constexpr int wc = FULLW ? 3 : 2;
    __shared__ uint16_t idx_s[kWarpsPerCta][1024];   // MRPNP_MAX_POINTS
    uint16_t* my_idx = idx_s[threadIdx.x >> 5];
    for (int obj = warp; obj < kp.n_obj; obj += n_warps) {
        const float* K = kp.cam_mats + (size_t)obj * kp.cam_stride;
        const float* rg = kp.uv_range + (size_t)obj * kp.range_stride;
        WarpPass<FULLW> pass{kp};
        pass.cam.fx = K[0]; pass.cam.fy = K[4]; pass.cam.cx = K[2]; pass.cam.cy = K[5];
        pass.cam.z_min = kp.z_min;
        pass.cam.u_min = rg[0]; pass.cam.u_max = rg[1]; pass.cam.v_min = rg[2]; pass.cam.v_max = rg[3];
        pass.c3 = kp.coords_3d + (size_t)obj * 3 * kp.n_pts;
        pass.c2 = kp.coords_2d + (size_t)obj * 2 * kp.n_pts;
        pass.cw = kp.weights + (size_t)obj * wc * kp.n_pts;
        {   // inlier index list (order-preserving: ballot + prefix popcount per row of 32 points)
            const uint32_t* mask = kp.inlier ? kp.inlier + (size_t)obj * ((kp.n_pts + 31) >> 5) : nullptr;
            int count = 0;
            __syncwarp();
            for (int base = 0; base < kp.n_pts; base += 32) {
                const int p = base + lane;
                bool on = p < kp.n_pts;
                if (mask) on = on && ((__ldg(mask + (base >> 5)) >> lane) & 1u);
                const unsigned m = __ballot_sync(0xffffffffu, on);
                if (on) my_idx[count + __popc(m & ((1u << lane) - 1u))] = (uint16_t)p;
                count += __popc(m);
            }
            __syncwarp();
            pass.idx = my_idx;
            pass.n_idx = count;
        }
        pass.lane = lane;
        double x[kNP];
        for (int k = 0; k < kNP; ++k) x[k] = kp.init[(size_t)obj * kNP + k];
        mrlm::LMOptions opt = mrlm::default_options();
        if (kp.max_iterations > 0) opt.max_num_iterations = kp.max_iterations;
        const mrlm::LMResult r = mrlm::minimize<kNP>(pass, x, opt);
        bool valid = (r.term == mrlm::kConvergence || r.term == mrlm::kNoConvergence);  // IsSolutionUsable
        double acc[kNAcc], cov[kNP * kNP];
        for (int i = 0; i < kNAcc; ++i) acc[i] = 0.0;
        pass(x, true, acc);
        const bool spd = covariance(acc, cov);
        if (lane == 0) {
            double* out = kp.result + (size_t)obj * kResultStride;
            for (int k = 0; k < kNP; ++k) out[k] = x[k];
            for (int i = 0; i < kNP * kNP; ++i) out[kNP + i] = (valid && spd) ? cov[i] : ((i % (kNP + 1) == 0) ? 1.0 : 0.0);
            out[42] = (valid && spd) ? 1.0 : 0.0;  // Covariance::Compute failing clears result_val (cpp:287)
            out[43] = r.iterations;
            out[44] = r.final_cost;
            out[45] = r.cost_evals;
            out[46] = r.term;
            out[47] = 0.0;
        }
    }
}

#endif  // __CUDACC__

}  // namespace mr6
