// pnp_6dof_fast.cuh -- the 6-DoF solve (pnp_6dof.cuh) on the MIXED scheme of the 4-DoF solver (MRPNP_PREC_MIXED): the
// residual / cost chain of every evaluation in fp64, exactly the expressions of mr6::add_point, so that Ceres'
// trust-region decisions (lm_dense.cuh) are taken on the same numbers as in the fp64 kernel up to summation order; the
// Jacobian, J^T r and J^T J in fp32 (their rounding moves a step by ~1e-7 of its length and the covariance by ~1e-6).
//
// What differs from pnp_6dof_kernel besides the arithmetic:
//   * one read of the correspondences: the object's inliers are compacted once into the warp's shared-memory slot
//     (x y z u v as fp32 planes, then the weight planes: the three entries of the symmetric 2x2 matrix or the two
//     per-axis inverse deviations as floats, log-std weights exponentiated once in fp64 and kept as doubles), every pass
//     reads that;
//   * every evaluation computes cost AND normal equations (a candidate is accepted ~85 % of the time) and leaves its
//     totals in one of two stash entries per warp; the Jacobian evaluation Ceres asks for after accepting a step, and the
//     covariance evaluation at the returned pose, find them there: 1 + (LM iterations) passes per object instead of
//     2 + 2 x (LM iterations);
//   * the trust-region controller is mrlm::minimize cut at its cost evaluation (lm_advance: from one evaluation's totals
//     to the next candidate or the end), run by lane 0 on a per-warp state in shared memory -- with the L1 carved out for
//     the slots, the per-lane stack arrays of mrlm::minimize would be served by L2 (measured 1.8 ms against 1.05 ms) --
//     and handed to the warp through __syncwarp; the six columns of the covariance are back-substituted by six lanes;
//   * 27 fp32 sums (two residual rows side by side in packed FMAs) leave the lanes through a 31-shuffle transposed
//     reduction (lane L ends with total L), the cost through a 5-step fp64 butterfly;
//   * objects are handed out by an atomic counter (they differ in LM iterations), 8 warps per CTA, one CTA per SM.
// Checked against oracle/pnp_6dof_oracle.cpp and the fp64 kernel in tests/test_6dof_gpu.py; profiles/r02_6dof.txt.
#pragma once
#include "pnp_6dof.cuh"

namespace mr6 {

// The per-warp areas live in shared memory; functions that receive them by pointer (not inlined into the kernel) would use
// generic loads and stores with 64-bit address arithmetic.  MR6_NO_ASSUME_SHARED builds without the hint.
#if defined(__CUDA_ARCH__) && !defined(MR6_NO_ASSUME_SHARED)
#define MR6_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#else
#define MR6_ASSUME_SHARED(p) ((void)0)
#endif

struct StashEntry {
    double x[kNP];
    double cost;
    int pad[2];
    float tot[32];   // [0..5] J^T r, [6..26] upper triangle of J^T J (the order of mrlm::Layout<6> after the cost)
};
static_assert(sizeof(StashEntry) == 192, "stash entry layout");

enum { kCmdEvaluate = 0, kCmdDone = 1 };

// mrlm::minimize's variables for one object (lane 0 reads and writes them; `cmd`, `cur`, `cand`, `best` and the covariance
// in Hs are what the other lanes read, after a __syncwarp).
struct LMState {
    double x[kNP], best[kNP], grad[kNP], scale[kNP], diag[kNP], bs[kNP], step[kNP], delta[kNP], cand[kNP], z[kNP];
    double Hs[kNP * kNP], A[kNP * kNP], L[kNP * kNP], acc[kNAcc];
    double radius, decrease_factor, x_cost, x_norm, gradient_max_norm, minimum_cost, step_norm, model_cost_change, final_cost;
    int iteration, iterations, num_invalid, reuse_diagonal, step_is_successful, term, cost_evals, jac_evals;
    int cur;   // stash entry of the current point (cur ^ 1: the candidate's)
    int cmd, spd, need_final;
    mrlm::LMOptions opt;
};

// mrlm::minimize (lm_dense.cuh) from one cost evaluation to the next: `first` -- the totals of the start point are in
// stash[cur]; otherwise the candidate S.cand has just been evaluated into stash[cur ^ 1].  Returns kCmdEvaluate with the
// next candidate in S.cand, or kCmdDone with S.best / S.term / S.iterations / S.final_cost / S.cost_evals set as
// minimize() returns them.  Same statements in the same order as minimize(); the Jacobian evaluation after an accepted
// step is the candidate's stash entry.  Lane 0 only.
// Host-compilable: tests/harness/sixdof_host_harness.cpp runs it against mrlm::minimize on identical numbers (bit-equal).
MRLM_HD_NOINLINE int lm_advance(LMState& S, const StashEntry* stash, bool first) {
    constexpr int NP = kNP;
    MR6_ASSUME_SHARED(&S);
    MR6_ASSUME_SHARED(stash);
    const mrlm::LMOptions& opt = S.opt;
    auto fetch = [&](int e) {
        S.acc[0] = stash[e].cost;
#pragma unroll
        for (int i = 0; i < kNAcc - 1; ++i) S.acc[1 + i] = (double)stash[e].tot[i];
    };
    auto load_point = [&](bool initial) {
        S.x_cost = S.acc[0];
        S.gradient_max_norm = 0.0;
        for (int k = 0; k < NP; ++k) {
            S.grad[k] = S.acc[kAccG + k];
            S.gradient_max_norm = fmax(S.gradient_max_norm, fabs(S.grad[k]));
        }
        if (initial)
            for (int k = 0; k < NP; ++k) S.scale[k] = 1.0 / (1.0 + sqrt(S.acc[kAccH + mrlm::tri<NP>(k, k)]));
        for (int a = 0; a < NP; ++a) {
            S.bs[a] = S.scale[a] * S.grad[a];
            for (int b = a; b < NP; ++b) {
                const double h = S.acc[kAccH + mrlm::tri<NP>(a, b)] * S.scale[a] * S.scale[b];
                S.Hs[a * NP + b] = h;
                S.Hs[b * NP + a] = h;
            }
        }
        S.x_norm = 0.0;
        for (int k = 0; k < NP; ++k) S.x_norm += S.x[k] * S.x[k];
        S.x_norm = sqrt(S.x_norm);
    };
    bool finished = false;
    if (first) {
        S.radius = opt.initial_radius; S.decrease_factor = 2.0; S.reuse_diagonal = 0; S.num_invalid = 0;
        S.minimum_cost = 1.7976931348623157e308;
        S.term = mrlm::kFailure; S.iterations = 0; S.cost_evals = 1; S.jac_evals = 1; S.final_cost = 0.0;
        for (int k = 0; k < NP; ++k) S.best[k] = S.x[k];
        fetch(S.cur);
        if (!mrlm::all_finite(S.acc, kNAcc)) { S.final_cost = S.acc[0]; return kCmdDone; }
        load_point(true);
        S.iteration = 0; S.step_is_successful = 1; S.term = mrlm::kNoConvergence;
    } else {
        S.cost_evals++;
        const double c = stash[S.cur ^ 1].cost;
        const double cand_cost = isfinite(c) ? c : 1.7976931348623157e308;
        const double cost_change = S.x_cost - cand_cost;
        if (S.step_norm <= opt.parameter_tolerance * (S.x_norm + opt.parameter_tolerance)) {
            S.term = mrlm::kConvergence; finished = true;
        } else if (fabs(cost_change) <= opt.function_tolerance * S.x_cost) {
            S.term = mrlm::kConvergence; finished = true;
        } else {
            const double relative_decrease = cost_change / S.model_cost_change;
            if (relative_decrease > opt.min_relative_decrease) {  // HandleSuccessfulStep
                for (int k = 0; k < NP; ++k) S.x[k] = S.cand[k];
                S.cur ^= 1;
                fetch(S.cur);
                S.jac_evals++;
                if (!mrlm::all_finite(S.acc, kNAcc)) {
                    S.term = mrlm::kFailure; finished = true;
                } else {
                    load_point(false);
                    S.step_is_successful = 1;
                    const double q = 2.0 * relative_decrease - 1.0;  // LevenbergMarquardtStrategy::StepAccepted
                    S.radius = S.radius / fmax(1.0 / 3.0, 1.0 - q * q * q);
                    S.radius = fmin(opt.max_radius, S.radius);
                    S.decrease_factor = 2.0;
                    S.reuse_diagonal = 0;
                }
            } else {  // StepRejected
                S.radius /= S.decrease_factor; S.decrease_factor *= 2.0;
            }
        }
    }
    while (!finished) {
        if (S.step_is_successful && S.x_cost < S.minimum_cost) {
            S.minimum_cost = S.x_cost;
            for (int k = 0; k < NP; ++k) S.best[k] = S.x[k];
        }
        S.iterations = S.iteration;
        if (S.iteration >= opt.max_num_iterations) { S.term = mrlm::kNoConvergence; break; }
        if (S.step_is_successful && S.gradient_max_norm <= opt.gradient_tolerance) { S.term = mrlm::kConvergence; break; }
        if (S.radius <= opt.min_radius) { S.term = mrlm::kConvergence; break; }
        ++S.iteration;
        S.step_is_successful = 0;

        // LevenbergMarquardtStrategy::ComputeStep
        if (!S.reuse_diagonal)
            for (int k = 0; k < NP; ++k)
                S.diag[k] = fmin(fmax(S.Hs[k * NP + k], opt.min_lm_diagonal), opt.max_lm_diagonal);
        for (int i = 0; i < NP * NP; ++i) S.A[i] = S.Hs[i];
        for (int k = 0; k < NP; ++k) S.A[k * NP + k] += S.diag[k] / S.radius;
        const bool solved = mrlm::cholesky_factor<NP>(S.A, S.L) && mrlm::cholesky_backsolve<NP>(S.L, S.bs, S.step, S.z);
        S.reuse_diagonal = 1;
        bool step_is_valid = false;
        S.model_cost_change = 0.0;
        if (solved) {
            double lin = 0.0, quad = 0.0;  // -(J s)^T (r + J s / 2) with s = -y
            for (int a = 0; a < NP; ++a) S.step[a] = -S.step[a];
            for (int a = 0; a < NP; ++a) {
                double hs = 0.0;
                for (int b = 0; b < NP; ++b) hs += S.Hs[a * NP + b] * S.step[b];
                lin += S.step[a] * S.bs[a];
                quad += S.step[a] * hs;
            }
            S.model_cost_change = -(lin + 0.5 * quad);
            step_is_valid = S.model_cost_change > 0.0;
        }
        if (!step_is_valid) {  // HandleInvalidStep
            if (++S.num_invalid >= opt.max_consecutive_invalid) { S.term = mrlm::kFailure; break; }
            S.radius /= S.decrease_factor; S.decrease_factor *= 2.0;
            continue;
        }
        S.num_invalid = 0;
        double step_norm = 0.0;
        for (int k = 0; k < NP; ++k) {
            S.delta[k] = S.step[k] * S.scale[k];
            S.cand[k] = S.x[k] + S.delta[k];
            step_norm += S.delta[k] * S.delta[k];
        }
        S.step_norm = sqrt(step_norm);
        return kCmdEvaluate;
    }
    S.final_cost = S.minimum_cost;
    return kCmdDone;
}

#ifdef __CUDACC__

constexpr int kMixMaxWarps = 8;
enum { kWLogstd = 0, kWIstd = 1, kWFull = 2 };   // = MRPNP_W_*
__host__ __device__ inline int mixed_weight_bytes(int wkind) { return wkind == kWFull ? 12 : wkind == kWIstd ? 8 : 16; }
#ifndef MR6_MIX_UNROLL
#define MR6_MIX_UNROLL 1   // 2 measured 4 % slower (instruction cache; profiles/r02_6dof.txt)
#endif
constexpr int kMixUnroll = MR6_MIX_UNROLL;
#ifndef MR6_STAGE_ROWS
#define MR6_STAGE_ROWS 5
#endif
constexpr int kStageRows = MR6_STAGE_ROWS;

struct WarpArea {   // per warp, in front of the point planes
    StashEntry stash[2];
    Pose6 pose;
    LMState lm;
};
__host__ __device__ inline int mixed_cap(int n_pts) { return (n_pts + 3) & ~3; }
__host__ __device__ inline int mixed_slot_bytes(int n_pts, int wkind) {
    return (int)((sizeof(WarpArea) + 15) & ~(size_t)15) + mixed_cap(n_pts) * (20 + mixed_weight_bytes(wkind));
}

// Transposed reduction of 32 per-lane partial sums: lane L ends with the warp total of value L.  16+8+4+2+1 shuffles.
__device__ __forceinline__ float warp_reduce32_scatter(float v[32], int lane) {
#pragma unroll
    for (int half = 16, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    return v[0];
}

// make_pose with one reciprocal instead of 27 divisions in the rotation derivatives (they only feed the fp32 Jacobian here);
// R and t are computed by the same expressions as in make_pose, so the cost chain sees identical numbers.
__device__ __noinline__ void make_pose_fast(const double* x, Pose6* p) {
    MR6_ASSUME_SHARED(p);
    MR6_ASSUME_SHARED(x);
    const double w[3] = {x[0], x[1], x[2]};
    const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    p->t[0] = x[3]; p->t[1] = x[4]; p->t[2] = x[5];
    double W[9];
    skew(w, W);
    if (theta2 > 2.220446049250313e-16) {
        const double theta = sqrt(theta2), c = cos(theta), s = sin(theta), inv = 1.0 / theta, inv2 = 1.0 / theta2;
        const double k[3] = {w[0] * inv, w[1] * inv, w[2] * inv};
        double Kx[9];
        skew(k, Kx);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                p->R[i * 3 + j] = (i == j ? c : 0.0) + s * Kx[i * 3 + j] + (1.0 - c) * k[i] * k[j];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double col[3] = {(a == 0 ? 1.0 : 0.0) - p->R[a], (a == 1 ? 1.0 : 0.0) - p->R[3 + a],
                                   (a == 2 ? 1.0 : 0.0) - p->R[6 + a]};
            const double u[3] = {w[1] * col[2] - w[2] * col[1], w[2] * col[0] - w[0] * col[2],
                                 w[0] * col[1] - w[1] * col[0]};
            double U[9], M[9];
            skew(u, U);
#pragma unroll
            for (int i = 0; i < 9; ++i) M[i] = (w[a] * W[i] + U[i]) * inv2;
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

template <int WKIND>
struct MixedPass {
    static constexpr bool FULLW = WKIND == kWFull;
    const KParams& kp;
    Camera cam;
    const float* slot;   // shared memory: planes of `cap` entries x y z u v, then the weight planes
    WarpArea* area;      // shared memory: stash, pose matrices, minimiser arrays
    int cap, n, lane;

    __device__ __forceinline__ float at(const float* base, int c, int nc, int p) const {
        return kp.planar ? __ldg(base + (size_t)c * kp.n_pts + p) : __ldg(base + (size_t)p * nc + c);
    }

    // Compacts the object's inliers into the slot (point order kept); returns their number.
    __device__ __forceinline__ int stage(int obj, float* slot_w) {
        constexpr int wc = FULLW ? 3 : 2;
        const float* c3 = kp.coords_3d + (size_t)obj * 3 * kp.n_pts;
        const float* c2 = kp.coords_2d + (size_t)obj * 2 * kp.n_pts;
        const float* cw = kp.weights + (size_t)obj * wc * kp.n_pts;
        const uint32_t* mask = kp.inlier ? kp.inlier + (size_t)obj * ((kp.n_pts + 31) >> 5) : nullptr;
        float *sx = slot_w, *sy = slot_w + cap, *sz = slot_w + 2 * cap, *su = slot_w + 3 * cap, *sv = slot_w + 4 * cap;
        float* swf = slot_w + 5 * cap;
        double* swd = reinterpret_cast<double*>(slot_w + 5 * cap);
        int count = 0;
        // kStageRows rows of 32 points per round: all their loads are issued before the first ballot (the compiler does not
        // move loads across the warp votes by itself, and one row per round left the compaction waiting on memory 25 times
        // per object)
        for (int base0 = 0; base0 < kp.n_pts; base0 += 32 * kStageRows) {
            float X[kStageRows], Y[kStageRows], Z[kStageRows], u[kStageRows], v[kStageRows], w0[kStageRows], w1[kStageRows],
                w2[kStageRows];
            unsigned bits[kStageRows];
#pragma unroll
            for (int r = 0; r < kStageRows; ++r) {
                const int p = base0 + 32 * r + lane;
                const bool in = p < kp.n_pts;
                const int q = in ? p : 0;
                X[r] = at(c3, 0, 3, q); Y[r] = at(c3, 1, 3, q); Z[r] = at(c3, 2, 3, q);
                u[r] = at(c2, 0, 2, q); v[r] = at(c2, 1, 2, q);
                w0[r] = at(cw, 0, wc, q); w1[r] = at(cw, 1, wc, q); w2[r] = FULLW ? at(cw, 2, wc, q) : 0.f;
                const bool row = base0 + 32 * r < kp.n_pts;
                bits[r] = !row ? 0u : (mask ? __ldg(mask + ((base0 >> 5) + r)) : 0xffffffffu);
            }
#pragma unroll
            for (int r = 0; r < kStageRows; ++r) {
                const bool on = (base0 + 32 * r + lane < kp.n_pts) && ((bits[r] >> lane) & 1u);
                const unsigned m = __ballot_sync(0xffffffffu, on);
                if (on) {
                    const int j = count + __popc(m & ((1u << lane) - 1u));
                    sx[j] = X[r]; sy[j] = Y[r]; sz[j] = Z[r]; su[j] = u[r]; sv[j] = v[r];
                    if (WKIND == kWLogstd) {
                        swd[j] = exp(-(double)w0[r]) / kp.std_scale;
                        swd[cap + j] = exp(-(double)w1[r]) / kp.std_scale;
                    } else {
                        swf[j] = w0[r];
                        swf[cap + j] = w1[r];
                        if (FULLW) swf[2 * cap + j] = w2[r];
                    }
                }
                count += __popc(m);
            }
        }
        __syncwarp();
        return count;
    }

    // One evaluation at x: cost (fp64) and normal equations (fp32) of all staged points, totals into stash[e].
    __device__ __noinline__ void evaluate(const double* x, int e) const {
        Pose6* ps = &area->pose;
        __syncwarp();
        if (lane == 0) make_pose_fast(x, ps);
        __syncwarp();
        const double R0 = ps->R[0], R1 = ps->R[1], R2 = ps->R[2], R3 = ps->R[3], R4 = ps->R[4], R5 = ps->R[5], R6 = ps->R[6],
                     R7 = ps->R[7], R8 = ps->R[8], t0 = ps->t[0], t1 = ps->t[1], t2 = ps->t[2];
        float D[27];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int i = 0; i < 9; ++i) D[k * 9 + i] = (float)ps->dR[k][i];
        const double fx = cam.fx, fy = cam.fy, cx = cam.cx, cy = cam.cy, z_min = cam.z_min, u_min = cam.u_min,
                     u_max = cam.u_max, v_min = cam.v_min, v_max = cam.v_max;
        const float fxf = (float)fx, fyf = (float)fy;
        float2 a2[kNAcc - 1];   // .x: the first residual row's share, .y: the second's
#pragma unroll
        for (int i = 0; i < kNAcc - 1; ++i) a2[i] = make_float2(0.f, 0.f);
        double cost = 0.0;
        const int cap_ = cap, n_ = n;
        MR6_ASSUME_SHARED(slot);   // LDS instead of generic loads: 19 of the loop's 209 instructions
        MR6_ASSUME_SHARED(area);
        MR6_ASSUME_SHARED(x);
        const float *sx = slot, *sy = slot + cap_, *sz = slot + 2 * cap_, *su = slot + 3 * cap_, *sv = slot + 4 * cap_;
        const float* swf = slot + 5 * cap_;
        const double* swd = reinterpret_cast<const double*>(slot + 5 * cap_);
#pragma unroll kMixUnroll
        for (int j = lane; j < n_; j += 32) {
            const float Xf = sx[j], Yf = sy[j], Zf = sz[j];
            const double X = Xf, Y = Yf, Z = Zf;
            // fp64 residual chain: the expressions of add_point
            const double xc = R0 * X + R1 * Y + R2 * Z + t0;
            const double yc = R3 * X + R4 * Y + R5 * Z + t1;
            const double zc = R6 * X + R7 * Y + R8 * Z + t2;
            const bool z_free = !(zc < z_min);
            const double z = z_free ? zc : z_min, iz = 1.0 / z;
            double pu = fx * xc * iz + cx, pv = fy * yc * iz + cy;
            bool u_free = true, v_free = true;
            if (pu < u_min) { pu = u_min; u_free = false; } else if (pu > u_max) { pu = u_max; u_free = false; }
            if (pv < v_min) { pv = v_min; v_free = false; } else if (pv > v_max) { pv = v_max; v_free = false; }
            const double du = pu - (double)su[j], dv = pv - (double)sv[j];
            double w00, w01, w11;
            float f00, f01, f11;
            if (WKIND == kWLogstd) {
                w00 = swd[j]; w01 = 0.0; w11 = swd[cap_ + j];
                f00 = (float)w00; f01 = 0.f; f11 = (float)w11;
            } else {
                f00 = swf[j];
                f01 = FULLW ? swf[cap_ + j] : 0.f;
                f11 = FULLW ? swf[2 * cap_ + j] : swf[cap_ + j];
                w00 = f00; w01 = f01; w11 = f11;
            }
            const double r0 = w00 * du + w01 * dv, r1 = w01 * du + w11 * dv;
            cost += 0.5 * (r0 * r0 + r1 * r1);
            // fp32 Jacobian rows and normal equations
            const float izf = (float)iz, xcf = (float)xc, ycf = (float)yc, mz = z_free ? 1.f : 0.f;
            const float au = u_free ? fxf * izf : 0.f, bu = u_free ? -fxf * xcf * izf * izf * mz : 0.f;
            const float av = v_free ? fyf * izf : 0.f, bv = v_free ? -fyf * ycf * izf * izf * mz : 0.f;
            float j0[kNP], j1[kNP];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float* Dk = D + k * 9;
                const float dx = Dk[0] * Xf + Dk[1] * Yf + Dk[2] * Zf, dy = Dk[3] * Xf + Dk[4] * Yf + Dk[5] * Zf,
                            dz = Dk[6] * Xf + Dk[7] * Yf + Dk[8] * Zf;
                const float ju = au * dx + bu * dz, jv = av * dy + bv * dz;
                j0[k] = f00 * ju + f01 * jv;
                j1[k] = f01 * ju + f11 * jv;
            }
            j0[3] = f00 * au;            j1[3] = f01 * au;
            j0[4] = f01 * av;            j1[4] = f11 * av;
            j0[5] = f00 * bu + f01 * bv; j1[5] = f01 * bu + f11 * bv;
            const float r0f = (float)r0, r1f = (float)r1;
            // the two residual rows side by side: one packed FMA per sum (3.5 % faster than 54 scalar FMAs)
            const float2 rr = make_float2(r0f, r1f);
#pragma unroll
            for (int p = 0; p < kNP; ++p) {
                const float2 jp = make_float2(j0[p], j1[p]);
                a2[p] = __ffma2_rn(jp, rr, a2[p]);
#pragma unroll
                for (int q = p; q < kNP; ++q)
                    a2[kNP + mrlm::tri<kNP>(p, q)] = __ffma2_rn(jp, make_float2(j0[q], j1[q]), a2[kNP + mrlm::tri<kNP>(p, q)]);
            }
        }
        float a[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = i < kNAcc - 1 ? a2[i].x + a2[i].y : 0.f;
        const float tot = warp_reduce32_scatter(a, lane);
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, m);
        __syncwarp();
        StashEntry* s = area->stash + e;
        s->tot[lane] = tot;
        if (lane < kNP) s->x[lane] = x[lane];
        if (lane == 0) s->cost = cost;
        __syncwarp();
    }
};

template <int WKIND>
__global__ void __launch_bounds__(kMixMaxWarps * 32, 1) pnp_6dof_mixed_kernel(const KParams kp, int* counters, int slot_bytes) {
    extern __shared__ __align__(16) unsigned char mix_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = mix_smem + (size_t)warp * slot_bytes;
    MixedPass<WKIND> pass{kp};
    pass.cap = mixed_cap(kp.n_pts);
    pass.area = reinterpret_cast<WarpArea*>(mine);
    float* planes = reinterpret_cast<float*>(mine + ((sizeof(WarpArea) + 15) & ~(size_t)15));
    pass.slot = planes;
    pass.lane = lane;
    pass.cam.z_min = kp.z_min;
    LMState& S = pass.area->lm;
    const StashEntry* stash = pass.area->stash;
    if (lane == 0) {
        S.opt = mrlm::default_options();
        if (kp.max_iterations > 0) S.opt.max_num_iterations = kp.max_iterations;
    }
    while (true) {
        int obj = 0;
        if (lane == 0) obj = atomicAdd(counters, 1);
        obj = __shfl_sync(0xffffffffu, obj, 0);
        if (obj >= kp.n_obj) break;
#ifndef MR6_EXP_NO_PREFETCH
        {   // the object one generation of warps ahead: into L2 now, so that whoever claims it compacts from L2 instead of HBM
            const int ahead = obj + (int)gridDim.x * (int)(blockDim.x >> 5);
            if (ahead < kp.n_obj) {
                constexpr int wc = WKIND == kWFull ? 3 : 2;
                const char* b3 = reinterpret_cast<const char*>(kp.coords_3d + (size_t)ahead * 3 * kp.n_pts);
                const char* b2 = reinterpret_cast<const char*>(kp.coords_2d + (size_t)ahead * 2 * kp.n_pts);
                const char* bw = reinterpret_cast<const char*>(kp.weights + (size_t)ahead * wc * kp.n_pts);
                for (int off = lane * 128; off < 12 * kp.n_pts; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b3 + off));
                for (int off = lane * 128; off < 8 * kp.n_pts; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + off));
                for (int off = lane * 128; off < 4 * wc * kp.n_pts; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(bw + off));
            }
        }
#endif
        const float* K = kp.cam_mats + (size_t)obj * kp.cam_stride;
        const float* rg = kp.uv_range + (size_t)obj * kp.range_stride;
        pass.cam.fx = K[0]; pass.cam.fy = K[4]; pass.cam.cx = K[2]; pass.cam.cy = K[5];
        pass.cam.u_min = rg[0]; pass.cam.u_max = rg[1]; pass.cam.v_min = rg[2]; pass.cam.v_max = rg[3];
        __syncwarp();
        if (lane < kNP) S.x[lane] = kp.init[(size_t)obj * kNP + lane];
        if (lane == 0) S.cur = 0;
        pass.n = pass.stage(obj, planes);   // ends with __syncwarp
        pass.evaluate(S.x, 0);
        bool first = true;
        while (true) {
            if (lane == 0) S.cmd = lm_advance(S, stash, first);
            __syncwarp();
            if (S.cmd == kCmdDone) break;
            pass.evaluate(S.cand, S.cur ^ 1);
            first = false;
        }
        // covariance at the returned pose: its totals are the current stash entry (the returned pose is the last accepted
        // point) unless the minimiser failed before accepting anything
        if (lane == 0) {
            const StashEntry* e = stash + S.cur;
            bool same = true;
            for (int k = 0; k < kNP; ++k) same = same && (e->x[k] == S.best[k]);
            S.need_final = same ? 0 : 1;
        }
        __syncwarp();
        if (S.need_final) pass.evaluate(S.best, S.cur);
        if (lane == 0) {   // (J^T J) = L L^T, like mr6::covariance
            const StashEntry* e = stash + S.cur;
            for (int a = 0; a < kNP; ++a)
                for (int b = a; b < kNP; ++b) {
                    S.A[a * kNP + b] = (double)e->tot[kNP + mrlm::tri<kNP>(a, b)];
                    S.A[b * kNP + a] = S.A[a * kNP + b];
                }
            S.spd = mrlm::cholesky_factor<kNP>(S.A, S.L) ? 1 : 0;
        }
        __syncwarp();
        bool col_ok = true;
        if (lane < kNP && S.spd) {   // one column of the inverse per lane: the same back-substitutions as covariance()
            double e[kNP], y[kNP], z[kNP];
#pragma unroll
            for (int i = 0; i < kNP; ++i) e[i] = (i == lane) ? 1.0 : 0.0;
            col_ok = mrlm::cholesky_backsolve<kNP>(S.L, e, y, z);
#pragma unroll
            for (int i = 0; i < kNP; ++i) S.Hs[i * kNP + lane] = y[i];
        }
        const bool spd = S.spd && __all_sync(0xffffffffu, col_ok);
        __syncwarp();
        {   // rows of 48 doubles, written by the warp
            double* out = kp.result + (size_t)obj * kResultStride;
            const bool valid = (S.term == mrlm::kConvergence || S.term == mrlm::kNoConvergence);  // IsSolutionUsable
            const bool good = valid && spd;
            for (int i = lane; i < kResultStride; i += 32) {
                double v = 0.0;
                if (i < kNP) v = S.best[i];
                else if (i < kNP + kNP * kNP) v = good ? S.Hs[i - kNP] : (((i - kNP) % (kNP + 1) == 0) ? 1.0 : 0.0);
                else if (i == 42) v = good ? 1.0 : 0.0;   // Covariance::Compute failing clears result_val (cpp:287)
                else if (i == 43) v = S.iterations;
                else if (i == 44) v = S.final_cost;
                else if (i == 45) v = S.cost_evals;
                else if (i == 46) v = S.term;
                out[i] = v;
            }
        }
        __syncwarp();
    }
    // self-resetting work counters, as in pnp_kernel.cuh: the last CTA to finish rearms them for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int done = atomicAdd(counters + 1, 1);
        if (done == (int)gridDim.x - 1) {
            counters[0] = 0;
            counters[1] = 0;
            __threadfence();
        }
    }
}

#endif  // __CUDACC__

}  // namespace mr6
