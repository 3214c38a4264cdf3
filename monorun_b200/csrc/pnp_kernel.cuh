// pnp_kernel.cuh -- persistent warp-per-object LM kernel; T = double (MRPNP_PREC_FP64) or float (MRPNP_PREC_FP32).
#pragma once
#include "pnp_device.cuh"

namespace mrpnp {

// Ceres 1.14 Solver::Options defaults left untouched by the reference (pnp_uncert_cpu.cpp:270-271).
constexpr double kFunctionTol = 1e-6;
constexpr double kGradientTol = 1e-10;
constexpr double kParameterTol = 1e-8;
constexpr double kInitialRadius = 1e4;
constexpr double kMaxRadius = 1e16;
constexpr double kMinRadius = 1e-32;
constexpr double kMinRelDecrease = 1e-3;
constexpr double kMinLmDiag = 1e-6;
constexpr double kMaxLmDiag = 1e32;
constexpr int kMaxInvalidSteps = 5;
constexpr double kDblMax = 1.7976931348623157e308;

enum Termination { kConvergence = 0, kNoConvergence = 1, kFailure = 2 };

__device__ __forceinline__ bool finite_value(double v) { return fabs(v) < kDblMax; }  // false for NaN too
__device__ __forceinline__ bool finite_value(float v) { return fabsf(v) < 3.0e38f; }

// Stage one object's slab into the warp's slot.  TMA path: three 1-D bulk copies completing on the
// warp's mbarrier; fallback: coalesced loads through registers.
template <int WC>
__device__ __forceinline__ void load_object(const KParams& kp, int obj, float* slot, uint64_t* bar, uint32_t& parity,
                                            int lane) {
    const int P = kp.n_pts;
    const float* g3 = kp.c3d + (size_t)obj * 3 * P;
    const float* g2 = kp.c2d + (size_t)obj * 2 * P;
    const float* gw = kp.wgt + (size_t)obj * WC * P;
    float* s3 = slot;
    float* s2 = slot + 3 * P;
    float* sw = slot + 5 * P;
    __syncwarp();  // every lane is done with the previous object's data
    if (kp.use_tma) {
        if (lane == 0) {
            fence_proxy_async();  // order our generic-proxy accesses before the async-proxy writes
            mbar_expect_tx(bar, (uint32_t)((5 + WC) * P * sizeof(float)));
            bulk_g2s(s3, g3, (uint32_t)(3 * P * sizeof(float)), bar);
            bulk_g2s(s2, g2, (uint32_t)(2 * P * sizeof(float)), bar);
            bulk_g2s(sw, gw, (uint32_t)(WC * P * sizeof(float)), bar);
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
    } else {
        for (int i = lane; i < 3 * P; i += 32) s3[i] = __ldg(g3 + i);
        for (int i = lane; i < 2 * P; i += 32) s2[i] = __ldg(g2 + i);
        for (int i = lane; i < WC * P; i += 32) sw[i] = __ldg(gw + i);
        __syncwarp();
    }
}

// istd from log-std in place, per-axis mean, inlier decision, inlier_out, in-place compaction.
// Returns the number of active points (compacted inliers, or P) and the per-lane inlier bits.
template <int WMODE, int LAYOUT>
__device__ __forceinline__ int prepare_object(const KParams& kp, int obj, float* slot, int lane, uint32_t& bits,
                                              bool& compacted, int& n_inliers) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    constexpr int CV = WC - 1;  // channel of the v-axis weight (wyy for full W)
    const int P = kp.n_pts;
    float* s3 = slot;
    float* s2 = slot + 3 * P;
    float* sw = slot + 5 * P;
    float su = 0.f, sv = 0.f;
    for (int p = lane; p < P; p += 32) {
        float wu = sw[sidx<LAYOUT, WC>(p, 0, P)], wv = sw[sidx<LAYOUT, WC>(p, CV, P)];
        if (WMODE == MRPNP_W_LOGSTD) {  // uncert_prop_pnp_optimizer.py:73
            wu = __fdiv_rn(expf(-wu), kp.std_scale);
            wv = __fdiv_rn(expf(-wv), kp.std_scale);
            sw[sidx<LAYOUT, WC>(p, 0, P)] = wu;
            sw[sidx<LAYOUT, WC>(p, CV, P)] = wv;
        }
        su += wu;
        sv += wv;
    }
    su = warp_sum(su);
    sv = warp_sum(sv);
    // pnp_uncert_cpu.py:164-168: istd >= thres * mean on both axes (fp32 like numpy)
    const float thr_u = kp.istd_thres * __fdiv_rn(su, (float)P);
    const float thr_v = kp.istd_thres * __fdiv_rn(sv, (float)P);
    const uint8_t* gin = kp.inl_in ? kp.inl_in + (size_t)obj * P : nullptr;
    const bool test = kp.istd_thres > 0.f;
    bits = 0u;
    int cnt = 0;
    const int rows = (P + 31) >> 5;
    for (int k = 0; k < rows; ++k) {
        const int p = k * 32 + lane;
        bool inl = p < P;
        if (inl) {
            if (gin) inl = gin[p] != 0;
            else if (test) inl = sw[sidx<LAYOUT, WC>(p, 0, P)] >= thr_u && sw[sidx<LAYOUT, WC>(p, CV, P)] >= thr_v;
        }
        bits |= (inl ? 1u : 0u) << k;
        cnt += __popc(__ballot_sync(kFull, inl));
    }
    if (cnt <= 4) {  // pnp_uncert_cpu.py:23-32: too few inliers -> every point is an inlier
        cnt = P;
        bits = 0u;
        for (int k = 0; k < rows; ++k) bits |= ((k * 32 + lane) < P ? 1u : 0u) << k;
    }
    if (kp.inl_out) {
        uint8_t* gout = kp.inl_out + (size_t)obj * P;
        for (int k = 0; k < rows; ++k) {
            const int p = k * 32 + lane;
            if (p < P) gout[p] = (uint8_t)((bits >> k) & 1u);
        }
    }
    compacted = false;
    n_inliers = cnt;
    if (kp.inlier_opt_only && cnt < P) {  // pnp_uncert_cpu.py:62-66: LM sees the inliers only, order kept
        int base = 0;
        for (int k = 0; k < rows; ++k) {
            const int p = k * 32 + lane;
            const bool inl = (bits >> k) & 1u;
            float v3[3], v2[2], vw[WC];
            if (inl) {
#pragma unroll
                for (int c = 0; c < 3; ++c) v3[c] = s3[sidx<LAYOUT, 3>(p, c, P)];
#pragma unroll
                for (int c = 0; c < 2; ++c) v2[c] = s2[sidx<LAYOUT, 2>(p, c, P)];
#pragma unroll
                for (int c = 0; c < WC; ++c) vw[c] = sw[sidx<LAYOUT, WC>(p, c, P)];
            }
            const unsigned m = __ballot_sync(kFull, inl);
            __syncwarp();  // this row's reads (all lanes) happen before any lane's compacted writes
            if (inl) {
                const int d = base + __popc(m & ((1u << lane) - 1u));
#pragma unroll
                for (int c = 0; c < 3; ++c) s3[sidx<LAYOUT, 3>(d, c, P)] = v3[c];
#pragma unroll
                for (int c = 0; c < 2; ++c) s2[sidx<LAYOUT, 2>(d, c, P)] = v2[c];
#pragma unroll
                for (int c = 0; c < WC; ++c) sw[sidx<LAYOUT, WC>(d, c, P)] = vw[c];
            }
            base += __popc(m);
        }
        __syncwarp();
        compacted = true;
        return cnt;
    }
    __syncwarp();
    return P;
}

template <typename T, int WMODE, int LAYOUT>
__global__ void __launch_bounds__(kMaxThreads, 1) pnp_lm_kernel(const KParams kp) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw) + warp;
    float* slot = reinterpret_cast<float*>(smem_raw + kBarrierBytes) + (size_t)warp * kp.slot_floats;
    const int P = kp.n_pts;
    const float* s3 = slot;
    const float* s2 = slot + 3 * P;
    const float* sw = slot + 5 * P;

    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncwarp();
    uint32_t parity = 0;
    const int max_iter = kp.max_iter > 0 ? kp.max_iter : 50;

    while (true) {
        int obj = 0;
        if (lane == 0) obj = atomicAdd(kp.counters, 1);
        obj = __shfl_sync(kFull, obj, 0);
        if (obj >= kp.n_obj) break;

        load_object<WC>(kp, obj, slot, bar, parity, lane);

        Camera<T> cam;
        {
            const float* K = kp.cam + (size_t)obj * kp.cam_stride;
            const float* R = kp.range + (size_t)obj * kp.range_stride;
            cam.fx = (T)__ldg(K + 0); cam.fy = (T)__ldg(K + 4);  // pnp_uncert_cpu.cpp:265
            cam.cx = (T)__ldg(K + 2); cam.cy = (T)__ldg(K + 5);
            cam.z_min = (T)kp.z_min;
            cam.u_min = (T)__ldg(R + 0); cam.u_max = (T)__ldg(R + 1);
            cam.v_min = (T)__ldg(R + 2); cam.v_max = (T)__ldg(R + 3);
        }

        uint32_t bits;
        bool compacted;
        int n_inliers;
        const int n = prepare_object<WMODE, LAYOUT>(kp, obj, slot, lane, bits, compacted, n_inliers);

        // ---------------- Levenberg-Marquardt, Ceres 1.14 TrustRegionMinimizer control flow ----------------
        double x[4];
        bool init_ok = true;
        if (kp.init_mode == MRPNP_INIT_GIVEN) {
            const float* ip = kp.init + (size_t)obj * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = (double)__ldg(ip + i);
        } else {
            init_ok = linear_init<T, WMODE, LAYOUT>(s3, s2, sw, P, n, lane, cam, x);
            if (!init_ok) { x[0] = x[1] = x[2] = x[3] = 0.0; }  // pnp_uncert_cpu.py:119-125
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = (double)(float)x[i];  // fp32 hand-over, like the EPnP result
        }
        T acc[kNumAcc];
        bool clip_x;
        eval_pass<T, WMODE, LAYOUT, 0, false>(s3, s2, sw, P, n, lane, bits, x, cam, acc, clip_x);
        double cost = 0.5 * (double)acc[0];
        double g[4], H[10];  // unscaled gradient / Gauss-Newton matrix at x
#pragma unroll
        for (int i = 0; i < 4; ++i) g[i] = (double)acc[1 + i];
#pragma unroll
        for (int i = 0; i < 10; ++i) H[i] = (double)acc[5 + i];
        bool finite = finite_value(acc[0]);
#pragma unroll
        for (int i = 1; i < kNumAcc; ++i) finite = finite && finite_value(acc[i]);

        int term = kNoConvergence;
        int iteration = 0, cost_evals = 1;
        double radius = kInitialRadius;
        if (!finite || !init_ok) {
            term = kFailure;  // IterationZero failed: parameters stay at init
        } else {
            double scale[4];  // jacobi_scaling from the initial Jacobian only
#pragma unroll
            for (int i = 0; i < 4; ++i) scale[i] = 1.0 / (1.0 + sqrt(H[tri(i, i)]));
            double diag[4];
            double decrease_factor = 2.0;
            bool reuse_diagonal = false;
            int num_invalid = 0;
            double x_norm = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
            bool step_ok = true;  // iteration 0 counts as a successful step
            while (true) {
                // FinalizeIterationAndCheckIfMinimizerCanContinue
                if (iteration >= max_iter) { term = kNoConvergence; break; }
                if (step_ok) {
                    const double gmax = fmax(fmax(fabs(g[0]), fabs(g[1])), fmax(fabs(g[2]), fabs(g[3])));
                    if (gmax <= kGradientTol) { term = kConvergence; break; }
                }
                if (radius <= kMinRadius) { term = kConvergence; break; }
                ++iteration;
                step_ok = false;

                // LevenbergMarquardtStrategy::ComputeStep on the column-scaled system
                double Hs[10], gs[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    gs[i] = g[i] * scale[i];
#pragma unroll
                    for (int j = i; j < 4; ++j) Hs[tri(i, j)] = H[tri(i, j)] * scale[i] * scale[j];
                }
                if (!reuse_diagonal) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) diag[i] = fmin(fmax(Hs[tri(i, i)], kMinLmDiag), kMaxLmDiag);
                }
                reuse_diagonal = true;
                double A[10], L[10], y[4];
#pragma unroll
                for (int i = 0; i < 10; ++i) A[i] = Hs[i];
#pragma unroll
                for (int i = 0; i < 4; ++i) A[tri(i, i)] += diag[i] / radius;
                bool valid = chol4(A, L);
                double model_change = 0.0;
                if (valid) {
                    chol4_solve(L, gs, y);  // step = -y
                    // model_cost_change = -step^T gs - 1/2 step^T Hs step = y^T gs - 1/2 y^T Hs y
                    double hy[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        double s = 0.0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) s = fma(Hs[tri(i, j)], y[j], s);
                        hy[i] = s;
                    }
                    double yg = 0.0, yhy = 0.0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) { yg = fma(y[i], gs[i], yg); yhy = fma(y[i], hy[i], yhy); }
                    model_change = yg - 0.5 * yhy;
                    valid = (model_change > 0.0) && (fabs(y[0]) + fabs(y[1]) + fabs(y[2]) + fabs(y[3]) < kDblMax);
                }
                if (!valid) {  // HandleInvalidStep
                    if (++num_invalid >= kMaxInvalidSteps) { term = kFailure; break; }
                    radius /= decrease_factor;
                    decrease_factor *= 2.0;
                    continue;
                }
                num_invalid = 0;
                double delta[4], cand[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { delta[i] = -y[i] * scale[i]; cand[i] = x[i] + delta[i]; }

                // candidate cost, fused with its gradient / Gauss-Newton matrix (used only if accepted)
                bool clip_c;
                eval_pass<T, WMODE, LAYOUT, 0, false>(s3, s2, sw, P, n, lane, bits, cand, cam, acc, clip_c);
                ++cost_evals;
                const bool cfinite = finite_value(acc[0]);
                bool jfinite = cfinite;  // the fused pass also produced the candidate's Jacobian sums
#pragma unroll
                for (int i = 1; i < kNumAcc; ++i) jfinite = jfinite && finite_value(acc[i]);
                const double cand_cost = cfinite ? 0.5 * (double)acc[0] : kDblMax;

                // ParameterToleranceReached
                const double step_norm =
                    sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2] + delta[3] * delta[3]);
                if (step_norm <= kParameterTol * (x_norm + kParameterTol)) { term = kConvergence; break; }
                // FunctionToleranceReached (Ceres 1.14: candidate is not adopted on this exit)
                const double cost_change = cost - cand_cost;
                if (fabs(cost_change) <= kFunctionTol * cost) {
                    if (kp.adopt_ftol && cand_cost < cost) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) { x[i] = cand[i]; g[i] = (double)acc[1 + i]; }
#pragma unroll
                        for (int i = 0; i < 10; ++i) H[i] = (double)acc[5 + i];
                        cost = cand_cost;
                        clip_x = clip_c;
                    }
                    term = kConvergence;
                    break;
                }
                const double rho = cost_change / model_change;
                if (rho > kMinRelDecrease) {  // HandleSuccessfulStep
                    if (!jfinite) { term = kFailure; break; }  // re-evaluation at the new point fails in Ceres
#pragma unroll
                    for (int i = 0; i < 4; ++i) { x[i] = cand[i]; g[i] = (double)acc[1 + i]; }
#pragma unroll
                    for (int i = 0; i < 10; ++i) H[i] = (double)acc[5 + i];
                    cost = cand_cost;
                    clip_x = clip_c;
                    x_norm = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
                    step_ok = true;
                    const double q = 2.0 * rho - 1.0;
                    radius = radius / fmax(1.0 / 3.0, 1.0 - q * q * q);
                    radius = fmin(kMaxRadius, radius);
                    decrease_factor = 2.0;
                    reuse_diagonal = false;
                } else {  // HandleUnsuccessfulStep
                    radius /= decrease_factor;
                    decrease_factor *= 2.0;
                }
            }
        }
        bool usable = term != kFailure;  // Summary::IsSolutionUsable (pnp_uncert_cpu.cpp:276)

        // ---------------- pose covariance ----------------
        double cov[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) cov[i] = (i % 5 == 0) ? 1.0 : 0.0;
        if (kp.cov_mode != MRPNP_COV_NONE && usable) {
            // H already holds J^T J at the returned x with Ceres masks.  The pipeline covariance
            // (hessian.py:67-87) differs only when a point is clipped at x or outliers were kept in LM.
            const bool any_clip = __any_sync(kFull, clip_x);
            const bool need_pass = kp.cov_mode == MRPNP_COV_PIPELINE && (any_clip || (!compacted && n_inliers < P));
            if (need_pass) {
                bool c2;
                eval_pass<T, WMODE, LAYOUT, 1, true>(s3, s2, sw, P, n, lane, compacted ? 0xffffffffu : bits, x, cam,
                                                       acc, c2);
#pragma unroll
                for (int i = 0; i < 10; ++i) H[i] = (double)acc[5 + i];
            }
            if (!spd_inverse4(H, cov)) {  // pnp_uncert.py:79-85 fallback: H := I, object invalid
#pragma unroll
                for (int i = 0; i < 16; ++i) cov[i] = (i % 5 == 0) ? 1.0 : 0.0;
                usable = false;
            }
        }

        // ---------------- result row: one coalesced 96-byte store ----------------
        {
            float v = 0.f;  // unrolled selects: no dynamically indexed local arrays
#pragma unroll
            for (int i = 0; i < 4; ++i) v = (lane == i) ? (float)x[i] : v;
#pragma unroll
            for (int i = 0; i < 16; ++i) v = (lane == 4 + i) ? (float)cov[i] : v;
            v = (lane == 20) ? (usable ? 1.f : 0.f) : v;
            v = (lane == 21) ? (float)iteration : v;
            v = (lane == 22) ? (float)cost : v;
            v = (lane == 23) ? (float)radius : v;
            if (lane < MRPNP_RESULT_STRIDE) kp.result[(size_t)obj * MRPNP_RESULT_STRIDE + lane] = v;
            if (kp.result64) {
                double d = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) d = (lane == i) ? x[i] : d;
                d = (lane == 4) ? cost : d;
                d = (lane == 5) ? radius : d;
                d = (lane == 6) ? (double)cost_evals : d;
                d = (lane == 7) ? (double)term : d;
                if (lane < 8) kp.result64[(size_t)obj * 8 + lane] = d;
            }
        }
    }

    // self-resetting work counters: the last CTA to finish rearms them for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int done = atomicAdd(kp.counters + 1, 1);
        if (done == (int)gridDim.x - 1) {
            kp.counters[0] = 0;
            kp.counters[1] = 0;
            __threadfence();
        }
    }
}

}  // namespace mrpnp
