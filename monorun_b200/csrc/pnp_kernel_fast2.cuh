// pnp_kernel_fast2.cuh -- MRPNP_PREC_FAST with TWO warps per object (same arithmetic as pnp_kernel_fast.cuh).
//
// The 22 KB slab of an object limits an SM to ten objects; with one warp each that is 2.5 warps per scheduler, and
// every phase of the solve is a dependent chain (ncu: issue slots 53 % busy, "wait" the top stall).  Here each slot is
// worked on by a pair of warps:
//   * row-parallel phases (log-std -> weights, inlier mask + compaction, the fused passes) are split between the two
//     warps, each warp always revisiting the same rows (it owns their tracked residuals);
//   * after a pass both warps publish their 16 reduced sums in the pair's shared-memory header and meet at a named
//     barrier (bar.sync id, 64);
//   * the leader alone runs the fp32 trust-region algebra (state in its registers) and posts the next candidate step;
//     the follower sleeps at the barrier and costs no issue slots;
//   * both warps run the same code at the same time, so they also share instruction-cache lines.
#pragma once
#include "pnp_kernel_fast.cuh"

namespace mrpnp {

constexpr int kPairFastHeaderBytes = 512;
constexpr int kMaxPairsPerCta = 10;

enum PairCmd { kCmdFirst = 0, kCmdAnchor = 1, kCmdDelta = 2, kCmdUndoDelta = 3, kCmdDone = 4 };

struct PairHeader {
    uint64_t bar;          // staging mbarrier
    int cmd, obj;
    int seg_n[2];          // inliers per compaction segment
    float wsum[2][2];      // per-warp weight sums (u, v)
    int flagged[2];
    float pad0[2];
    DeltaStep step;        // candidate step of the next delta pass / the roll-back            (40 B, offset 56 -> 96)
    float pt[4];           // evaluation point of the next pass from the observations
    float sn, cs;
    float pad1[2];
    float scratch[24];     // leader: totals broadcast, H / x staging for the covariance        (offset 128)
    float part[2][16];     // per-warp reduced sums of the last pass
};
static_assert(sizeof(PairHeader) <= kPairFastHeaderBytes, "pair header too small");

__device__ __forceinline__ void pair_sync(int pair) {
    asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
}

// weights -> inverse std, rows dealt alternately to the two warps; returns this warp's share of the per-axis sums
template <int WMODE, int LAYOUT>
__device__ __forceinline__ void pair_weights(const KParams& kp, float* sw, int P, int lane, int t, float& su, float& sv) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    constexpr int CV = WC - 1;
    const float k2 = -1.4426950408889634f, off = -__log2f(kp.std_scale);
    su = 0.f; sv = 0.f;
#pragma unroll 2
    for (int p = lane + 32 * t; p < P; p += 64) {
        float wu = sw[sidx<LAYOUT, WC>(p, 0, P)], wv = sw[sidx<LAYOUT, WC>(p, CV, P)];
        if (WMODE == MRPNP_W_LOGSTD) {
            wu = fast_ex2(fmaf(wu, k2, off));
            wv = fast_ex2(fmaf(wv, k2, off));
            sw[sidx<LAYOUT, WC>(p, 0, P)] = wu;
            sw[sidx<LAYOUT, WC>(p, CV, P)] = wv;
        }
        su += wu;
        sv += wv;
    }
    su = warp_sum(su);
    sv = warp_sum(sv);
}

// fast_mask_and_compact restricted to rows [row_lo, row_hi): the inliers of these rows are compacted, in order, into
// the segment that starts at point row_lo * 32.  Returns their number.
template <int WMODE, int LAYOUT>
__device__ __forceinline__ int pair_mask_and_compact(const KParams& kp, int obj, float* slot, int P, int lane, float thr_u,
                                                     float thr_v, bool all_inliers, int row_lo, int row_hi) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    constexpr int CV = WC - 1;
    constexpr int U = 2;
    float* s3 = slot;
    float* s2 = slot + 3 * P;
    float* sw = slot + 5 * P;
    const int rows = (P + 31) >> 5;
    const bool test = kp.istd_thres > 0.f;
    uint32_t in_word = 0u;
    if (kp.inl_in && lane < rows) in_word = __ldg(kp.inl_in + (size_t)obj * rows + lane);
    uint32_t out_word = 0u;
    int base = row_lo * 32;
#pragma unroll 1
    for (int k0 = row_lo; k0 < row_hi; k0 += U) {
        float v3[U][3], v2[U][2], wv_[U][3];
        bool inl[U];
        int dst[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = k0 + u;
            const int p = k * 32 + lane;
            bool ok = k < row_hi && p < P;
            const int pc = ok ? p : 0;
            wv_[u][0] = sw[sidx<LAYOUT, WC>(pc, 0, P)];
            wv_[u][2] = sw[sidx<LAYOUT, WC>(pc, CV, P)];
            wv_[u][1] = (WMODE == MRPNP_W_FULL) ? sw[sidx<LAYOUT, WC>(pc, 1, P)] : 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) v3[u][c] = s3[sidx<LAYOUT, 3>(pc, c, P)];
#pragma unroll
            for (int c = 0; c < 2; ++c) v2[u][c] = s2[sidx<LAYOUT, 2>(pc, c, P)];
            if (!all_inliers) {
                if (kp.inl_in) {
                    const uint32_t row_word = __shfl_sync(kFull, in_word, k & 31);  // every lane takes part
                    ok = ok && ((row_word >> lane) & 1u);
                } else if (test) {
                    ok = ok && (wv_[u][0] >= thr_u) && (wv_[u][2] >= thr_v);
                }
            }
            const unsigned m = __ballot_sync(kFull, ok);
            if (lane == k && k < row_hi) out_word = m;
            inl[u] = ok;
            dst[u] = base + __popc(m & ((1u << lane) - 1u));
            base += __popc(m);
        }
        if (!all_inliers) {
            __syncwarp();  // the reads of these rows (all lanes) happen before any lane's compacted writes
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (inl[u]) {
                    const int d = dst[u];
#pragma unroll
                    for (int c = 0; c < 3; ++c) s3[sidx<LAYOUT, 3>(d, c, P)] = v3[u][c];
#pragma unroll
                    for (int c = 0; c < 2; ++c) s2[sidx<LAYOUT, 2>(d, c, P)] = v2[u][c];
                    sw[sidx<LAYOUT, WC>(d, 0, P)] = wv_[u][0];
                    if (WMODE == MRPNP_W_FULL) sw[sidx<LAYOUT, WC>(d, 1, P)] = wv_[u][1];
                    sw[sidx<LAYOUT, WC>(d, CV, P)] = wv_[u][2];
                }
            }
        }
    }
    __syncwarp();
    if (kp.inl_out && lane >= row_lo && lane < row_hi) kp.inl_out[(size_t)obj * rows + lane] = out_word;
    return base - row_lo * 32;
}

template <int WMODE, int LAYOUT>
__device__ __noinline__ bool pair_linear_init(const KParams& kp, int obj, float* slot, int n0, int n1, int base1, int lane,
                                              float* scratch, float* x_out) {
    const int P = kp.n_pts;
    const Camera<float> cam = load_camera<float>(kp, obj);
    double x[4];
    const bool ok = linear_init_impl<WMODE, LAYOUT>(slot, slot + 3 * P, slot + 5 * P, P, TeamRows(n0, n1, base1, 0, 1), lane,
                                                    cam, scratch, x);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) x_out[i] = ok ? (float)x[i] : 0.f;
    return ok;
}

// Leader: take the next object off the work counter, start its bulk copies, publish the index.
template <int WC>
__device__ __forceinline__ void pair_fetch_and_stage(const KParams& kp, float* slot, int P, PairHeader* hd, int lane) {
    if (lane == 0) {
        int obj = atomicAdd(kp.counters, 1);
        if (obj >= kp.n_obj) obj = -1;
        if (obj >= 0 && kp.use_tma) {
            const float *g3, *g2, *gw;
            object_slabs<WC>(kp, obj, g3, g2, gw);
            fence_proxy_async();
            if (kp.dense) {
                mbar_expect_tx(&hd->bar, (uint32_t)(5 * P * sizeof(float)));
                bulk_g2s(slot, g3, (uint32_t)(3 * P * sizeof(float)), &hd->bar);
                bulk_g2s(slot + 5 * P, gw, (uint32_t)(2 * P * sizeof(float)), &hd->bar);
            } else {
                mbar_expect_tx(&hd->bar, (uint32_t)((5 + WC) * P * sizeof(float)));
                bulk_g2s(slot, g3, (uint32_t)(3 * P * sizeof(float)), &hd->bar);
                bulk_g2s(slot + 3 * P, g2, (uint32_t)(2 * P * sizeof(float)), &hd->bar);
                bulk_g2s(slot + 5 * P, gw, (uint32_t)(WC * P * sizeof(float)), &hd->bar);
            }
        }
        hd->obj = obj;
    }
}

template <int WMODE, int LAYOUT, int PCT>
__global__ void __launch_bounds__(kMaxPairsPerCta * 64, 1) pnp_lm_fast2_kernel(const __grid_constant__ KParams kp) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    constexpr int R = 1;  // rows per loop iteration of a pass: the second warp of the pair provides the overlap
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = warp >> 1, t = warp & 1;
    const int npairs = blockDim.x >> 6;
    PairHeader* hd = reinterpret_cast<PairHeader*>(smem_raw + (size_t)pair * kPairFastHeaderBytes);
    float* slot = reinterpret_cast<float*>(smem_raw + (size_t)npairs * kPairFastHeaderBytes) + (size_t)pair * kp.slot_floats;
    float* scratch = hd->scratch;
    const int P = PCT ? PCT : kp.n_pts;
    float* s3 = slot;
    float* s2 = slot + 3 * P;
    float* sw = slot + 5 * P;
    const bool leader = t == 0;
    const int max_iter = kp.max_iter > 0 ? kp.max_iter : (kp.max_iter < 0 ? 0 : 50);
    const int rows_all = (P + 31) >> 5;
    const int rows_half = (rows_all + 1) >> 1;
    const int row_lo = min(rows_all, t * rows_half), row_hi = min(rows_all, row_lo + rows_half);

    if (leader && lane == 0) {
        mbar_init(&hd->bar, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncwarp();
    uint32_t parity = 0;
    if (leader) pair_fetch_and_stage<WC>(kp, slot, P, hd, lane);
    pair_sync(pair);
    int obj = hd->obj;

#pragma unroll 1
    while (obj >= 0) {
        const Camera<float> camf = load_camera<float>(kp, obj);
        // ---------------- stage + weights + inlier mask + compaction ----------------
        int n0 = 0, n1 = 0;
#pragma unroll 1
        for (int attempt = 0; attempt < 2; ++attempt) {
            if (kp.use_tma) {
                if (attempt) {  // re-stage the same object (the first attempt compacted the slot)
                    fence_proxy_async();
                    pair_sync(pair);
                    if (leader && lane == 0) {
                        const float *g3, *g2, *gw;
                        object_slabs<WC>(kp, obj, g3, g2, gw);
                        fence_proxy_async();
                        if (kp.dense) {
                            mbar_expect_tx(&hd->bar, (uint32_t)(5 * P * sizeof(float)));
                            bulk_g2s(s3, g3, (uint32_t)(3 * P * sizeof(float)), &hd->bar);
                            bulk_g2s(sw, gw, (uint32_t)(2 * P * sizeof(float)), &hd->bar);
                        } else {
                            mbar_expect_tx(&hd->bar, (uint32_t)((5 + WC) * P * sizeof(float)));
                            bulk_g2s(s3, g3, (uint32_t)(3 * P * sizeof(float)), &hd->bar);
                            bulk_g2s(s2, g2, (uint32_t)(2 * P * sizeof(float)), &hd->bar);
                            bulk_g2s(sw, gw, (uint32_t)(WC * P * sizeof(float)), &hd->bar);
                        }
                    }
                }
                mbar_wait(&hd->bar, parity);
                parity ^= 1u;
            } else {
                pair_sync(pair);
                if (leader) stage_object_plain<WC>(kp, obj, slot, lane);
                pair_sync(pair);
            }
            float su, sv;
            if (WMODE == MRPNP_W_LOGSTD && LAYOUT == MRPNP_LAYOUT_PLANAR && kp.dense) {
                // fused head prologue: the leader decodes the whole slot (thresholds come back through the scratch)
                if (leader) fast_dense_decode(kp, obj, slot, lane, scratch);
                pair_sync(pair);
                su = scratch[0]; sv = scratch[1];
            } else {
                pair_weights<WMODE, LAYOUT>(kp, sw, P, lane, t, su, sv);
                if (lane == 0) { hd->wsum[t][0] = su; hd->wsum[t][1] = sv; }
                pair_sync(pair);
                const float invP = 1.f / (float)P;
                su = kp.istd_thres * ((hd->wsum[0][0] + hd->wsum[1][0]) * invP);
                sv = kp.istd_thres * ((hd->wsum[0][1] + hd->wsum[1][1]) * invP);
            }
            // second attempt == pnp_uncert_cpu.py:28-32: <= 4 inliers -> every point is an inlier (slot re-staged)
            const bool all = attempt == 1;
            const int cnt = pair_mask_and_compact<WMODE, LAYOUT>(kp, obj, slot, P, lane, su, sv, all, row_lo, row_hi);
            if (lane == 0) hd->seg_n[t] = cnt;
            pair_sync(pair);
            n0 = hd->seg_n[0];
            n1 = hd->seg_n[1];
            if (all || n0 + n1 > 4) break;
        }
        const RowMap<2> rows(n0, n1, rows_half * 32, t);

        // ---------------- leader: initial point, LM state ----------------
        float x[4], pt[4];
        bool init_ok = true;
        float sn_x = 0.f, cs_x = 1.f, sn_p = 0.f, cs_p = 1.f;
        if (leader) {
            if (kp.init_mode == MRPNP_INIT_GIVEN) {
                const float* ip = kp.init + (size_t)obj * 4;
#pragma unroll
                for (int i = 0; i < 4; ++i) pt[i] = __ldg(ip + i);
            } else {
                init_ok = pair_linear_init<WMODE, LAYOUT>(kp, obj, slot, n0, n1, rows_half * 32, lane, scratch, pt);
            }
            sincos_cold(pt[0], &sn_x, &cs_x);
            cs_x += 1.f;
            sn_p = sn_x; cs_p = cs_x;
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) hd->pt[i] = pt[i];
                hd->sn = sn_p; hd->cs = cs_p;
                hd->cmd = kCmdFirst;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = pt[i];
        float cost = 0.f, g[4], H[10], scale[4], diag[4], delta[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { g[i] = 0.f; scale[i] = 1.f; diag[i] = 1.f; delta[i] = 0.f; }
#pragma unroll
        for (int i = 0; i < 10; ++i) H[i] = 0.f;
        int term = kNoConvergence, iteration = 0, cost_evals = 0, num_invalid = 0;
        float radius = (float)kInitialRadius, decrease_factor = 2.f, x_norm = 0.f, model_change = 1.f;
        bool reuse_diagonal = false, step_ok = true, first = true, redo = false;
        DeltaStep dstep = {};
        const ClipWindow cwin = make_clip_window(camf);

#pragma unroll 1
        while (true) {
            pair_sync(pair);  // A: the leader's command (and its parameters) are visible
            const int cmd = hd->cmd;
            if (cmd == kCmdDone) break;
            // ---- the fused pass over this warp's rows ----
            float a[16];
            bool flagged;
            if (cmd <= kCmdAnchor) {
                const float ptf[4] = {hd->pt[0], hd->pt[1], hd->pt[2], hd->pt[3]};
                eval_pass_first<WMODE, LAYOUT, R>(s3, s2, sw, P, rows, lane, cmd == kCmdAnchor, ptf, hd->sn, hd->cs, camf, a,
                                                  flagged);
            } else {
                const DeltaStep ds = hd->step;
                if (cmd == kCmdUndoDelta) {
                    // roll the rejected candidate back (its step is still in this thread's registers), then evaluate
                    undo_pass_delta<WMODE, LAYOUT>(slot, P, rows, lane, dstep, camf.fx, camf.fy);
                }
                dstep = ds;
                eval_pass_delta<WMODE, LAYOUT, R>(s3, s2, sw, P, rows, lane, ds, camf, cwin, a, flagged);
            }
            if (cmd == kCmdAnchor) dstep = hd->step;  // the anchor candidate's step, should it be rejected
            const float tot = warp_reduce16_scatter(a, lane);  // lane L: total of sum (L >> 1)
            if ((lane & 1) == 0) hd->part[t][lane >> 1] = tot;
            if (lane == 0) hd->flagged[t] = flagged ? 1 : 0;
            fence_proxy_async();  // this thread's stores to the slot precede a later bulk copy into it
            pair_sync(pair);  // B: both warps' sums are visible
            if (!leader) continue;

            // ================= leader only: combine, then Ceres' trust-region logic in fp32 =================
            const float both = lane < 16 ? hd->part[0][lane] + hd->part[1][lane] : 0.f;
            const bool jfin = __all_sync(kFull, fabsf(both) < kFltMax);
            if (lane < 16) scratch[lane] = both;
            __syncwarp();
            const bool from_observations = cmd <= kCmdAnchor;
            bool done = true;   // every `break` of the block below ends the object
            bool undo = false;
            do {
                if (hd->flagged[0] | hd->flagged[1]) { redo = true; break; }
                ++cost_evals;
                const float c_term = scratch[14];  // first two evaluations: sum |r|^2; afterwards: its change
                const bool cfinite = fabsf(c_term) < kFltMax;
                const bool jfinite = jfin && cfinite;
                bool accept = false;
                if (first) {  // IterationZero
                    first = false;
                    if (!jfinite || !init_ok) { term = kFailure; break; }  // parameters stay at init
                    accept = true;
                    cost = 0.5f * c_term;
                } else {
                    // ParameterToleranceReached
                    const float step_norm2 = delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2] + delta[3] * delta[3];
                    const float ptol = (float)kParameterTol * (x_norm + (float)kParameterTol);
                    if (step_norm2 <= ptol * ptol) { term = kConvergence; break; }
                    // FunctionToleranceReached (Ceres 1.14: the candidate is not adopted on this exit)
                    const float cost_change = !cfinite ? -kFltMax : (from_observations ? cost - 0.5f * c_term : -0.5f * c_term);
                    bool stop_after = false;
                    if (fabsf(cost_change) <= (float)kFunctionTol * cost) {
                        term = kConvergence;
                        if (!(kp.adopt_ftol && cost_change > 0.f)) break;
                        stop_after = true;  // documented switch: take the candidate, then stop
                    }
                    const float rho = cost_change * fast_rcp(model_change);
                    if (stop_after || rho > (float)kMinRelDecrease) {  // HandleSuccessfulStep
                        if (!jfinite) { term = kFailure; break; }
                        accept = true;
                        cost = from_observations ? 0.5f * c_term : cost - cost_change;
                        const float q = 2.f * rho - 1.f;
                        radius = fminf((float)kMaxRadius, radius * fast_rcp(fmaxf(1.f / 3.f, 1.f - q * q * q)));
                        decrease_factor = 2.f;
                        reuse_diagonal = false;
                        if (stop_after) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) x[i] = pt[i];
#pragma unroll
                            for (int i = 0; i < 10; ++i) H[i] = scratch[4 + i];
                            break;
                        }
                    } else {  // HandleUnsuccessfulStep: both warps roll the candidate back before the next pass
                        radius = radius * fast_rcp(decrease_factor);
                        decrease_factor *= 2.f;
                        undo = true;
                    }
                }
                if (accept) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { x[i] = pt[i]; g[i] = scratch[i]; }
#pragma unroll
                    for (int i = 0; i < 10; ++i) H[i] = scratch[4 + i];
                    sn_x = sn_p; cs_x = cs_p;
                    x_norm = fast_sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
                    step_ok = true;
                    if (cost_evals == 1) {  // jacobi_scaling from the initial Jacobian only
#pragma unroll
                        for (int i = 0; i < 4; ++i) scale[i] = fast_rcp(1.f + fast_sqrtf(H[tri(i, i)]));
                    }
                }
                // ---- next trust-region step (invalid steps shrink the radius without a new evaluation) ----
                bool stop = false;
#pragma unroll 1
                while (true) {
                    // FinalizeIterationAndCheckIfMinimizerCanContinue
                    if (iteration >= max_iter) { term = kNoConvergence; stop = true; break; }
                    if (step_ok) {
                        const float gmax = fmaxf(fmaxf(fabsf(g[0]), fabsf(g[1])), fmaxf(fabsf(g[2]), fabsf(g[3])));
                        if (gmax <= (float)kGradientTol) { term = kConvergence; stop = true; break; }
                    }
                    if (radius <= (float)kMinRadius) { term = kConvergence; stop = true; break; }
                    ++iteration;
                    step_ok = false;
                    // LevenbergMarquardtStrategy::ComputeStep on the column-scaled system
                    float A[10], gs[4], y[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        gs[i] = g[i] * scale[i];
#pragma unroll
                        for (int j = i; j < 4; ++j) A[tri(i, j)] = H[tri(i, j)] * (scale[i] * scale[j]);
                    }
                    if (!reuse_diagonal) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) diag[i] = fminf(fmaxf(A[tri(i, i)], (float)kMinLmDiag), (float)kMaxLmDiag);
                    }
                    reuse_diagonal = true;
                    const float inv_radius = fast_rcp(radius);
                    float dmp[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        dmp[i] = fminf(diag[i] * inv_radius, 1e30f);
                        A[tri(i, i)] += dmp[i];
                    }
                    const Ldl4f f = ldl4f_factor(A);
                    bool valid = f.ok;
                    if (valid) {
                        ldl4f_solve(f, gs, y);  // step = -y
                        float yg = 0.f, ydy = 0.f;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            yg = fmaf(y[i], gs[i], yg);
                            ydy = fmaf(dmp[i] * y[i], y[i], ydy);
                        }
                        model_change = 0.5f * (yg + ydy);
                        valid = (model_change > 0.f) && ((fabsf(y[0]) + fabsf(y[1])) + (fabsf(y[2]) + fabsf(y[3])) < kFltMax);
                    }
                    if (valid) {
                        num_invalid = 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) { delta[i] = -y[i] * scale[i]; pt[i] = x[i] + delta[i]; }
                        break;
                    }
                    // HandleInvalidStep
                    if (++num_invalid >= kMaxInvalidSteps) { term = kFailure; stop = true; break; }
                    radius = radius * fast_rcp(decrease_factor);
                    decrease_factor *= 2.f;
                }
                if (stop) break;
                done = false;  // a new candidate is ready
            } while (false);

            if (done) {
                if (lane == 0) hd->cmd = kCmdDone;
                continue;  // -> barrier A, where both warps leave the loop
            }
            // rotation at the candidate by angle addition; the same sin / cos - 1 of the yaw step drive the delta pass
            float sd, cdm1;
            sincos_cm1(delta[0], sd, cdm1);
            sn_p = fmaf(sn_x, cdm1, fmaf(cs_x, sd, sn_x));
            cs_p = fmaf(cs_x, cdm1, fmaf(-sn_x, sd, cs_x));
            if (lane == 0) {
                DeltaStep ds;
                ds.cp = cs_p; ds.sp = sn_p;
                ds.txp = pt[1]; ds.typ = pt[2]; ds.tzp = pt[3];
                ds.ncdm1 = -cdm1; ds.sd = sd;
                ds.dtx = delta[1]; ds.dty = delta[2]; ds.dtz = delta[3];
                hd->step = ds;
#pragma unroll
                for (int i = 0; i < 4; ++i) hd->pt[i] = pt[i];
                hd->sn = sn_p; hd->cs = cs_p;
                hd->cmd = cost_evals == 1 ? kCmdAnchor : (undo ? kCmdUndoDelta : kCmdDelta);
            }
        }
        // ---------------- the slot is free: the leader starts staging the next object, then finishes this one ----------------
        if (leader) {
            pair_fetch_and_stage<WC>(kp, slot, P, hd, lane);
            if (redo) {
                if (lane == 0) kp.redo_list[atomicAdd(kp.redo_count, 1)] = obj;
            } else if (lane == 0) {  // every lane holds the same H and x
#pragma unroll
                for (int i = 0; i < 10; ++i) scratch[i] = H[i];
#pragma unroll
                for (int i = 0; i < 4; ++i) scratch[12 + i] = x[i];
            }
        }
        pair_sync(pair);  // C: the next object's index is visible to the follower
        const int next_obj = hd->obj;
        if (leader && !redo)
            fast_finish_object(kp, obj, lane, scratch, scratch + 12, cost, radius, iteration, cost_evals, term);
        obj = next_obj;
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
