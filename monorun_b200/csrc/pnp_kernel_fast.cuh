// pnp_kernel_fast.cuh -- MRPNP_PREC_FAST kernel: warp per object, tracked residuals, fp32 delta passes, fp32 scalar LM,
// and a hot path small enough for the SM's instruction cache.
//
// What bounds the warp-per-object kernels of pnp_kernel.cuh on B200 is neither HBM nor the math pipes but instruction
// FETCH: each of the ten resident warps of an SM is in a different phase of its own object (staging, compaction,
// a pass, the 4x4 algebra), the fp64 trust-region algebra alone is ~36 KB of straight-line SASS, and the whole kernel
// is 64-120 KB against a 32 KB L1.5 / 6 KB L0 instruction cache (ncu: sm__icc_request_hit_rate 82 %, GPC instruction
// requests at 55 % of peak, "no_instruction" + fetch-bound "wait" the top stalls; tools/trace_run.py: a candidate
// evaluation costs 3.5 k cycles of FIXED time and only 26 cycles per row of 32 points).  So this kernel is written
// for code size first:
//   * residuals are evaluated once in fp64 and then tracked incrementally (pnp_fast.cuh), so the pass that runs
//     3-4 times per object is ~250 fp32 instructions in total;
//   * the trust-region algebra (Ceres 1.14 control flow, unchanged) is fp32: the normal equations come from fp32 sums
//     anyway (their error, amplified by the conditioning, already bounds the step accuracy), Jacobi scaling keeps
//     the 4x4 system well inside fp32 range, and the accept / function-tolerance decisions read the cost CHANGE
//     straight from the delta pass (relative error ~1e-6 of itself);
//   * everything cold (linear initialiser, fused head prologue, unaligned staging, roll-back of a rejected step, the
//     fp64 covariance) is out of line;
//   * loops are not unrolled beyond what latency hiding needs.
// Objects with a point near a clip bound go to the redo list for the exact kernel (see pnp_fast.cuh).
#pragma once
#include "pnp_kernel.cuh"
#include "pnp_fast.cuh"

// Phase trace (tools/trace_run.py, build with -DMRPNP_TRACE): clock64 ticks per phase, summed per object, written to
// the result64 buffer viewed as [N,32] doubles.  Phases: 0 staging wait, 1 weights + mask + compaction, 2 initialiser,
// 3 first evaluation, 4 candidate evaluations, 5 scalar trust-region algebra, 6 roll-backs, 7 covariance + stores.
#ifdef MRPNP_TRACE
#define TR_DECL long long tr_t = clock64(), tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long tr_t0 = tr_t;
#define TR_MARK(k) { const long long tr_now = clock64(); tr_acc[k] += tr_now - tr_t; tr_t = tr_now; }
#else
#define TR_DECL
#define TR_MARK(k)
#endif

namespace mrpnp {

#ifndef MRPNP_FAST_WARPS
#define MRPNP_FAST_WARPS 10
#endif
constexpr int kFastMaxWarps = MRPNP_FAST_WARPS;   // resident warps (= objects in flight) per SM
constexpr int kFastHeaderBytes = 128;  // per warp: mbarrier (8 B) + 24-float scratch at +16 (reduction broadcast)
constexpr int kFastScratch = 16;       // byte offset of the scratch

// Global addresses of one object's three slabs (for the fused head entry: the class slice of the head's full output,
// FCNNOCDecoder.slice_pred, fcn_noc_decoder.py:242-267, and `rois` riding in the coords_2d slot).
template <int WC>
__device__ __forceinline__ void object_slabs(const KParams& kp, int obj, const float*& g3, const float*& g2,
                                             const float*& gw) {
    const int P = kp.n_pts;
    g3 = kp.c3d + (size_t)obj * 3 * P;
    g2 = kp.c2d + (size_t)obj * (kp.dense ? 0 : 2 * P);
    gw = kp.wgt + (size_t)obj * WC * P;
    if (kp.dense && kp.pred_stride) {
        const long long c = kp.labels ? __ldg(kp.labels + obj) : 0;
        g3 = kp.c3d + (size_t)obj * kp.pred_stride + (size_t)(3 * c) * P;
        gw = kp.wgt + (size_t)obj * kp.pred_stride + (size_t)(2 * c) * P;
    }
}

// Unaligned shapes (P % 4 != 0 or unaligned pointers): coalesced loads through registers instead of bulk copies.
template <int WC>
__device__ __noinline__ void stage_object_plain(const KParams& kp, int obj, float* slot, int lane) {
    const int P = kp.n_pts;
    const float *g3, *g2, *gw;
    object_slabs<WC>(kp, obj, g3, g2, gw);
    for (int i = lane; i < 3 * P; i += 32) slot[i] = __ldg(g3 + i);
    if (!kp.dense) for (int i = lane; i < 2 * P; i += 32) slot[3 * P + i] = __ldg(g2 + i);
    for (int i = lane; i < WC * P; i += 32) slot[5 * P + i] = __ldg(gw + i);
    __syncwarp();
}

// Take the next object off the work counter and start its bulk copies into the slot (lane 0 issues; the caller has
// made sure that every lane is done with the slot).  Returns the object index or -1.
template <int WC>
__device__ __forceinline__ int fetch_and_stage(const KParams& kp, float* slot, int P, uint64_t* bar, int lane) {
    int obj = 0;
    if (lane == 0) {
        obj = atomicAdd(kp.counters, 1);
        if (obj >= kp.n_obj) obj = -1;
        if (obj >= 0 && kp.use_tma) {
            const float *g3, *g2, *gw;
            object_slabs<WC>(kp, obj, g3, g2, gw);
            fence_proxy_async();  // order our generic-proxy accesses before the async-proxy writes
            if (kp.dense) {
                mbar_expect_tx(bar, (uint32_t)(5 * P * sizeof(float)));
                bulk_g2s(slot, g3, (uint32_t)(3 * P * sizeof(float)), bar);
                bulk_g2s(slot + 5 * P, gw, (uint32_t)(2 * P * sizeof(float)), bar);
            } else {
                mbar_expect_tx(bar, (uint32_t)((5 + WC) * P * sizeof(float)));
                bulk_g2s(slot, g3, (uint32_t)(3 * P * sizeof(float)), bar);
                bulk_g2s(slot + 3 * P, g2, (uint32_t)(2 * P * sizeof(float)), bar);
                bulk_g2s(slot + 5 * P, gw, (uint32_t)(WC * P * sizeof(float)), bar);
            }
        }
    }
    return __shfl_sync(kFull, obj, 0);
}

__device__ __forceinline__ float fast_ex2(float a) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
    return y;
}

// weights -> inverse std in place (uncert_prop_pnp_optimizer.py:73) and the per-axis sums for the inlier thresholds
// (pnp_uncert_cpu.py:164-165)
template <int WMODE, int LAYOUT>
__device__ __forceinline__ void fast_weights(const KParams& kp, float* sw, int P, int lane, float& su, float& sv) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    constexpr int CV = WC - 1;
    // exp(-l) / s = exp2(-l log2(e) - log2(s)): one FFMA + one MUFU.EX2 per weight
    const float k2 = -1.4426950408889634f, off = -__log2f(kp.std_scale);
    su = 0.f; sv = 0.f;
#pragma unroll 4
    for (int p = lane; p < P; p += 32) {
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

// fused head -> PnP prologue, see dense_decode_and_thresholds in pnp_kernel.cuh (same arithmetic); out of line: the
// plain op-level entry never runs it
__device__ __noinline__ void fast_dense_decode(const KParams& kp, int obj, float* slot, int lane, float* sums) {
    float thr_u, thr_v;
    dense_decode_and_thresholds(kp, obj, slot, lane, thr_u, thr_v);
    if (lane == 0) { sums[0] = thr_u; sums[1] = thr_v; }
    __syncwarp();
}

// Inlier decision + packed inlier_out + in-place order-preserving compaction (boolean-mask indexing of
// pnp_uncert_cpu.py:24-27,62-66).  Rows are handled four at a time: four independent load batches, one warp barrier,
// four store batches.  Returns the number of inliers.
template <int WMODE, int LAYOUT>
__device__ __forceinline__ int fast_mask_and_compact(const KParams& kp, int obj, float* slot, int P, int lane, float thr_u,
                                                     float thr_v, bool all_inliers) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    constexpr int CV = WC - 1;
    constexpr int U = 4;
    float* s3 = slot;
    float* s2 = slot + 3 * P;
    float* sw = slot + 5 * P;
    const int rows = (P + 31) >> 5;
    const bool test = kp.istd_thres > 0.f;
    uint32_t in_word = 0u;
    if (kp.inl_in && lane < rows) in_word = __ldg(kp.inl_in + (size_t)obj * rows + lane);
    uint32_t out_word = 0u;
    int base = 0;
#pragma unroll 1
    for (int k0 = 0; k0 < rows; k0 += U) {
        float v3[U][3], v2[U][2], wv_[U][3];
        bool inl[U];
        int dst[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = k0 + u;
            const int p = k * 32 + lane;
            bool ok = p < P;
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
            if (lane == k) out_word = m;
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
            // (no barrier after the stores: they land at or below this batch's rows, never ahead of the read front)
        }
    }
    __syncwarp();
    if (kp.inl_out && lane < rows) kp.inl_out[(size_t)obj * rows + lane] = out_word;
    return base;
}

// Out-of-line wrapper of the on-device linear initialiser (cold for callers that pass init_pose)
template <int WMODE, int LAYOUT>
__device__ __noinline__ bool fast_linear_init(const KParams& kp, int obj, float* slot, int n, int lane, float* scratch,
                                              float* x_out) {
    const int P = kp.n_pts;
    const Camera<float> cam = load_camera<float>(kp, obj);
    double x[4];
    const bool ok = linear_init_impl<WMODE, LAYOUT>(slot, slot + 3 * P, slot + 5 * P, P, TeamRows(n, 0, 0, 0, 1), lane, cam,
                                                    scratch, x);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) x_out[i] = ok ? (float)x[i] : 0.f;  // fp32 hand-over; .py:119-125
    return ok;
}

// Pose covariance (fp64, once per object) and the result row.  H = J^T J at the returned x; no point is clipped there
// (or the object would be on the redo list), so this is both the pipeline covariance (hessian.py:67-87,
// pnp_uncert.py:77-85) and Ceres' own (pnp_uncert_cpu.cpp:279-291).  Lanes 0..3 each solve one column of H^-1.
__device__ __noinline__ void fast_finish_object(const KParams& kp, int obj, int lane, const float* H10, const float* x4,
                                                float cost, float radius, int iteration, int cost_evals, int term) {
    bool usable = term != kFailure;  // Summary::IsSolutionUsable (pnp_uncert_cpu.cpp:276)
    double y[4] = {0.0, 0.0, 0.0, 0.0};
    bool spd = true;
    if (kp.cov_mode != MRPNP_COV_NONE && usable) {
        double H[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) H[i] = (double)H10[i];
        const Ldl4 f = ldl4_factor(H);
        spd = f.ok;
        double e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = (lane == i) ? 1.0 : 0.0;
        ldl4_solve(f, e, y);  // lane c < 4 holds column c
    }
    const bool identity = kp.cov_mode == MRPNP_COV_NONE || !usable || !spd;  // pnp_uncert.py:79-85: H := I, invalid
    if (!spd) usable = false;
    float v = 0.f;
    {
        const int e = lane - 4, r = e >> 2, c = e & 3;
        float cv = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float t = (float)__shfl_sync(kFull, y[i], c);
            cv = (r == i) ? t : cv;
        }
        if (identity) cv = (r == c) ? 1.f : 0.f;
        v = cv;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) v = (lane == i) ? x4[i] : v;
    v = (lane == 20) ? (usable ? 1.f : 0.f) : v;
    v = (lane == 21) ? (float)iteration : v;
    v = (lane == 22) ? cost : v;
    v = (lane == 23) ? radius : v;
    store_result_row(kp, obj, lane, v);
#ifndef MRPNP_TRACE
    if (kp.result64) {
        double d = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) d = (lane == i) ? (double)x4[i] : d;
        d = (lane == 4) ? (double)cost : d;
        d = (lane == 5) ? (double)radius : d;
        d = (lane == 6) ? (double)cost_evals : d;
        d = (lane == 7) ? (double)term : d;
        if (lane < 8) kp.result64[(size_t)obj * 8 + lane] = d;
    }
#endif
}

// fp32 LDL^T of the packed SPD 4x4 (see ldl4_factor) and solve; MUFU reciprocals, no Newton steps.
struct Ldl4f {
    float l10, l20, l30, l21, l31, l32, i0, i1, i2, i3;
    bool ok;
};
__device__ __forceinline__ Ldl4f ldl4f_factor(const float A[10]) {
    Ldl4f f;
    const float d0 = A[0];
    f.i0 = fast_rcp(d0);
    f.l10 = A[1] * f.i0; f.l20 = A[2] * f.i0; f.l30 = A[3] * f.i0;
    const float d1 = fmaf(-f.l10, A[1], A[4]);
    f.i1 = fast_rcp(d1);
    const float t21 = fmaf(-f.l20, A[1], A[5]), t31 = fmaf(-f.l30, A[1], A[6]);
    f.l21 = t21 * f.i1; f.l31 = t31 * f.i1;
    const float d2 = fmaf(-f.l21, t21, fmaf(-f.l20, A[2], A[7]));
    f.i2 = fast_rcp(d2);
    const float t32 = fmaf(-f.l31, t21, fmaf(-f.l30, A[2], A[8]));
    f.l32 = t32 * f.i2;
    const float d3 = fmaf(-f.l32, t32, fmaf(-f.l31, t31, fmaf(-f.l30, A[3], A[9])));
    f.i3 = fast_rcp(d3);
    const float dmin = fminf(fminf(d0, d1), fminf(d2, d3)), dmax = fmaxf(fmaxf(d0, d1), fmaxf(d2, d3));
    f.ok = (dmin > 0.f) && (dmax < 3.0e38f);
    return f;
}
__device__ __forceinline__ void ldl4f_solve(const Ldl4f& f, const float b[4], float y[4]) {
    const float z0 = b[0];
    const float z1 = fmaf(-f.l10, z0, b[1]);
    const float z2 = fmaf(-f.l21, z1, fmaf(-f.l20, z0, b[2]));
    const float z3 = fmaf(-f.l32, z2, fmaf(-f.l31, z1, fmaf(-f.l30, z0, b[3])));
    y[3] = z3 * f.i3;
    y[2] = fmaf(-f.l32, y[3], z2 * f.i2);
    y[1] = fmaf(-f.l31, y[3], fmaf(-f.l21, y[2], z1 * f.i1));
    y[0] = fmaf(-f.l30, y[3], fmaf(-f.l20, y[2], fmaf(-f.l10, y[1], z0 * f.i0)));
}

// sin(d) and cos(d) - 1 with fp32 RELATIVE accuracy: polynomials for |d| <= 0.5 (every yaw step but a wild first one;
// truncation < 1e-8), the library routine -- out of line, it is ~100 instructions -- otherwise and for the initial yaw.
__device__ __noinline__ void sincos_cold(float d, float* sd, float* cdm1) {
    float cd;
    sincosf(d, sd, &cd);
    *cdm1 = cd - 1.f;
}
__device__ __forceinline__ void sincos_cm1(float d, float& sd, float& cdm1) {
    if (fabsf(d) <= 0.5f) {
        const float d2 = d * d;
        float ps = fmaf(d2, 2.7557319e-6f, -1.9841270e-4f);   // 1/9!, -1/7!
        ps = fmaf(ps, d2, 8.3333333e-3f);                      // 1/5!
        ps = fmaf(ps, d2, -1.6666667e-1f);                     // -1/3!
        sd = fmaf(ps * d2, d, d);
        float pc = fmaf(d2, -2.7557319e-7f, 2.4801587e-5f);    // -1/10!, 1/8!
        pc = fmaf(pc, d2, -1.3888889e-3f);                     // -1/6!
        pc = fmaf(pc, d2, 4.1666667e-2f);                      // 1/4!
        pc = fmaf(pc, d2, -0.5f);
        cdm1 = pc * d2;
    } else {
        sincos_cold(d, &sd, &cdm1);
    }
}

constexpr float kFltMax = 3.0e38f;
__device__ __forceinline__ float fast_sqrtf(float a) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
    return y;
}

// ------------------------------------------------------------------ the kernel
// PCT: points per object known at compile time (784 = 28 x 28, every reference config) or 0 = kp.n_pts; with a
// compile-time P the plane offsets of the slot become immediates of the shared-memory loads.
template <int WMODE, int LAYOUT, int PCT>
__global__ void __launch_bounds__(kFastMaxWarps * 32, 1) pnp_lm_fast_kernel(const __grid_constant__ KParams kp) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    unsigned char* header = smem_raw + (size_t)warp * kFastHeaderBytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(header);
    float* scratch = reinterpret_cast<float*>(header + kFastScratch);
    float* slot = reinterpret_cast<float*>(smem_raw + (size_t)nwarps * kFastHeaderBytes) + (size_t)warp * kp.slot_floats;
    const int P = PCT ? PCT : kp.n_pts;
    float* s3 = slot;
    float* s2 = slot + 3 * P;
    float* sw = slot + 5 * P;
    const int max_iter = kp.max_iter > 0 ? kp.max_iter : (kp.max_iter < 0 ? 0 : 50);

    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncwarp();
    uint32_t parity = 0;

    // The bulk copies of an object are started as soon as the slot is free, i.e. BEFORE the covariance / result row of
    // the previous object, so that part of the staging latency is hidden behind it.
    int obj = fetch_and_stage<WC>(kp, slot, P, bar, lane);
#pragma unroll 1
    while (obj >= 0) {
        TR_DECL
        const Camera<float> camf = load_camera<float>(kp, obj);

        // ---------------- stage + weights + inlier mask + compaction ----------------
        int n = P;
#pragma unroll 1
        for (int attempt = 0; attempt < 2; ++attempt) {
            if (kp.use_tma) {
                if (attempt) {  // re-stage the same object (the first attempt compacted the slot)
                    __syncwarp();
                    if (lane == 0) {
                        const float *g3, *g2, *gw;
                        object_slabs<WC>(kp, obj, g3, g2, gw);
                        fence_proxy_async();
                        if (kp.dense) {
                            mbar_expect_tx(bar, (uint32_t)(5 * P * sizeof(float)));
                            bulk_g2s(s3, g3, (uint32_t)(3 * P * sizeof(float)), bar);
                            bulk_g2s(sw, gw, (uint32_t)(2 * P * sizeof(float)), bar);
                        } else {
                            mbar_expect_tx(bar, (uint32_t)((5 + WC) * P * sizeof(float)));
                            bulk_g2s(s3, g3, (uint32_t)(3 * P * sizeof(float)), bar);
                            bulk_g2s(s2, g2, (uint32_t)(2 * P * sizeof(float)), bar);
                            bulk_g2s(sw, gw, (uint32_t)(WC * P * sizeof(float)), bar);
                        }
                    }
                }
                mbar_wait(bar, parity);
                parity ^= 1u;
            } else {
                stage_object_plain<WC>(kp, obj, slot, lane);
            }
            TR_MARK(0)
            float thr_u, thr_v;
            if (WMODE == MRPNP_W_LOGSTD && LAYOUT == MRPNP_LAYOUT_PLANAR && kp.dense) {
                fast_dense_decode(kp, obj, slot, lane, scratch);
                thr_u = scratch[0]; thr_v = scratch[1];
                __syncwarp();
            } else {
                float su, sv;
                fast_weights<WMODE, LAYOUT>(kp, sw, P, lane, su, sv);
                const float invP = 1.f / (float)P;
                thr_u = kp.istd_thres * (su * invP);
                thr_v = kp.istd_thres * (sv * invP);
                __syncwarp();
            }
            // second attempt == pnp_uncert_cpu.py:28-32: <= 4 inliers -> every point is an inlier (slot re-staged)
            const bool all = attempt == 1;
            n = fast_mask_and_compact<WMODE, LAYOUT>(kp, obj, slot, P, lane, thr_u, thr_v, all);
            if (all || n > 4) break;
        }
        TR_MARK(1)

        // ---------------- initial point ----------------
        float x[4], pt[4];
        bool init_ok = true;
        if (kp.init_mode == MRPNP_INIT_GIVEN) {
            const float* ip = kp.init + (size_t)obj * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) pt[i] = __ldg(ip + i);
        } else {
            init_ok = fast_linear_init<WMODE, LAYOUT>(kp, obj, slot, n, lane, scratch, pt);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = pt[i];
        float sn_x, cs_x;
        sincos_cold(pt[0], &sn_x, &cs_x);
        cs_x += 1.f;
        float sn_p = sn_x, cs_p = cs_x;
        TR_MARK(2)

        // ---------------- Levenberg-Marquardt, Ceres 1.14 TrustRegionMinimizer control flow ----------------
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
            // ---- the fused pass at pt: 15 sums, transposed warp reduction, broadcast through the scratch ----
            float a[16];
            bool flagged;
            const bool from_observations = cost_evals < 2;  // initial point (plain fp32), then the fp64 anchor
            if (from_observations) {
                eval_pass_first<WMODE, LAYOUT>(s3, s2, sw, P, RowMap<1>{n}, lane, cost_evals == 1, pt, sn_p, cs_p, camf, a, flagged);
            } else {
                eval_pass_delta<WMODE, LAYOUT>(s3, s2, sw, P, RowMap<1>{n}, lane, dstep, camf, cwin, a, flagged);
            }
            const float tot = warp_reduce16_scatter(a, lane);  // lane L: total of sum (L >> 1)
            // finite iff every total is finite
            const bool jfin = __all_sync(kFull, fabsf(tot) < kFltMax);
            __syncwarp();
            if ((lane & 1) == 0) scratch[lane >> 1] = tot;
            __syncwarp();
            TR_MARK(from_observations ? 3 : 4)
            if (flagged) { redo = true; break; }
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
                // cost - candidate cost
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
                } else {  // HandleUnsuccessfulStep
                    radius = radius * fast_rcp(decrease_factor);
                    decrease_factor *= 2.f;
                    TR_MARK(5)
                    // the slot holds the residuals at the rejected candidate (also after the anchor evaluation)
                    undo_pass_delta<WMODE, LAYOUT>(slot, P, RowMap<1>{n}, lane, dstep, camf.fx, camf.fy);
                    TR_MARK(6)
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
                    // model_cost_change = y^T gs - 1/2 y^T Hs y with (Hs + D) y = gs  =>  1/2 (y^T gs + sum_i D_i y_i^2)
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
            if (stop) { TR_MARK(5) break; }
            // rotation at the candidate by angle addition; the same sin / cos - 1 of the yaw step drive the delta pass
            float sd, cdm1;
            sincos_cm1(delta[0], sd, cdm1);
            sn_p = fmaf(sn_x, cdm1, fmaf(cs_x, sd, sn_x));
            cs_p = fmaf(cs_x, cdm1, fmaf(-sn_x, sd, cs_x));
            dstep.cp = cs_p; dstep.sp = sn_p;
            dstep.txp = pt[1]; dstep.typ = pt[2]; dstep.tzp = pt[3];
            dstep.ncdm1 = -cdm1; dstep.sd = sd;
            dstep.dtx = delta[1]; dstep.dty = delta[2]; dstep.dtz = delta[3];
            TR_MARK(5)
        }
        // the slot is free: start staging the next object before finishing this one
        __syncwarp();
        const int next_obj = fetch_and_stage<WC>(kp, slot, P, bar, lane);
        if (redo) {  // a point near a clip bound: the exact kernel solves this object
            if (lane == 0) kp.redo_list[atomicAdd(kp.redo_count, 1)] = obj;
            obj = next_obj;
            continue;
        }
        // ---------------- pose covariance + result row (out of line, fp64) ----------------
        if (lane == 0) {  // every lane holds the same H and x
#pragma unroll
            for (int i = 0; i < 10; ++i) scratch[i] = H[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) scratch[12 + i] = x[i];
        }
        __syncwarp();
        fast_finish_object(kp, obj, lane, scratch, scratch + 12, cost, radius, iteration, cost_evals, term);
        TR_MARK(7)
#ifdef MRPNP_TRACE
        if (kp.result64 && lane == 0) {
            double* tr = kp.result64 + (size_t)obj * 32;
            for (int i = 0; i < 8; ++i) tr[i] = (double)tr_acc[i];
            tr[8] = (double)cost_evals; tr[9] = (double)(tr_t - tr_t0); tr[10] = (double)n;
            tr[11] = (double)tr_t0; tr[12] = (double)tr_t; tr[13] = (double)blockIdx.x; tr[14] = (double)warp;
        }
#endif
        obj = next_obj;
    }

    // self-resetting work counters: the last CTA to finish rearms them for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        if (kp.n_peers) __threadfence_system();  // rows stored into peer memory are performed before the kernel ends
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
