// pnp_kernel.cuh -- persistent warp-per-object LM kernel.
// MIXED = false: MRPNP_PREC_FP64 (all fp64, reproduces the fp64 reference decisions exactly)
// MIXED = true : MRPNP_PREC_MIXED (fp64 residual/cost chain + fp32 Jacobian sums; fast path)
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

// Stage one object's slab into the warp's slot.  TMA path: three 1-D bulk copies completing on the
// warp's mbarrier; fallback: coalesced loads through registers.
template <int WC>
__device__ __forceinline__ void load_object(const KParams& kp, int obj, float* slot, uint64_t* bar, uint32_t& parity,
                                            int lane) {
    const int P = kp.n_pts;
    const float* g3 = kp.c3d + (size_t)obj * 3 * P;
    const float* g2 = kp.c2d + (size_t)obj * (kp.dense ? 0 : 2 * P);
    const float* gw = kp.wgt + (size_t)obj * WC * P;
    if (kp.dense && kp.pred_stride) {
        // class slice of the head's full output (FCNNOCDecoder.slice_pred, fcn_noc_decoder.py:242-267): channels
        // [3c, 3c+3) of the NOC block and [2c, 2c+2) of the log-std block that follows the 3*C NOC channels
        const long long c = kp.labels ? __ldg(kp.labels + obj) : 0;
        g3 = kp.c3d + (size_t)obj * kp.pred_stride + (size_t)(3 * c) * P;
        gw = kp.wgt + (size_t)obj * kp.pred_stride + (size_t)(2 * c) * P;
    }
    float* s3 = slot;
    float* s2 = slot + 3 * P;
    float* sw = slot + 5 * P;
    __syncwarp();  // every lane is done with the previous object's data
    if (kp.use_tma) {
        if (lane == 0) {
            fence_proxy_async();  // order our generic-proxy accesses before the async-proxy writes
            if (kp.dense) {  // fused head entry: only the NOC map and the raw log-std come from HBM
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
        mbar_wait(bar, parity);
        parity ^= 1u;
    } else if (kp.global_interleaved) {
        // [N,P,C] tensors into a PLANAR slot (the exact routine inside the fast kernel): transpose on the way
        for (int i = lane; i < 3 * P; i += 32) { const int p = i / 3; s3[(i - 3 * p) * P + p] = __ldg(g3 + i); }
        for (int i = lane; i < 2 * P; i += 32) { const int p = i >> 1; s2[(i & 1) * P + p] = __ldg(g2 + i); }
        for (int i = lane; i < WC * P; i += 32) { const int p = i / WC; sw[(i - WC * p) * P + p] = __ldg(gw + i); }
        __syncwarp();
    } else {
        for (int i = lane; i < 3 * P; i += 32) s3[i] = __ldg(g3 + i);
        if (!kp.dense) for (int i = lane; i < 2 * P; i += 32) s2[i] = __ldg(g2 + i);
        for (int i = lane; i < WC * P; i += 32) sw[i] = __ldg(gw + i);
        __syncwarp();
    }
}

// Sweep A: weights -> inverse std in place (uncert_prop_pnp_optimizer.py:73) and the per-axis thresholds
// thres * mean (pnp_uncert_cpu.py:164-165; fp32 like numpy).
template <int WMODE, int LAYOUT, bool FASTEXP>
__device__ __forceinline__ void weights_and_thresholds(const KParams& kp, float* sw, int lane, float& thr_u,
                                                       float& thr_v) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    constexpr int CV = WC - 1;  // channel of the v-axis weight (wyy for full W)
    const int P = kp.n_pts;
    const float inv_scale = 1.f / kp.std_scale;
    float su = 0.f, sv = 0.f;
#pragma unroll 4
    for (int p = lane; p < P; p += 32) {
        float wu = sw[sidx<LAYOUT, WC>(p, 0, P)], wv = sw[sidx<LAYOUT, WC>(p, CV, P)];
        if (WMODE == MRPNP_W_LOGSTD) {
            wu = (FASTEXP ? __expf(-wu) : expf(-wu)) * inv_scale;
            wv = (FASTEXP ? __expf(-wv) : expf(-wv)) * inv_scale;
            sw[sidx<LAYOUT, WC>(p, 0, P)] = wu;
            sw[sidx<LAYOUT, WC>(p, CV, P)] = wv;
        }
        su += wu;
        sv += wv;
    }
    su = warp_sum(su);
    sv = warp_sum(sv);
    const float invP = 1.f / (float)P;
    thr_u = kp.istd_thres * (su * invP);
    thr_v = kp.istd_thres * (sv * invP);
    __syncwarp();
}

// Fused head -> PnP prologue (planar layout, log-std weights): the slot holds the dense head's raw outputs, NOC map in
// the coords_3d planes and proj_logstd in the weight planes.  One sweep decodes them in place and generates coords_2d:
//   coords_3d = (noc*std + mean) * dims                                   NOCCoder.decode, coord_coder/noc_coder.py:50-73
//   var_3d    = dims_var * (noc*std + mean)^2                              (noc_var is None, :68-69)
//   var_2d    = (0.5 (var_x + var_z), var_y)                              distance_invar_proj_error_coder.py:46-50
//   logstd'   = 0.5 log(var_2d (f gain / sd)^2 + exp(2 logstd))           :52-55 with distance = scaling_denominator
//   istd      = exp(-logstd') / std_scale                                 uncert_prop_pnp_optimizer.py:73
//   u[j] = x1 - 0.5 + (j + 0.5)(x2 - x1)/W,  v[i] likewise                roi_align of the pixel grid, monorun_roi_head.py:521-523
// and accumulates the per-axis weight sums for the inlier thresholds (pnp_uncert_cpu.py:164-165).
// Dimensions of one object, decoded when the caller passed the coder's tables (MultiClassNormDimCoder.decode,
// dim_coder/multiclass_norm_dim_coder.py:28-36: dims = dim * std_c + mean_c, dims_var = dim_var * std_c^2; mul and add
// rounded separately, as the two torch launches do).  A label outside the table leaves NaN dimensions: every residual
// of the object is then NaN and the solve reports failure.
__device__ __forceinline__ void object_dims(const KParams& kp, int obj, int lane, float d[3], float v[3]) {
    const float* dm = kp.dims + (size_t)obj * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        d[k] = __ldg(dm + k);
        v[k] = kp.dims_var ? __ldg(kp.dims_var + (size_t)obj * 3 + k) : 0.f;
    }
    if (kp.dim_means) {
        const long long c = kp.dim_labels[obj];
        const bool ok = c >= 0 && c < kp.n_dim_classes;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float sd = ok ? __ldg(kp.dim_stds + c * 3 + k) : __int_as_float(0x7fc00000);
            const float mu = ok ? __ldg(kp.dim_means + c * 3 + k) : 0.f;
            d[k] = __fadd_rn(__fmul_rn(d[k], sd), mu);
            v[k] = __fmul_rn(v[k], __fmul_rn(sd, sd));
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (kp.dims_out) kp.dims_out[(size_t)obj * 3 + k] = d[k];
            if (kp.dims_var_out && kp.dims_var) kp.dims_var_out[(size_t)obj * 3 + k] = v[k];
        }
    }
}

__device__ __forceinline__ void dense_decode_and_thresholds(const KParams& kp, int obj, float* slot, int lane,
                                                            float& thr_u, float& thr_v) {
    const int P = kp.n_pts, W = kp.roi_w, H = P / W;
    float* s3 = slot;
    float* s2 = slot + 3 * P;
    float* sw = slot + 5 * P;
    const float* roi = kp.c2d + (size_t)obj * 4;
    const float x1 = __ldg(roi + 0), y1 = __ldg(roi + 1), x2 = __ldg(roi + 2), y2 = __ldg(roi + 3);
    const float dx = x2 - x1, dy = y2 - y1, x0 = x1 - 0.5f, y0 = y1 - 0.5f;
    const float inv_w = 1.f / (float)W, inv_h = 1.f / (float)H;
    float dm[3], dv[3];
    object_dims(kp, obj, lane, dm, dv);
    const float d0 = dm[0], d1 = dm[1], d2 = dm[2], v0 = dv[0], v1 = dv[1], v2 = dv[2];
    // distance given: the variance is divided by clamp(distance)^2 instead of scaling_denominator^2 (:41-44)
    float inv_scale = 1.f / kp.std_scale;
    if (kp.distance) inv_scale *= fmaxf(__ldg(kp.distance + obj), kp.distance_min) * kp.inv_scaling_denominator;
    float su = 0.f, sv = 0.f;
    // coords_3d and coords_2d use un-contracted roundings so that they are bit-identical to the separate torch
    // launches of the unfused path (mul, add, mul  /  (j+0.5)/W, mul, add)
#pragma unroll 2
    for (int p = lane; p < P; p += 32) {
        const int i = p / W, j = p - i * W;
        const float pn0 = __fadd_rn(__fmul_rn(s3[p], kp.noc_std[0]), kp.noc_mean[0]);
        const float pn1 = __fadd_rn(__fmul_rn(s3[P + p], kp.noc_std[1]), kp.noc_mean[1]);
        const float pn2 = __fadd_rn(__fmul_rn(s3[2 * P + p], kp.noc_std[2]), kp.noc_mean[2]);
        s3[p] = pn0 * d0;
        s3[P + p] = pn1 * d1;
        s3[2 * P + p] = pn2 * d2;
        const float vu = 0.5f * (v0 * pn0 * pn0 + v2 * pn2 * pn2), vv = v1 * pn1 * pn1;
        const float lu = sw[p], lv = sw[P + p];
        const float wu = rsqrtf(fmaf(vu, kp.proj_gain2, __expf(2.f * lu))) * inv_scale;
        const float wv = rsqrtf(fmaf(vv, kp.proj_gain2, __expf(2.f * lv))) * inv_scale;
        sw[p] = wu;
        sw[P + p] = wv;
        s2[p] = __fadd_rn(x0, __fmul_rn(__fmul_rn((float)j + 0.5f, inv_w), dx));
        s2[P + p] = __fadd_rn(y0, __fmul_rn(__fmul_rn((float)i + 0.5f, inv_h), dy));
        su += wu;
        sv += wv;
    }
    su = warp_sum(su);
    sv = warp_sum(sv);
    const float invP = 1.f / (float)P;
    thr_u = kp.istd_thres * (su * invP);
    thr_v = kp.istd_thres * (sv * invP);
    __syncwarp();
}

// Sweep B: inlier decision (pnp_uncert_cpu.py:166-168 or the caller's packed mask), packed inlier_out,
// and in-place order-preserving compaction (boolean-mask indexing of pnp_uncert_cpu.py:24-27,62-66).
// Returns the inlier count; `bits` = this lane's inlier flags (bit k <-> point 32k+lane).
// With `compact` the slot is rewritten; the caller must re-load it if the count turns out to be <= 4.
template <int WMODE, int LAYOUT>
__device__ __forceinline__ int mask_and_compact(const KParams& kp, int obj, float* slot, int lane, float thr_u,
                                                float thr_v, bool all_inliers, bool compact, uint32_t& bits) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    constexpr int CV = WC - 1;
    const int P = kp.n_pts;
    float* s3 = slot;
    float* s2 = slot + 3 * P;
    float* sw = slot + 5 * P;
    const int rows = (P + 31) >> 5;
    const bool test = kp.istd_thres > 0.f;
    uint32_t in_word = 0u;
    if (kp.inl_in && lane < rows) in_word = __ldg(kp.inl_in + (size_t)obj * rows + lane);
    uint32_t out_word = 0u;
    bits = 0u;
    int base = 0;
    for (int k = 0; k < rows; ++k) {
        const int p = k * 32 + lane;
        bool inl = p < P;
        float wu = 0.f, wv = 0.f;
        if (inl) {
            wu = sw[sidx<LAYOUT, WC>(p, 0, P)];
            wv = sw[sidx<LAYOUT, WC>(p, CV, P)];
        }
        if (!all_inliers) {
            if (kp.inl_in) {
                const uint32_t row_word = __shfl_sync(kFull, in_word, k);  // every lane takes part
                inl = inl && ((row_word >> lane) & 1u);
            } else if (test) {
                inl = inl && (wu >= thr_u) && (wv >= thr_v);
            }
        }
        const unsigned m = __ballot_sync(kFull, inl);
        bits |= (inl ? 1u : 0u) << k;
        if (lane == k) out_word = m;
        if (compact) {
            float v3[3], v2[2], w1 = 0.f;
            if (inl) {
#pragma unroll
                for (int c = 0; c < 3; ++c) v3[c] = s3[sidx<LAYOUT, 3>(p, c, P)];
#pragma unroll
                for (int c = 0; c < 2; ++c) v2[c] = s2[sidx<LAYOUT, 2>(p, c, P)];
                if (WMODE == MRPNP_W_FULL) w1 = sw[sidx<LAYOUT, WC>(p, 1, P)];
            }
            __syncwarp();  // this row's reads (all lanes) happen before any lane's compacted writes
            if (inl) {
                const int d = base + __popc(m & ((1u << lane) - 1u));
#pragma unroll
                for (int c = 0; c < 3; ++c) s3[sidx<LAYOUT, 3>(d, c, P)] = v3[c];
#pragma unroll
                for (int c = 0; c < 2; ++c) s2[sidx<LAYOUT, 2>(d, c, P)] = v2[c];
                sw[sidx<LAYOUT, WC>(d, 0, P)] = wu;
                if (WMODE == MRPNP_W_FULL) sw[sidx<LAYOUT, WC>(d, 1, P)] = w1;
                sw[sidx<LAYOUT, WC>(d, CV, P)] = wv;
            }
        }
        base += __popc(m);
    }
    __syncwarp();
    if (kp.inl_out && lane < rows) kp.inl_out[(size_t)obj * rows + lane] = out_word;
    return base;
}

// One object, start to finish, by one warp: stage, weights, inlier mask, compaction, Ceres-1.14 LM, covariance, result
// row.  Out of line: it is the whole body of pnp_lm_kernel below AND the routine the MRPNP_PREC_FAST kernel calls
// (MIXED = false, exact fp64 decisions) for the few objects it hands back (a point near a clip bound, an accept /
// function-tolerance decision within rounding of its threshold).  `scratch`: 40 doubles of the warp's header.
// TEAM: warp 0 of a CTA whose other warps sit in team_worker(): the fp64 evaluations go through team_eval_fp64.
template <bool MIXED, int WMODE, int LAYOUT, bool TEAM = false>
__device__ __noinline__ uint32_t solve_object_exact(const KParams& kp, int obj, float* slot, uint64_t* bar, uint32_t parity,
                                                    double* scratch, int lane, RedoTeam* team = nullptr,
                                                    uint32_t* team_phase = nullptr) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    const int P = kp.n_pts;
    const float* s3 = slot;
    const float* s2 = slot + 3 * P;
    const float* sw = slot + 5 * P;
    const int max_iter = kp.max_iter > 0 ? kp.max_iter : (kp.max_iter < 0 ? 0 : 50);
    const Camera<float> camf = load_camera<float>(kp, obj);
    Camera<double> cam;
    cam.fx = camf.fx; cam.fy = camf.fy; cam.cx = camf.cx; cam.cy = camf.cy;  // only these four are used inline
    uint32_t tphase = TEAM ? *team_phase : 0u;   // round count of the team barrier, in a register while this routine runs

    // ---------------- stage + istd + inlier mask + compaction ----------------
    uint32_t bits = 0u;
    int n_inliers = P, n = P;
    bool compacted = false;
    for (int attempt = 0; attempt < 2; ++attempt) {
        load_object<WC>(kp, obj, slot, bar, parity, lane);
        float thr_u, thr_v;
        if (WMODE == MRPNP_W_LOGSTD && LAYOUT == MRPNP_LAYOUT_PLANAR && kp.dense)
            dense_decode_and_thresholds(kp, obj, slot, lane, thr_u, thr_v);
        else
            weights_and_thresholds<WMODE, LAYOUT, MIXED>(kp, slot + 5 * P, lane, thr_u, thr_v);
        // second attempt == pnp_uncert_cpu.py:28-32: <= 4 inliers -> every point is an inlier (slot re-staged)
        const bool all = attempt == 1;
        const bool compact = !all && kp.inlier_opt_only != 0;
        n_inliers = mask_and_compact<WMODE, LAYOUT>(kp, obj, slot, lane, thr_u, thr_v, all, compact, bits);
        if (all || n_inliers > 4) {
            compacted = compact;
            break;
        }
    }
    n = compacted ? n_inliers : P;

    // ---------------- Levenberg-Marquardt, Ceres 1.14 TrustRegionMinimizer control flow ----------------
    // One evaluation site: `pt` is the initial point in phase 0 and the candidate x + delta afterwards.
    double x[4], pt[4];
    bool init_ok = true;
    if (kp.init_mode == MRPNP_INIT_GIVEN) {
        const float* ip = kp.init + (size_t)obj * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) pt[i] = (double)__ldg(ip + i);
    } else {
        init_ok = linear_init<WMODE, LAYOUT>(kp, obj, slot, n, lane, scratch);
#pragma unroll
        for (int i = 0; i < 4; ++i) pt[i] = scratch[kScrPt + i];
        __syncwarp();
    }
    if ((kp.ransac_thres || kp.ransac_ratio > 0.f) && compacted && init_ok) {
        // reprojection-threshold consensus with the start pose as the model (consensus_prune, pnp_device.cuh)
        float thr;
        if (kp.dense) {
            const float* roi = kp.c2d + (size_t)obj * 4;
            const int Hh = P / kp.roi_w;
            thr = kp.ransac_ratio * (__ldg(roi + 3) - __ldg(roi + 1)) * (float)(Hh - 1) / (float)Hh;
        } else {
            thr = __ldg(kp.ransac_thres + obj);
        }
        if (thr > 0.f) {
            // model = the linear initialiser's pose (init_pose only says where LM starts)
            bool model_ok = init_ok;
            if (kp.init_mode == MRPNP_INIT_GIVEN) model_ok = linear_init<WMODE, LAYOUT>(kp, obj, slot, n, lane, scratch);
            for (int round = 0; round < 3 && model_ok; ++round)   // graduated trimmed fits, as in the fast kernel
                if (!linear_init<WMODE, LAYOUT>(kp, obj, slot, n, lane, scratch, thr * (round == 0 ? 3.f : round == 1 ? 2.f : 1.4f))) break;
            if (model_ok) {
                float* fs = reinterpret_cast<float*>(scratch + 24);
                __syncwarp();
                if (lane < 4) fs[lane] = (float)scratch[kScrPt + lane];
                __syncwarp();
                const int n2 = consensus_prune<WMODE, LAYOUT>(kp, obj, slot, n, lane, fs, thr);
                __syncwarp();
                if (n2 != n) {
                    n = n2;
                    n_inliers = n2;
                    model_ok = linear_init<WMODE, LAYOUT>(kp, obj, slot, n, lane, scratch);
                }
            }
            if (kp.init_mode != MRPNP_INIT_GIVEN) {
                init_ok = model_ok;
#pragma unroll
                for (int i = 0; i < 4; ++i) pt[i] = scratch[kScrPt + i];
                __syncwarp();
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = pt[i];

    double cost = 0.0, g[4], H[10], scale[4], diag[4], delta[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { g[i] = 0.0; scale[i] = 1.0; diag[i] = 1.0; delta[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < 10; ++i) H[i] = 0.0;
    int term = kNoConvergence, iteration = 0, cost_evals = 0, num_invalid = 0;
    double radius = kInitialRadius, decrease_factor = 2.0, x_norm = 0.0, model_change = 1.0;
    bool reuse_diagonal = false, step_ok = true, clip_x = false, first = true;
    double sn_x = 0.0, cs_x = 1.0, sn_p = 0.0, cs_p = 1.0;  // sin/cos of the accepted point / of pt

    while (true) {
        // ---- the fused pass at pt ----
        double acc[16];
        bool clip_p = false, jfinite = true;
        bool exact = !MIXED;
        if (MIXED) {
            // rotation at pt: library sincos once per object, then the angle-addition update from the accepted
            // point with a small-angle polynomial (the step in yaw is tiny after the first iteration)
            double sn, cs;
            const double dyaw = pt[0] - x[0];
            if (first || fabs(dyaw) > 0.5) {
                sincos(pt[0], &sn, &cs);
            } else {
                double sd, cd;
                sincos_small(dyaw, &sd, &cd);
                sn = fma(sn_x, cd, cs_x * sd);
                cs = fma(cs_x, cd, -sn_x * sd);
            }
            sn_p = sn; cs_p = cs;
            eval_pass_mixed<WMODE, LAYOUT>(s3, s2, sw, P, n, lane, pt, sn, cs, cam, camf, acc, scratch, exact, jfinite);
        }
        if (exact) {  // MRPNP_PREC_FP64, or a point near a clip bound in the mixed pass: exact fp64 clip semantics
            __syncwarp();
            if (lane < 4) {
                double v = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) v = (lane == i) ? pt[i] : v;
                scratch[kScrPt + lane] = v;
            }
            __syncwarp();
            if (TEAM) team_eval_fp64<WMODE, LAYOUT>(kp, team, tphase, obj, slot, n, lane, 0, scratch);
            else eval_pass_fp64<WMODE, LAYOUT>(kp, obj, slot, n, lane, bits, 0, false, scratch);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = scratch[kScrSums + i];
            clip_p = scratch[kScrClip] != 0.0;
            __syncwarp();
            const double t0 = (fabs(acc[1]) + fabs(acc[2])) + (fabs(acc[3]) + fabs(acc[4]));
            const double t1 = (fabs(acc[5]) + fabs(acc[6])) + (fabs(acc[7]) + fabs(acc[8]));
            const double t2 = (fabs(acc[9]) + fabs(acc[10])) + (fabs(acc[11]) + fabs(acc[12]));
            jfinite = finite_value((t0 + t1) + (t2 + (fabs(acc[13]) + fabs(acc[14]))));
        }
        ++cost_evals;
        const bool cfinite = finite_value(acc[0]);
        jfinite = jfinite && cfinite;
        bool accept = false;
        if (first) {  // IterationZero
            first = false;
            if (!jfinite || !init_ok) { term = kFailure; break; }  // parameters stay at init
            accept = true;
            cost = 0.5 * acc[0];
#pragma unroll
            for (int i = 0; i < 4; ++i)  // jacobi_scaling from the initial Jacobian only
                scale[i] = fast_rcp(1.0 + fast_sqrt(acc[5 + tri(i, i)]));
        } else {
            const double cand_cost = cfinite ? 0.5 * acc[0] : kDblMax;
            // ParameterToleranceReached
            const double step_norm2 = delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2] + delta[3] * delta[3];
            const double ptol = kParameterTol * (x_norm + kParameterTol);
            if (step_norm2 <= ptol * ptol) { term = kConvergence; break; }
            // FunctionToleranceReached (Ceres 1.14: the candidate is not adopted on this exit)
            const double cost_change = cost - cand_cost;
            if (fabs(cost_change) <= kFunctionTol * cost) {
                term = kConvergence;
                if (!(kp.adopt_ftol && cand_cost < cost)) break;
                accept = true;  // documented switch: take the candidate, then stop
            }
            const double rho = cost_change * fast_rcp1(model_change);
            if (accept || rho > kMinRelDecrease) {  // HandleSuccessfulStep
                if (!jfinite) { term = kFailure; break; }  // re-evaluation at the new point fails in Ceres
                const bool stop = accept;
                accept = true;
                cost = cand_cost;
                const double q = 2.0 * rho - 1.0;
                radius = fmin(kMaxRadius, radius * fast_rcp(fmax(1.0 / 3.0, 1.0 - q * q * q)));
                decrease_factor = 2.0;
                reuse_diagonal = false;
                if (stop) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { x[i] = pt[i]; g[i] = acc[1 + i]; }
#pragma unroll
                    for (int i = 0; i < 10; ++i) H[i] = acc[5 + i];
                    clip_x = clip_p;
                    sn_x = sn_p; cs_x = cs_p;
                    break;
                }
            } else {  // HandleUnsuccessfulStep
                radius /= decrease_factor;
                decrease_factor *= 2.0;
            }
        }
        if (accept) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { x[i] = pt[i]; g[i] = acc[1 + i]; }
#pragma unroll
            for (int i = 0; i < 10; ++i) H[i] = acc[5 + i];
            clip_x = clip_p;
            sn_x = sn_p; cs_x = cs_p;
            x_norm = fast_sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
            step_ok = true;
        }
        // ---- next trust-region step (invalid steps shrink the radius without a new evaluation) ----
        bool stop = false;
        while (true) {
            // FinalizeIterationAndCheckIfMinimizerCanContinue
            if (iteration >= max_iter) { term = kNoConvergence; stop = true; break; }
            if (step_ok) {
                const double gmax = fmax(fmax(fabs(g[0]), fabs(g[1])), fmax(fabs(g[2]), fabs(g[3])));
                if (gmax <= kGradientTol) { term = kConvergence; stop = true; break; }
            }
            if (radius <= kMinRadius) { term = kConvergence; stop = true; break; }
            ++iteration;
            step_ok = false;
            // LevenbergMarquardtStrategy::ComputeStep on the column-scaled system
            double Hs[10], gs[4], A[10], y[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                gs[i] = g[i] * scale[i];
#pragma unroll
                for (int j = i; j < 4; ++j) Hs[tri(i, j)] = H[tri(i, j)] * (scale[i] * scale[j]);
            }
            if (!reuse_diagonal) {
#pragma unroll
                for (int i = 0; i < 4; ++i) diag[i] = fmin(fmax(Hs[tri(i, i)], kMinLmDiag), kMaxLmDiag);
            }
            reuse_diagonal = true;
            const double inv_radius = fast_rcp1(radius);
#pragma unroll
            for (int i = 0; i < 10; ++i) A[i] = Hs[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) A[tri(i, i)] = fma(diag[i], inv_radius, A[tri(i, i)]);
            const Ldl4 f = ldl4_factor(A);
            bool valid = f.ok;
            if (valid) {
                ldl4_solve(f, gs, y);  // step = -y
                // model_cost_change = y^T gs - 1/2 y^T Hs y with (Hs + D) y = gs  =>  1/2 (y^T gs + sum_i D_i y_i^2)
                double yg = 0.0, ydy = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    yg = fma(y[i], gs[i], yg);
                    ydy = fma(diag[i] * inv_radius * y[i], y[i], ydy);
                }
                model_change = 0.5 * (yg + ydy);
                valid = (model_change > 0.0) && finite_value(fabs(y[0]) + fabs(y[1]) + fabs(y[2]) + fabs(y[3]));
            }
            if (valid) {
                num_invalid = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) { delta[i] = -y[i] * scale[i]; pt[i] = x[i] + delta[i]; }
                break;
            }
            // HandleInvalidStep
            if (++num_invalid >= kMaxInvalidSteps) { term = kFailure; stop = true; break; }
            radius /= decrease_factor;
            decrease_factor *= 2.0;
        }
        if (stop) break;
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
            __syncwarp();
            if (lane < 4) {
                double v = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) v = (lane == i) ? x[i] : v;
                scratch[kScrPt + lane] = v;
            }
            __syncwarp();
            if (TEAM && compacted) team_eval_fp64<WMODE, LAYOUT>(kp, team, tphase, obj, slot, n, lane, 1, scratch);
            else eval_pass_fp64<WMODE, LAYOUT>(kp, obj, slot, n, lane, bits, 1, !compacted, scratch);
#pragma unroll
            for (int i = 0; i < 10; ++i) H[i] = scratch[kScrSums + 5 + i];
            __syncwarp();
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
        store_result_row(kp, obj, lane, v);
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
    if (TEAM) {
        __syncwarp();
        if (lane == 0) *team_phase = tphase;
        __syncwarp();
    }
    return parity;   // mbarrier phase after this object's copies (by value: no local of the caller has its address taken)
}

template <bool MIXED, int WMODE, int LAYOUT>
__global__ void __launch_bounds__(kMaxThreads, 1) pnp_lm_kernel(const __grid_constant__ KParams kp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    unsigned char* header = smem_raw + (size_t)warp * kWarpHeaderBytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(header);
    double* scratch = reinterpret_cast<double*>(header + 128);
    float* slot = reinterpret_cast<float*>(smem_raw + (size_t)nwarps * kWarpHeaderBytes) + (size_t)warp * kp.slot_floats;

    // work list: all objects, or (follow-up launch of a FAST solve) the objects the fast kernel handed back
    int n_work = kp.n_obj;
    if (kp.work_list) {
        n_work = *reinterpret_cast<const volatile int*>(kp.work_count);
        if (n_work == 0) return;  // the usual case of the follow-up launch: nothing was handed back, no counter was touched
    }

    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncwarp();
    wait_for_acks(kp);
    uint32_t parity = 0;

    while (true) {
        int obj = 0;
        if (lane == 0) {
            obj = atomicAdd(kp.counters, 1);
            if (obj >= n_work) obj = -1;
            else if (kp.work_list) obj = kp.work_list[obj];
        }
        obj = __shfl_sync(kFull, obj, 0);
        if (obj < 0) break;

        parity = solve_object_exact<MIXED, WMODE, LAYOUT>(kp, obj, slot, bar, parity, scratch, lane);
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
            if (kp.work_list) *kp.work_count = 0;  // the follow-up launch consumed the redo list
            __threadfence();
            raise_peer_flags(kp);
        }
    }
}

}  // namespace mrpnp
