// pnp_kernel_pool.cuh -- MRPNP_PREC_FAST with pooled shared memory: more objects in flight per SM.
//
// The 22-25 KB slab of an object is only needed WHOLE while it is staged and compacted; the LM loop works on the
// inliers alone (35-50 % of the points on the bench data).  pnp_kernel_fast.cuh nevertheless gives every warp a
// slab-sized slot for the whole life of its object, which limits an SM to 9-10 warps -- 2.5 per scheduler, each
// running a dependent instruction chain (ncu: issue slots 53 % busy, "wait" the top stall).  Here a CTA owns
//   * NS staging buffers of slab size (bulk-copy target, weights, inlier mask, in-place compaction), and
//   * NSM small slots that hold up to kPoolCap compacted points,
// shared by NS + NSM warps.  A warp takes a staging buffer for the prologue of its object, moves the compacted inliers
// into a small slot, hands the staging buffer back, and runs the LM loop out of the small slot; an object with more
// than kPoolCap inliers simply keeps its staging buffer as its slot (no cliff, no fallback launch).  Buffers are
// handed out by atomics on two bit masks in shared memory; a warp never waits for a second resource of a kind it
// already holds, LM warps wait for nothing, so the scheme cannot deadlock.  The arithmetic is that of
// pnp_kernel_fast.cuh (same device functions); the plane stride of a small slot is the compile-time kPoolCap.
#pragma once
#include "pnp_kernel_fast.cuh"

namespace mrpnp {

constexpr int kPoolCap = 448;            // points a small slot holds (14 rows of 32)
constexpr int kPoolMaxWarps = 16;
constexpr int kPoolCtaHeaderBytes = 256; // masks, per-staging-buffer mbarrier + parity
constexpr int kPoolMaxStage = 8;

struct PoolCtaHeader {
    unsigned int stage_mask, small_mask;
    unsigned int parity[kPoolMaxStage];
    uint64_t bar[kPoolMaxStage];
};
static_assert(sizeof(PoolCtaHeader) <= kPoolCtaHeaderBytes, "pool header too small");

// Take one free buffer out of `count` (bit set = in use); lane 0 spins, everyone gets the index.
__device__ __forceinline__ int pool_acquire(unsigned int* mask, int count, int lane) {
    int bit = 0;
    if (lane == 0) {
        const unsigned int all = count >= 32 ? 0xffffffffu : ((1u << count) - 1u);
        while (true) {
            const unsigned int m = *reinterpret_cast<volatile unsigned int*>(mask);
            const unsigned int free_bits = ~m & all;
            if (free_bits) {
                bit = __ffs(free_bits) - 1;
                const unsigned int old = atomicOr(mask, 1u << bit);
                if (!(old & (1u << bit))) break;
            } else {
                __nanosleep(100);
            }
        }
        __threadfence_block();
    }
    return __shfl_sync(kFull, bit, 0);
}
__device__ __forceinline__ void pool_release(unsigned int* mask, int bit, int lane) {
    fence_proxy_async();  // this thread's generic accesses to the buffer precede a later bulk copy into it
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        atomicAnd(mask, ~(1u << bit));
    }
}

struct LmResult {
    float x[4], H[10], cost, radius;
    int iteration, cost_evals, term;
    bool redo;
};

// The LM loop of pnp_kernel_fast.cuh on a slot whose planes are PS floats apart (see there for the commentary).
template <int WMODE, int LAYOUT, int PS_CT>
__device__ __forceinline__ void pool_lm(const KParams& kp, float* slot, int ps_rt, int n, int lane, float* scratch,
                                        const Camera<float>& camf, int max_iter, const float pt0[4], bool init_ok,
                                        LmResult& out) {
    const int PS = PS_CT ? PS_CT : ps_rt;
    float* s3 = slot;
    float* s2 = slot + 3 * PS;
    float* sw = slot + 5 * PS;
    float x[4], pt[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { pt[i] = pt0[i]; x[i] = pt0[i]; }
    float sn_x, cs_x;
    sincos_cold(pt[0], &sn_x, &cs_x);
    cs_x += 1.f;
    float sn_p = sn_x, cs_p = cs_x;
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
            eval_pass_first<WMODE, LAYOUT>(s3, s2, sw, PS, RowMap<1>{n}, lane, cost_evals == 1, pt, sn_p, cs_p, camf, a, flagged);
        } else {
            eval_pass_delta<WMODE, LAYOUT>(s3, s2, sw, PS, RowMap<1>{n}, lane, dstep, camf, cwin, a, flagged);
        }
        const float tot = warp_reduce16_scatter(a, lane);  // lane L: total of sum (L >> 1)
        // finite iff every total is finite
        const bool jfin = __all_sync(kFull, fabsf(tot) < kFltMax);
        __syncwarp();
        if ((lane & 1) == 0) scratch[lane >> 1] = tot;
        __syncwarp();
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
                // the slot holds the residuals at the rejected candidate (also after the anchor evaluation)
                undo_pass_delta<WMODE, LAYOUT>(slot, PS, RowMap<1>{n}, lane, dstep, camf.fx, camf.fy);
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
        if (stop) break;
        // rotation at the candidate by angle addition; the same sin / cos - 1 of the yaw step drive the delta pass
        float sd, cdm1;
        sincos_cm1(delta[0], sd, cdm1);
        sn_p = fmaf(sn_x, cdm1, fmaf(cs_x, sd, sn_x));
        cs_p = fmaf(cs_x, cdm1, fmaf(-sn_x, sd, cs_x));
        dstep.cp = cs_p; dstep.sp = sn_p;
        dstep.txp = pt[1]; dstep.typ = pt[2]; dstep.tzp = pt[3];
        dstep.ncdm1 = -cdm1; dstep.sd = sd;
        dstep.dtx = delta[1]; dstep.dty = delta[2]; dstep.dtz = delta[3];
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) out.x[i] = x[i];
#pragma unroll
    for (int i = 0; i < 10; ++i) out.H[i] = H[i];
    out.cost = cost; out.radius = radius; out.iteration = iteration; out.cost_evals = cost_evals; out.term = term;
    out.redo = redo;
}

// PCT: points per object known at compile time (784) or 0 = kp.n_pts.
template <int WMODE, int LAYOUT, int PCT>
__global__ void __launch_bounds__(kPoolMaxWarps * 32, 1) pnp_lm_pool_kernel(const __grid_constant__ KParams kp) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const int P = PCT ? PCT : kp.n_pts;
    const int NS = kp.pool_stage, NSM = nwarps - NS;
    PoolCtaHeader* ch = reinterpret_cast<PoolCtaHeader*>(smem_raw);
    float* scratch = reinterpret_cast<float*>(smem_raw + kPoolCtaHeaderBytes + (size_t)warp * kFastHeaderBytes + kFastScratch);
    float* stage0 = reinterpret_cast<float*>(smem_raw + kPoolCtaHeaderBytes + (size_t)nwarps * kFastHeaderBytes);
    float* small0 = stage0 + (size_t)NS * kp.slot_floats;
    const int small_floats = (5 + WC) * kPoolCap;
    const int max_iter = kp.max_iter > 0 ? kp.max_iter : (kp.max_iter < 0 ? 0 : 50);

    if (threadIdx.x == 0) {
        ch->stage_mask = 0u; ch->small_mask = 0u;
        for (int i = 0; i < NS; ++i) { ch->parity[i] = 0u; mbar_init(&ch->bar[i], 1); }
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();

#pragma unroll 1
    while (true) {
        int obj = 0;
        if (lane == 0) obj = atomicAdd(kp.counters, 1);
        obj = __shfl_sync(kFull, obj, 0);
        if (obj >= kp.n_obj) break;
        const Camera<float> camf = load_camera<float>(kp, obj);

        // ---------------- staging buffer: bulk copy, weights, inlier mask, compaction ----------------
        const int sb = pool_acquire(&ch->stage_mask, NS, lane);
        float* slot = stage0 + (size_t)sb * kp.slot_floats;
        uint64_t* bar = &ch->bar[sb];
        uint32_t parity = ch->parity[sb];
        int n = P;
#pragma unroll 1
        for (int attempt = 0; attempt < 2; ++attempt) {
            if (kp.use_tma) {
                __syncwarp();
                if (lane == 0) {
                    const float *g3, *g2, *gw;
                    object_slabs<WC>(kp, obj, g3, g2, gw);
                    fence_proxy_async();
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
                mbar_wait(bar, parity);
                parity ^= 1u;
            } else {
                stage_object_plain<WC>(kp, obj, slot, lane);
            }
            float thr_u, thr_v;
            if (WMODE == MRPNP_W_LOGSTD && LAYOUT == MRPNP_LAYOUT_PLANAR && kp.dense) {
                fast_dense_decode(kp, obj, slot, lane, scratch);
                thr_u = scratch[0]; thr_v = scratch[1];
                __syncwarp();
            } else {
                float su, sv;
                fast_weights<WMODE, LAYOUT>(kp, slot + 5 * P, P, lane, su, sv);
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
        if (lane == 0) ch->parity[sb] = parity;

        // ---------------- initial point (the linear initialiser reads the compacted staging buffer) ----------------
        float pt0[4];
        bool init_ok = true;
        if (kp.init_mode == MRPNP_INIT_GIVEN) {
            const float* ip = kp.init + (size_t)obj * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) pt0[i] = __ldg(ip + i);
        } else {
            init_ok = fast_linear_init<WMODE, LAYOUT>(kp, obj, slot, n, lane, scratch, pt0);
        }

        // ---------------- LM: out of a small slot if the inliers fit, else in place ----------------
        LmResult res;
        if (n <= kPoolCap) {
            const int ss = pool_acquire(&ch->small_mask, NSM, lane);
            float* sm = small0 + (size_t)ss * small_floats;
            if (LAYOUT == MRPNP_LAYOUT_PLANAR) {
#pragma unroll
                for (int c = 0; c < 5 + WC; ++c) {
#pragma unroll 4
                    for (int p = lane; p < n; p += 32) sm[c * kPoolCap + p] = slot[c * P + p];
                }
            } else {
                for (int i = lane; i < 3 * n; i += 32) sm[i] = slot[i];
                for (int i = lane; i < 2 * n; i += 32) sm[3 * kPoolCap + i] = slot[3 * P + i];
                for (int i = lane; i < WC * n; i += 32) sm[5 * kPoolCap + i] = slot[5 * P + i];
            }
            pool_release(&ch->stage_mask, sb, lane);
            pool_lm<WMODE, LAYOUT, kPoolCap>(kp, sm, kPoolCap, n, lane, scratch, camf, max_iter, pt0, init_ok, res);
            pool_release(&ch->small_mask, ss, lane);
        } else {
            pool_lm<WMODE, LAYOUT, PCT>(kp, slot, P, n, lane, scratch, camf, max_iter, pt0, init_ok, res);
            pool_release(&ch->stage_mask, sb, lane);
        }
        if (res.redo) {  // a point near a clip bound: the exact kernel solves this object
            if (lane == 0) kp.redo_list[atomicAdd(kp.redo_count, 1)] = obj;
            continue;
        }
        // ---------------- pose covariance + result row (out of line, fp64) ----------------
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 10; ++i) scratch[i] = res.H[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) scratch[12 + i] = res.x[i];
        }
        __syncwarp();
        fast_finish_object(kp, obj, lane, scratch, scratch + 12, res.cost, res.radius, res.iteration, res.cost_evals, res.term);
    }

    // self-resetting work counters: the last CTA to finish rearms them for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        if (kp.n_peers) __threadfence_system();
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
