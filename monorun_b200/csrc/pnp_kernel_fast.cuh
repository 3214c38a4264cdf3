// pnp_kernel_fast.cuh -- MRPNP_PREC_FAST kernel: warp per object, tracked residuals, packed-fp32 passes, fp32 scalar LM,
// a hot path small enough for the SM's instruction cache, and the exact fp64 routine for the few objects that need it.
//
// What bounded the round-1 kernels on B200 was neither HBM nor a math pipe but the issue rate of a latency-bound
// dependent chain at nine warps per SM (226 KB of shared memory hold nine 25 KB objects), on top of an instruction
// cache that the first kernels overflowed.  So this kernel is written for (a) few instructions per object and (b) code
// size:
//   * residuals are evaluated once in fp64 and then tracked incrementally (pnp_fast.cuh);
//   * every pass handles TWO points per lane with packed fp32 instructions (FFMA2 / FADD2 / FMUL2) on 64-bit
//     shared-memory accesses, in the normalised formulation of pnp_fast.cuh (a third fewer multiply-adds), over arrays
//     padded to whole 64-point groups (no validity predicates), and tests points against the clip bounds only when a
//     bounding box of the object is not inside the clip window;
//   * the trust-region algebra (Ceres 1.14 control flow, unchanged) is fp32: the normal equations come from fp32 sums
//     anyway, Jacobi scaling keeps the 4x4 system well inside fp32 range, and the accept / function-tolerance decisions
//     read the cost CHANGE straight from the delta pass (relative error ~1e-6 of itself);
//   * everything cold (linear initialiser, fused head prologue, unaligned / interleaved staging, roll-back of a rejected
//     step, the remainder of an unpadded pass, the fp64 covariance) is out of line.
// Objects the fp32 path must not decide -- a point near a clip bound, or an accept / function-tolerance decision within
// the rounding band of its threshold -- are handed, inside the same launch, to solve_object_exact<fp64> (pnp_kernel.cuh),
// which reproduces the fp64 reference decision for decision: a lock-free list in global memory that every warp polls
// before it takes a fresh object.
#pragma once
#include "pnp_kernel.cuh"
#include "pnp_fast.cuh"

// Phase trace (tools/trace_run.py, build with -DMRPNP_TRACE): clock64 ticks per phase, summed per object, written to
// the result64 buffer viewed as [N,32] doubles.  Phases: 0 staging wait, 1 weights + mask + compaction, 2 initialiser,
// 3 first two evaluations, 4 candidate evaluations, 5 scalar trust-region algebra, 6 roll-backs, 7 covariance + stores.
#ifdef MRPNP_TRACE
#define TR_DECL long long tr_t = clock64(), tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long tr_t0 = tr_t;
#define TR_MARK(k) { const long long tr_now = clock64(); tr_acc[k] += tr_now - tr_t; tr_t = tr_now; }
#else
#define TR_DECL
#define TR_MARK(k)
#endif

namespace mrpnp {

// Resident warps (= objects in flight) per SM.  Shared memory would hold 9 (full 2x2 weights) or 10 (diagonal) slots,
// but registers are handed out per four warps: above 8 warps a thread gets 168 registers, and at 168 the kernel
// spills its warp-uniform LM state to LOCAL memory -- ~106 KB per CTA against ~28 KB of L1 at this shared-memory
// carve-out, i.e. an L2 round trip per reload.  Measured (profiles/r02_ab_variants.txt): 8 warps x 255 registers, no
// spills, 162 us; 9 warps x 168 registers 196 us.
#ifndef MRPNP_FAST_WARPS
#define MRPNP_FAST_WARPS 8
#endif
constexpr int kFastMaxWarps = MRPNP_FAST_WARPS;
// Per-warp header, 512 B = 128 floats: [0] mbarrier | [4..27] sums buffer A | [32..55] sums buffer B | [56..63] Jacobi
// scale, LM diagonal | [64..95] argument stash of the out-of-line routines | [96..111] camera + clip window (pass constants).  The exact routine's 40-double scratch
// starts at float 32: it only runs between objects, when the fast path's state is dead.
// A sums buffer: [0..3] J^T r, [4..13] J^T J, [14] cost term, [16..18] extents; [20..23] linear-initialiser result.
// [128..671] the 15 x 36 tile of the passes' warp reduction (warp_reduce15_smem).
constexpr int kFastReduce = 128;
// MRPNP_EXP_FAST_TEAM only: [672..751] four partial-total rows of a split delta pass (team_part, pnp_fast.cuh).
constexpr int kFastParts = kFastReduce + ((kReduceTileFloats * 4 + 127) / 128) * 32;
#ifdef MRPNP_EXP_FAST_TEAM
constexpr int kFastHeaderBytes = 512 + ((kReduceTileFloats * 4 + 127) / 128) * 128 + 384;
#else
constexpr int kFastHeaderBytes = 512 + ((kReduceTileFloats * 4 + 127) / 128) * 128;
#endif
constexpr int kFastBufA = 4, kFastBufB = 32, kFastScaleDiag = 56, kFastArgStash = 64, kFastConsts = 96;
constexpr int kFastScratch64 = 128;   // byte offset
constexpr int kNoPending = -1;
// work counters of one launch (ints, all zero between launches; the last CTA re-arms them)
enum { kCntFresh = 0, kCntCtasDone = 1, kCntRedoCount = 2, kCntRedoTaken = 3, kCntCtasInRedo = 4 };
constexpr int kRedoKeepers = 24;   // CTAs that stay until every CTA has entered the redo phase

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

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

// Unaligned shapes / pointers and op-level [N,P,C] tensors: coalesced loads through registers instead of bulk copies;
// interleaved tensors are transposed on the way so that the slot is planar in every case.
template <int WC>
__device__ __noinline__ void stage_object_plain(const KParams& kp, int obj, float* slot, int lane) {
    const int P = kp.n_pts;
    const float *g3, *g2, *gw;
    object_slabs<WC>(kp, obj, g3, g2, gw);
    if (kp.global_interleaved) {
        for (int i = lane; i < 3 * P; i += 32) { const int p = i / 3; slot[(i - 3 * p) * P + p] = __ldg(g3 + i); }
        for (int i = lane; i < 2 * P; i += 32) { const int p = i >> 1; slot[(3 + (i & 1)) * P + p] = __ldg(g2 + i); }
        for (int i = lane; i < WC * P; i += 32) { const int p = i / WC; slot[(5 + (i - WC * p)) * P + p] = __ldg(gw + i); }
    } else {
        for (int i = lane; i < 3 * P; i += 32) slot[i] = __ldg(g3 + i);
        if (!kp.dense) for (int i = lane; i < 2 * P; i += 32) slot[3 * P + i] = __ldg(g2 + i);
        for (int i = lane; i < WC * P; i += 32) slot[5 * P + i] = __ldg(gw + i);
    }
    __syncwarp();
}

// Lane 0: start the bulk copies of a (fresh) object into the slot; the caller has made sure every lane is done with it.
template <int WC>
__device__ __forceinline__ void issue_bulk_copies(const KParams& kp, int obj, float* slot, int P, uint64_t* bar) {
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

// Work distribution.  Every warp draws fresh objects from one counter (kCntFresh).  An object the fp32 path must not
// decide is appended to a list in global memory (hand_back: kCntRedoCount over kp.redo_list) and the warp goes on with
// the next fresh object; the list is worked off in the redo phase at the end of the kernel, by whole CTAs (eight warps to
// one fp64 evaluation), on SMs that would otherwise idle through the launch's tail.  (Measured and dropped: the warp that
// hands an object back solves it itself right away -- MRPNP_EXP_INLINE_REDO keeps that variant: a second, fp64 solve by
// one warp that starts when the first ends put +50 us on the launch, profiles/r02_band_sweep.txt; drawing the ticket one
// object ahead and prefetching the next object's slabs into L2 -- 5 to 10 % slower, profiles/r02_ab_variants.txt.)
// Returns the object or -1 (nothing left); for a fresh object on the TMA path the bulk copies are started.
template <int WC>
__device__ __forceinline__ int fetch_job(const KParams& kp, float* slot, int P, uint64_t* bar, int lane, int& pending,
                                         bool& is_redo) {
    int obj = -1, redo = 0;
#ifndef MRPNP_EXP_INLINE_REDO
    // handed-back objects are solved by whole CTAs after the fresh objects (redo phase at the end of the kernel)
    if (lane == 0) {
        const int fresh = atomicAdd(kp.counters + kCntFresh, 1);
        if (fresh < kp.n_obj) {
            obj = fresh;
            if (kp.use_tma) issue_bulk_copies<WC>(kp, obj, slot, P, bar);
        }
    }
    obj = __shfl_sync(kFull, obj, 0);
    is_redo = false;
    return obj;
#endif
    if (lane == 0) {
        int* c = kp.counters;
        int fresh = pending;
        if (fresh == kNoPending) fresh = atomicAdd(c + kCntFresh, 1);   // in flight together with the two loads below
        // relaxed: nothing but the entry itself is read through these counters, and the entry is polled until written
        const int rc = ld_relaxed(c + kCntRedoCount), rt = ld_relaxed(c + kCntRedoTaken);
        pending = fresh;
        if (rt < rc && atomicCAS(c + kCntRedoTaken, rt, rt + 1) == rt) {
            // entries are object + 1; 0 = reserved but not yet written.  The reader re-arms the entry.
            volatile int* e = kp.redo_list + rt;
            int v;
            while ((v = *e) == 0) {}
            *e = 0;
            obj = v - 1;
            redo = 1;
        } else if (fresh < kp.n_obj) {
            obj = fresh;
            pending = kNoPending;
            if (kp.use_tma) issue_bulk_copies<WC>(kp, obj, slot, P, bar);
        } else if (rt < rc) {
            obj = -2;
        }
    }
    obj = __shfl_sync(kFull, obj, 0);
    is_redo = __shfl_sync(kFull, redo, 0) != 0;
    return obj;
}

// A fresh object the fp32 path must not decide: append it to the hand-back list.
__device__ __forceinline__ void hand_back(const KParams& kp, int obj, int lane) {
    if (lane == 0) {
        const int i = atomicAdd(kp.counters + kCntRedoCount, 1);
        *reinterpret_cast<volatile int*>(kp.redo_list + i) = obj + 1;
        __threadfence();
    }
}

__device__ __forceinline__ float fast_ex2(float a) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
    return y;
}

// weights -> inverse std in place (uncert_prop_pnp_optimizer.py:73) and the per-axis sums for the inlier thresholds
// (pnp_uncert_cpu.py:164-165); two points per lane and access
template <int WMODE>
__device__ __forceinline__ void fast_weights(const KParams& kp, float* sw, int P, int lane, float& su, float& sv) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    // exp(-l) / s = exp2(-l log2(e) - log2(s)): one FFMA + one MUFU.EX2 per weight
    const float k2 = -1.4426950408889634f, off = -__log2f(kp.std_scale);
    float2 au = make_float2(0.f, 0.f), av = au;
    float* pu = sw;
    float* pv = sw + (WC - 1) * P;
#pragma unroll 4
    for (int i = 2 * lane; i < P; i += 64) {
        float2 wu = *reinterpret_cast<const float2*>(pu + i), wv = *reinterpret_cast<const float2*>(pv + i);
        if (WMODE == MRPNP_W_LOGSTD) {
            wu = __ffma2_rn(wu, make_float2(k2, k2), make_float2(off, off));
            wv = __ffma2_rn(wv, make_float2(k2, k2), make_float2(off, off));
            wu = make_float2(fast_ex2(wu.x), fast_ex2(wu.y));
            wv = make_float2(fast_ex2(wv.x), fast_ex2(wv.y));
            *reinterpret_cast<float2*>(pu + i) = wu;
            *reinterpret_cast<float2*>(pv + i) = wv;
        }
        au = __fadd2_rn(au, wu);
        av = __fadd2_rn(av, wv);
    }
    su = warp_sum(au.x + au.y);
    sv = warp_sum(av.x + av.y);
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
// pnp_uncert_cpu.py:24-27,62-66) of the planar slot.  Returns the number of inliers.
//
// General form (out of line): caller-supplied masks, "every point" (second attempt, test disabled) and the rows the
// unrolled istd-test loop below leaves over.  One row of 32 points at a time, rows [k_begin, rows).
template <int WMODE>
__device__ __noinline__ int fast_compact_rows(const KParams& kp, int obj, float* slot, int lane, float thr_u, float thr_v,
                                              bool all_inliers, int k_begin, int base, uint32_t out_word) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    constexpr int NPL = 5 + WC;
    const int P = kp.n_pts;
    const int rows = (P + 31) >> 5;
    const bool test = kp.istd_thres > 0.f;
    uint32_t in_word = 0u;
    if (kp.inl_in && lane < rows) in_word = __ldg(kp.inl_in + (size_t)obj * rows + lane);
#pragma unroll 1
    for (int k = k_begin; k < rows; ++k) {
        const int p = k * 32 + lane;
        bool ok = p < P;
        const int pc = ok ? p : 0;
        float val[NPL];
#pragma unroll
        for (int c = 0; c < NPL; ++c) val[c] = slot[c * P + pc];
        if (!all_inliers) {
            if (kp.inl_in) {
                const uint32_t row_word = __shfl_sync(kFull, in_word, k & 31);  // every lane takes part
                ok = ok && ((row_word >> lane) & 1u);
            } else if (test) {
                ok = ok && (val[5] >= thr_u) && (val[NPL - 1] >= thr_v);
            }
        }
        const unsigned m = __ballot_sync(kFull, ok);
        if (lane == k) out_word = m;
        const int dst = base + __popc(m & ((1u << lane) - 1u));
        base += __popc(m);
        if (!all_inliers) {
            __syncwarp();  // the reads of this row (all lanes) happen before any lane's compacted writes
            if (ok) {
#pragma unroll
                for (int c = 0; c < NPL; ++c) slot[c * P + dst] = val[c];
            }
        }
    }
    __syncwarp();
    if (kp.inl_out && lane < rows) kp.inl_out[(size_t)obj * rows + lane] = out_word;
    return base;
}

// The hot form: the istd test (pnp_uncert_cpu.py:164-168) on rows that lie entirely inside the planes, four rows of 32
// points at a time -- four independent load batches, one warp barrier, four store batches, no bounds tests and no
// branches: a lane whose point is an outlier stores to the first free index behind this batch's inliers (dead space
// between the write front and the read front), which the next batch or the padding overwrites.
template <int WMODE>
__device__ __forceinline__ int fast_mask_and_compact(const KParams& kp, int obj, float* slot, int P, int lane, float thr_u,
                                                     float thr_v, bool all_inliers) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    constexpr int NPL = 5 + WC;   // planes of the slot: X Y Z | u v | w...
    constexpr int U = 4;
    if (all_inliers || kp.inl_in || !(kp.istd_thres > 0.f))
        return fast_compact_rows<WMODE>(kp, obj, slot, lane, thr_u, thr_v, all_inliers || !kp.inl_in, 0, 0, 0u);
    const int full_rows = P >> 5;
    uint32_t out_word = 0u;
    int base = 0, k0 = 0;
    float* row = slot + lane;
#pragma unroll 1
    for (; k0 + U <= full_rows; k0 += U, row += 32 * U) {
        float val[U][NPL];
        int dst[U];
        bool inl[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int c = 0; c < NPL; ++c) val[u][c] = row[c * P + 32 * u];
            inl[u] = (val[u][5] >= thr_u) && (val[u][NPL - 1] >= thr_v);
            const unsigned m = __ballot_sync(kFull, inl[u]);
            out_word = (lane == k0 + u) ? m : out_word;
            dst[u] = base + __popc(m & ((1u << lane) - 1u));
            base += __popc(m);
        }
        __syncwarp();  // the reads of these rows (all lanes) happen before any lane's compacted writes
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (inl[u]) {   // predicated stores
                float* d = slot + dst[u];
#pragma unroll
                for (int c = 0; c < NPL; ++c) d[c * P] = val[u][c];
            }
        }
        // (no barrier after the stores: they land at or below this batch's rows, never ahead of the read front)
    }
    if (k0 * 32 < P) return fast_compact_rows<WMODE>(kp, obj, slot, lane, thr_u, thr_v, false, k0, base, out_word);
    __syncwarp();
    const int rows = (P + 31) >> 5;
    if (kp.inl_out && lane < rows) kp.inl_out[(size_t)obj * rows + lane] = out_word;
    return base;
}

// Null points [n, n_pad): a copy of point 0's coordinates with zero weights, so that the packed loops need no validity
// predicates (M = 0: no contribution to any sum; finite projection: no spurious clip flag beyond point 0's own).
template <int WMODE>
__device__ __forceinline__ void fast_pad(float* slot, int P, int n, int n_pad, int lane) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    float v[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) v[c] = slot[c * P];
    for (int i = n + lane; i < n_pad; i += 32) {
#pragma unroll
        for (int c = 0; c < 5; ++c) slot[c * P + i] = v[c];
#pragma unroll
        for (int c = 0; c < WC; ++c) slot[(5 + c) * P + i] = 0.f;
    }
    __syncwarp();
}

// Out-of-line wrapper of the on-device linear initialiser (cold for callers that pass init_pose)
// gate > 0: trimmed fit around the pose currently in scratch[20..23] (see linear_init_impl).
template <int WMODE>
__device__ __noinline__ bool fast_linear_init(const KParams& kp, int obj, float* slot, int n, int lane, float* scratch,
                                              float gate = 0.f) {
    const int P = kp.n_pts;
    const Camera<float> cam = load_camera<float>(kp, obj);
    double x[4];
    float prior[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) prior[i] = gate > 0.f ? scratch[20 + i] : 0.f;
    __syncwarp();
    const bool ok = linear_init_impl<WMODE, MRPNP_LAYOUT_PLANAR>(slot, slot + 3 * P, slot + 5 * P, P, TeamRows(n, 0, 0, 0, 1),
                                                                 lane, cam, scratch, x, gate > 0.f ? prior : nullptr, gate * gate);
    __syncwarp();
    if (lane == 0 && (ok || !(gate > 0.f))) {   // result through the scratch (scratch[20..23]): fp32 hand-over; .py:119-125
#pragma unroll                                  // (a failed TRIMMED fit leaves the pose it started from in place)
        for (int i = 0; i < 4; ++i) scratch[20 + i] = ok ? (float)x[i] : 0.f;
    }
    __syncwarp();
    return ok;
}

// Pose covariance (fp64, once per object) and the result row.  H = J^T J at the returned x; no point is clipped there
// (or the object would have been handed back), so this is both the pipeline covariance (hessian.py:67-87,
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
__device__ __noinline__ float2 sincos_cold(float d) {   // (sin d, cos d - 1), by value
    float sd, cd;
    sincosf(d, &sd, &cd);
    return make_float2(sd, cd - 1.f);
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
        const float2 sc = sincos_cold(d);
        sd = sc.x; cdm1 = sc.y;
    }
}

constexpr float kFltMax = 3.0e38f;
__device__ __forceinline__ float fast_sqrtf(float a) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
    return y;
}

// ------------------------------------------------------------------ the kernel
// PCT: points per object known at compile time (784 = 28 x 28, every reference config) or 0 = kp.n_pts (even); with a
// compile-time P the plane offsets of the slot become immediates of the shared-memory accesses.
// Owner side of a split delta pass (see FastTeam).  The arguments are in the warp's stash.  Returns jfinite | flagged << 1 |
// barrier phase << 2.
template <int WMODE>
__device__ __noinline__ int team_delta_pass(FastTeam* ft, float* hdr, float* slot, int P, int n_main, int n, int lane, int warp,
                                            uint32_t phase, float* cand) {
    const int nw = blockDim.x >> 5;
    const float* args = hdr + kFastArgStash;
    const float* consts = hdr + kFastConsts;
    float* parts = hdr + kFastParts;
    float* red = hdr + kFastReduce;
    int idle = 0;
    if (lane == 0) idle = *reinterpret_cast<volatile int*>(&ft->idle);
    idle = __shfl_sync(kFull, idle, 0);
    if (nw >= kTeamParts && idle == nw - 1) {   // every other warp of the CTA sits in the waiting room: four warps, one part each
        if (lane == 0) {
            ft->cmd = kFtPass; ft->owner = warp; ft->n_main = n_main; ft->P = P; ft->owner_hdr = hdr; ft->owner_slot = slot;
        }
        ft_barrier(ft, phase, lane);
        team_part<WMODE>(args, consts, slot, P, n_main, 0, red, parts, lane);
        ft_barrier(ft, phase, lane);
    } else {
#pragma unroll 1
        for (int v = 0; v < kTeamParts; ++v) team_part<WMODE>(args, consts, slot, P, n_main, v, red, parts + 20 * v, lane);
    }
    __syncwarp();
    if (lane < 16) {
        const float t = (parts[lane] + parts[20 + lane]) + (parts[40 + lane] + parts[60 + lane]);
        cand[lane] = lane == 15 ? 0.f : (((kNegatedSums >> lane) & 1u) ? -t : t);
    }
    bool flagged = (parts[16] + parts[36]) + (parts[56] + parts[76]) != 0.f;
    __syncwarp();
    if (n > n_main) flagged = pass_remainder<WMODE>(slot, P, n_main, n, lane, kPassDelta, args, consts, cand) || flagged;
    const bool jfin = __all_sync(kFull, fabsf(cand[lane & 15]) < 3.0e38f);
    return (jfin ? 1 : 0) | (flagged ? 2 : 0) | (int)(phase << 2);
}

template <int WMODE, int PCT>
__global__ void __launch_bounds__(kFastMaxWarps * 32, 1) pnp_lm_fast_kernel(const __grid_constant__ KParams kp) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    unsigned char* header = smem_raw + (size_t)warp * kFastHeaderBytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(header);
    float* hdr = reinterpret_cast<float*>(header);
    double* scratch64 = reinterpret_cast<double*>(header + kFastScratch64);
    float* arg_stash = hdr + kFastArgStash;
    float* scale_diag = hdr + kFastScaleDiag;   // the warp-uniform LM state that must survive a pass lives here, not in
                                                // registers: 168 registers per thread are not enough for both, and
                                                // local memory is an L2 round trip at this shared-memory carve-out
    float* slot = reinterpret_cast<float*>(smem_raw + (size_t)nwarps * kFastHeaderBytes) + (size_t)warp * kp.slot_floats;
    const int P = PCT ? PCT : kp.n_pts;
    const int max_iter = kp.max_iter > 0 ? kp.max_iter : (kp.max_iter < 0 ? 0 : 50);
    const int pad_cap = P & ~63;   // the padded groups must fit the planes

    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncwarp();
    wait_for_acks(kp);
#ifdef MRPNP_EXP_FAST_TEAM
    __shared__ FastTeam ft;
    uint32_t ft_phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(&ft.bar, (uint32_t)nwarps);
        fence_mbar_init();
        ft.idle = 0;
    }
    __syncthreads();
#endif
    uint32_t parity = 0;
    int pending = kNoPending;
    bool is_redo = false;
    // camera and clip ranges shared by all objects (cam_mats / uv_range of batch size 1): read once
    const bool shared_cam = kp.cam_stride == 0 && kp.range_stride == 0;
    const Camera<float> cam0 = load_camera<float>(kp, 0);

    // The bulk copies of an object are started as soon as the slot is free, i.e. BEFORE the covariance / result row of
    // the previous object, so that part of the staging latency is hidden behind it.
    int obj;
    do { obj = fetch_job<WC>(kp, slot, P, bar, lane, pending, is_redo); } while (obj == -2);
#pragma unroll 1
    while (obj >= 0) {
        if (is_redo) {   // an object the fast path handed back: the exact fp64 routine, start to finish
#ifndef MRPNP_NO_EXACT
            parity = solve_object_exact<false, WMODE, MRPNP_LAYOUT_PLANAR>(kp, obj, slot, bar, parity, scratch64, lane);
#endif
            __syncwarp();
            do { obj = fetch_job<WC>(kp, slot, P, bar, lane, pending, is_redo); } while (obj == -2);
            continue;
        }
        TR_DECL
        const Camera<float> camf = shared_cam ? cam0 : load_camera<float>(kp, obj);

        // ---------------- stage + weights + inlier mask + compaction ----------------
        int n = P;
#pragma unroll 1
        for (int attempt = 0; attempt < 2; ++attempt) {
            if (kp.use_tma) {
                if (attempt) {  // re-stage the same object (the first attempt compacted the slot)
                    __syncwarp();
                    if (lane == 0) issue_bulk_copies<WC>(kp, obj, slot, P, bar);
                }
                mbar_wait(bar, parity);
                parity ^= 1u;
            } else {
                stage_object_plain<WC>(kp, obj, slot, lane);
            }
            TR_MARK(0)
            float thr_u, thr_v;
            if (WMODE == MRPNP_W_LOGSTD && kp.dense) {
                fast_dense_decode(kp, obj, slot, lane, hdr + kFastBufA);
                thr_u = hdr[kFastBufA]; thr_v = hdr[kFastBufA + 1];
                __syncwarp();
            } else {
                float su, sv;
                fast_weights<WMODE>(kp, slot + 5 * P, P, lane, su, sv);
                const float invP = 1.f / (float)P;
                thr_u = kp.istd_thres * (su * invP);
                thr_v = kp.istd_thres * (sv * invP);
                __syncwarp();
            }
            // second attempt == pnp_uncert_cpu.py:28-32: <= 4 inliers -> every point is an inlier (slot re-staged)
            const bool all = attempt == 1;
            n = fast_mask_and_compact<WMODE>(kp, obj, slot, P, lane, thr_u, thr_v, all);
            if (all || n > 4) break;
        }
        int n_main = (n + 63) & ~63;
        if (n_main <= pad_cap) fast_pad<WMODE>(slot, P, n, n_main, lane); else n_main = n & ~63;
        TR_MARK(1)

        // ---------------- initial point ----------------
        float x[4], pt[4];
        bool init_ok = true;
        if (kp.init_mode == MRPNP_INIT_GIVEN) {
            const float* ip = kp.init + (size_t)obj * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) pt[i] = __ldg(ip + i);
        } else {
            init_ok = fast_linear_init<WMODE>(kp, obj, slot, n, lane, hdr + kFastBufA);
#pragma unroll
            for (int i = 0; i < 4; ++i) pt[i] = hdr[kFastBufA + 20 + i];
            __syncwarp();
        }
        if (kp.ransac_thres || kp.ransac_ratio > 0.f) {
            // reprojection-threshold consensus with the start pose as the model (consensus_prune, pnp_device.cuh)
            float thr;
            if (kp.dense) {   // uncert_prop_pnp_optimizer.py:86-88 on the analytic RoI grid: v[H-1] - v[0] = (H-1)/H (y2 - y1)
                const float* roi = kp.c2d + (size_t)obj * 4;
                const int Hh = P / kp.roi_w;
                thr = kp.ransac_ratio * (__ldg(roi + 3) - __ldg(roi + 1)) * (float)(Hh - 1) / (float)Hh;
            } else {
                thr = __ldg(kp.ransac_thres + obj);
            }
            if (thr > 0.f) {
                // the model of the consensus is always the linear initialiser's pose (init_pose only says where LM starts)
                bool model_ok = init_ok;
                if (kp.init_mode == MRPNP_INIT_GIVEN) model_ok = fast_linear_init<WMODE>(kp, obj, slot, n, lane, hdr + kFastBufA);
                // graduated trimmed fits make the model robust to gross outliers before anything is dropped
#pragma unroll 1
                for (int round = 0; round < 3 && model_ok; ++round)
                    if (!fast_linear_init<WMODE>(kp, obj, slot, n, lane, hdr + kFastBufA, thr * (round == 0 ? 3.f : round == 1 ? 2.f : 1.4f))) break;
                if (model_ok) {
                    __syncwarp();
                    if (lane < 4) hdr[kFastBufB + lane] = hdr[kFastBufA + 20 + lane];
                    __syncwarp();
                    const int n2 = consensus_prune<WMODE, MRPNP_LAYOUT_PLANAR>(kp, obj, slot, n, lane, hdr + kFastBufB, thr);
                    if (n2 != n) {
                        n = n2;
                        n_main = (n + 63) & ~63;
                        if (n_main <= pad_cap) fast_pad<WMODE>(slot, P, n, n_main, lane); else n_main = n & ~63;
                        model_ok = fast_linear_init<WMODE>(kp, obj, slot, n, lane, hdr + kFastBufA);   // like OpenCV's final fit on the consensus set
                    }
                }
                if (kp.init_mode != MRPNP_INIT_GIVEN) {
                    init_ok = model_ok;
#pragma unroll
                    for (int i = 0; i < 4; ++i) pt[i] = hdr[kFastBufA + 20 + i];
                    __syncwarp();
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = pt[i];
        const float2 sc0 = sincos_cold(pt[0]);
        float sn_x = sc0.x, cs_x = sc0.y + 1.f;
        float sn_p = sn_x, cs_p = cs_x;
        TR_MARK(2)

        // ---------------- Levenberg-Marquardt, Ceres 1.14 TrustRegionMinimizer control flow ----------------
        float cost = 0.f, delta[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) delta[i] = 0.f;
        // the sums of the accepted point (`acc`: J^T r, J^T J) and of the candidate being evaluated (`cand`) alternate
        // between the two buffers of the header; Jacobi scale and LM diagonal sit next to them
        float* acc = hdr + kFastBufA;
        float* cand = hdr + kFastBufB;
        if (lane < 8) scale_diag[lane] = 1.f;
        __syncwarp();
        int term = kNoConvergence, iteration = 0, cost_evals = 0, num_invalid = 0;
        float radius = (float)kInitialRadius, decrease_factor = 2.f, x_norm = 0.f, model_change = 1.f;
        bool reuse_diagonal = false, step_ok = true, first = true, redo = false;
        float prev_step_norm2 = 0.f;   // |step|^2 of the last accepted step (adaptive decision band)
        int redo_why = 0;
        PassArgs pa;
        store_consts(hdr + kFastConsts, make_camn(camf), make_clip_window(camf), lane);
        pa.consts = hdr + kFastConsts;
        pa.check = true;
        pa.anchor = false;
        pa.step = DeltaStep{};
        Extent ext = {0.f, 0.f, 0.f};

#pragma unroll 1
        while (true) {
            // ---- the fused pass at pt: 15 sums, transposed warp reduction, broadcast through the scratch ----
            bool flagged, jfin;
            const bool from_observations = cost_evals < 2;  // initial point (plain fp32), then the fp64 anchor
            if (from_observations) {
                pa.cs = cs_p; pa.sn = sn_p; pa.tx = pt[1]; pa.ty = pt[2]; pa.tz = pt[3];
                pa.anchor = cost_evals == 1;
            } else {
                pa.check = !box_inside_window(args_win(pa), ext, pa.step.cp, pa.step.sp, pa.step.txp, pa.step.typ, pa.step.tzp);
            }
#ifdef MRPNP_EXP_FAST_TEAM
            if (!from_observations && cost_evals >= kTeamFromEval && n_main >= kTeamMinPoints) {
                // a long object: the pass in four parts with a fixed association, by four warps when the CTA is otherwise idle
                stash_args(arg_stash, pa, lane);
                const int rc = team_delta_pass<WMODE>(&ft, hdr, slot, P, n_main, n, lane, warp, ft_phase, cand);
                jfin = (rc & 1) != 0; flagged = (rc & 2) != 0; ft_phase = (uint32_t)(rc >> 2) & 1u;
            } else
#endif
            jfin = run_pass<WMODE>(slot, P, n_main, n, lane, pa, from_observations, cand, arg_stash, hdr + kFastReduce, flagged);
            if (cost_evals == 0) { ext.xm = cand[16]; ext.ym = cand[17]; ext.zm = cand[18]; }
            TR_MARK(from_observations ? 3 : 4)
            if (flagged) { redo = true; redo_why = 1; break; }
            ++cost_evals;
            const float c_term = cand[14];  // first two evaluations: sum |r|^2; afterwards: its change
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
                // cost - candidate cost, and a bound of its own rounding: the first step subtracts two independently
                // rounded sums; a delta pass sums products De (2 M e' - M De) that cancel down to the change
                float cost_change, err;
                if (!cfinite) {
                    cost_change = -kFltMax; err = 0.f;
                } else if (from_observations) {
                    cost_change = cost - 0.5f * c_term;
                    err = kp.band_first * cost;
                } else {
                    cost_change = -0.5f * c_term;
                    err = fmaf(kp.band_rel, fabsf(cost_change), kp.band_mix * fast_sqrtf(fabsf(model_change) * cost));
                    // Objects that converge slowly (many evaluations, consecutive steps of similar size): the deviation
                    // of this trajectory from the oracle's, ~1e-6 of the previous step, is small against the current
                    // step, so the band may shrink in proportion to |previous step| / |this step| -- and these are the
                    // objects that are expensive to solve twice.
                    if (kp.band_ratio > 0.f && cost_evals >= kp.band_ratio_from) {
                        const float r = kp.band_ratio * fast_sqrtf(prev_step_norm2 * fast_rcp(fmaxf(step_norm2, 1e-30f)));
                        err *= fminf(1.f, fmaxf(kp.band_rel_min, r) * fast_rcp(kp.band_rel));
                    }
                }
                const float ftol_cost = (float)kFunctionTol * cost;
                // a decision within the band of its threshold is not ours to take: the exact routine solves the object.
                // (The accept test only matters when the function-tolerance test has not ended the solve.)
                if (fabsf(fabsf(cost_change) - ftol_cost) <= err) { redo = true; redo_why = from_observations ? 2 : 3; break; }
                if (fabsf(cost_change) > ftol_cost && fabsf(cost_change - (float)kMinRelDecrease * model_change) <= err) {
                    redo = true;
                    redo_why = 4;
                    break;
                }
                bool stop_after = false;
                if (fabsf(cost_change) <= ftol_cost) {
                    term = kConvergence;
                    if (!(kp.adopt_ftol && cost_change > 0.f)) break;
                    stop_after = true;  // documented switch: take the candidate, then stop
                }
                const float rho = cost_change * fast_rcp(model_change);
                if (stop_after || rho > (float)kMinRelDecrease) {  // HandleSuccessfulStep
                    if (!jfinite) { term = kFailure; break; }
                    accept = true;
                    prev_step_norm2 = step_norm2;
                    cost = from_observations ? 0.5f * c_term : cost - cost_change;
                    const float q = 2.f * rho - 1.f;
                    radius = fminf((float)kMaxRadius, radius * fast_rcp(fmaxf(1.f / 3.f, 1.f - q * q * q)));
                    decrease_factor = 2.f;
                    reuse_diagonal = false;
                    if (stop_after) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) x[i] = pt[i];
                        float* t = acc; acc = cand; cand = t;
                        break;
                    }
                } else {  // HandleUnsuccessfulStep
                    radius = radius * fast_rcp(decrease_factor);
                    decrease_factor *= 2.f;
                    TR_MARK(5)
                    // the slot holds the residuals at the rejected candidate (also after the anchor evaluation)
                    stash_args(arg_stash, pa, lane);
                    undo_pass<WMODE>(slot, P, n_main, n, lane, arg_stash, pa.consts);
                    TR_MARK(6)
                }
            }
            if (accept) {
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = pt[i];
                float* t = acc; acc = cand; cand = t;   // the candidate's sums become the accepted point's
                sn_x = sn_p; cs_x = cs_p;
                x_norm = fast_sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
                step_ok = true;
                if (cost_evals == 1) {  // jacobi_scaling from the initial Jacobian only
                    if (lane < 4) scale_diag[lane] = fast_rcp(1.f + fast_sqrtf(acc[4 + ((0x9740 >> (4 * lane)) & 15)]));   // tri(i,i) = 0,4,7,9
                    __syncwarp();
                }
            }
            // ---- next trust-region step (invalid steps shrink the radius without a new evaluation) ----
            bool stop = false;
#pragma unroll 1
            while (true) {
                // FinalizeIterationAndCheckIfMinimizerCanContinue
                if (iteration >= max_iter) { term = kNoConvergence; stop = true; break; }
                float g[4], H[10], scale[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { g[i] = acc[i]; scale[i] = scale_diag[i]; }
#pragma unroll
                for (int i = 0; i < 10; ++i) H[i] = acc[4 + i];
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
                float diag[4];
                if (!reuse_diagonal) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) diag[i] = fminf(fmaxf(A[tri(i, i)], (float)kMinLmDiag), (float)kMaxLmDiag);
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) scale_diag[4 + i] = diag[i];
                    }
                    __syncwarp();
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) diag[i] = scale_diag[4 + i];
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
            pa.step.cp = cs_p; pa.step.sp = sn_p;
            pa.step.txp = pt[1]; pa.step.typ = pt[2]; pa.step.tzp = pt[3];
            pa.step.ncdm1 = -cdm1; pa.step.sd = sd;
            pa.step.dtx = delta[1]; pa.step.dty = delta[2]; pa.step.dtz = delta[3];
            TR_MARK(5)
        }
        // the slot is free: hand the object back if the fp32 path must not decide it, and start staging the next one
        // before finishing this one
        __syncwarp();
        const int this_obj = obj;
        if (redo) {
            hand_back(kp, this_obj, lane);
            // diagnostic: why (1 clip proximity, 2 first-step band, 3 function-tolerance band, 4 accept band) and after
            // how many evaluations
            if (kp.hand_back_log && lane == 0) kp.hand_back_log[this_obj] = redo_why | (cost_evals << 8);
        }
        bool next_redo;
        int next_obj;
        do { next_obj = fetch_job<WC>(kp, slot, P, bar, lane, pending, next_redo); } while (next_obj == -2);
        if (!redo) {
            // ---------------- pose covariance + result row (out of line, fp64) ----------------
            if (lane == 0) {  // every lane holds the same x
#pragma unroll
                for (int i = 0; i < 4; ++i) cand[i] = x[i];
            }
            __syncwarp();
            fast_finish_object(kp, this_obj, lane, acc + 4, cand, cost, radius, iteration, cost_evals, term);
            TR_MARK(7)
#ifdef MRPNP_TRACE
            if (kp.result64 && lane == 0) {
                double* tr = kp.result64 + (size_t)this_obj * 32;
                for (int i = 0; i < 8; ++i) tr[i] = (double)tr_acc[i];
                tr[8] = (double)cost_evals; tr[9] = (double)(tr_t - tr_t0); tr[10] = (double)n;
                tr[11] = (double)tr_t0; tr[12] = (double)tr_t; tr[13] = (double)blockIdx.x; tr[14] = (double)warp;
            }
#endif
        }
        obj = next_obj;
        is_redo = next_redo;
    }

#ifdef MRPNP_EXP_FAST_TEAM
    // ---------------- waiting room: out of fresh objects ----------------
    // A warp that still works on a long object may call on the warps waiting here for the parts of its delta passes
    // (team_delta_pass) once it is the only busy warp of the CTA; the last warp to arrive lets everybody go on.
    {
        int order = 0;
        if (lane == 0) order = atomicAdd(&ft.idle, 1) + 1;
        order = __shfl_sync(kFull, order, 0);
        if (order == nwarps) {
            if (lane == 0) ft.cmd = kFtExit;
            ft_barrier(&ft, ft_phase, lane);
        } else {
            while (true) {
                ft_barrier(&ft, ft_phase, lane);
                if (ft.cmd == kFtExit) break;
                const int r = (warp - ft.owner + nwarps) % nwarps;
                if (r < kTeamParts)
                    team_part<WMODE>(ft.owner_hdr + kFastArgStash, ft.owner_hdr + kFastConsts, ft.owner_slot, ft.P, ft.n_main, r,
                                     hdr + kFastReduce, ft.owner_hdr + kFastParts + 20 * r, lane);
                ft_barrier(&ft, ft_phase, lane);
            }
        }
    }
#endif

#if !defined(MRPNP_EXP_INLINE_REDO) && !defined(MRPNP_NO_EXACT)
    // ---------------- redo phase: the CTA's warps solve handed-back objects TOGETHER ----------------
    // Every warp of this CTA is out of fresh objects.  Most CTAs get here long before the launch ends (its tail is a few
    // long objects on a few SMs), so the handed-back objects -- appended to the list by whichever warp met them -- are
    // solved here by otherwise idle SMs, eight warps to an fp64 evaluation (RedoTeam, pnp_device.cuh).  A CTA leaves
    // when the list is empty AND every CTA has entered this phase (nobody can append any more).
    // Only the last kRedoKeepers CTAs to get here wait for that; an earlier one leaves as soon as it finds the list empty
    // -- later hand-backs are taken by the keepers -- so that its SM is free for the next launch on another stream.
    __shared__ RedoTeam team;
    __shared__ int keeper;
    __shared__ uint32_t leader_phase;
    uint32_t team_phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(&team.bar, (uint32_t)nwarps);
        fence_mbar_init();
        leader_phase = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int order = atomicAdd(kp.counters + kCntCtasInRedo, 1);
        keeper = order >= (int)gridDim.x - kRedoKeepers;
    }
    while (true) {
        if (threadIdx.x == 0) {
            int got = -1;
            const unsigned long long t0 = global_timer_ns();
            while (true) {
                const int all_in = !keeper || ld_acquire(kp.counters + kCntCtasInRedo) == (int)gridDim.x;
                const int rt = ld_relaxed(kp.counters + kCntRedoTaken), rc = ld_relaxed(kp.counters + kCntRedoCount);
                if (rt < rc) {
                    if (atomicCAS(kp.counters + kCntRedoTaken, rt, rt + 1) != rt) continue;
                    volatile int* e = kp.redo_list + rt;   // entries are object + 1; 0 = reserved but not yet written
                    int v;
                    while ((v = *e) == 0) {}
                    *e = 0;                                // the reader re-arms the entry
                    got = v - 1;
                    break;
                }
                if (all_in) break;                         // read BEFORE the counts: no append can follow
                __nanosleep(256);
                if (global_timer_ns() - t0 > 2000000000ull) break;   // never hang the device
            }
            team.obj = got;
        }
        __syncthreads();
        const int robj = team.obj;
        if (robj < 0) break;
        if (warp == 0) {
            // the leader's round count lives in shared memory across the call (a local passed by address to the
            // out-of-line routine would live in local memory)
            parity = solve_object_exact<false, WMODE, MRPNP_LAYOUT_PLANAR, true>(kp, robj, slot, bar, parity, scratch64, lane, &team,
                                                                                 &leader_phase);
            __syncwarp();
            team_phase = leader_phase;
            if (lane == 0) team.cmd = kTeamDone;
            team_barrier(&team, team_phase, lane);         // releases the workers
            __syncwarp();
            if (lane == 0) leader_phase = team_phase;
        } else {
            team_phase = team_worker<WMODE, MRPNP_LAYOUT_PLANAR>(kp, &team, team_phase, scratch64, warp, lane);
        }
        __syncthreads();
    }
#endif

    // self-resetting work counters: the last CTA to finish rearms them for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        if (kp.n_peers) __threadfence_system();  // rows stored into peer memory are performed before the kernel ends
        __threadfence();
        const int done = atomicAdd(kp.counters + kCntCtasDone, 1);
        if (done == (int)gridDim.x - 1) {
            if (kp.stats) atomicAdd(kp.stats, (unsigned long long)kp.counters[kCntRedoCount]);   // objects handed back, running total
#pragma unroll
            for (int i = 0; i < 5; ++i) kp.counters[i] = 0;
            __threadfence();
            raise_peer_flags(kp);
        }
    }
}

}  // namespace mrpnp
