// pnp_device.cuh -- device-side building blocks of the batched uncertainty-PnP solver (sm_100a).
//
// One warp owns one object: its correspondence slab ([3|2|2or3] x P fp32, 22-25 KB at P=784) is staged
// into the warp's shared-memory slot with 1-D bulk TMA copies (cp.async.bulk + mbarrier), the istd
// inlier test compacts it in place, and every Levenberg-Marquardt pass streams the compacted points
// from shared memory, accumulating cost, J^T r and the upper triangle of J^T J in registers, followed by
// a transposed warp-shuffle reduction.  The 4x4 damped solve and the trust-region bookkeeping run
// redundantly on all lanes in fp64 (no divergence, no broadcast).
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
constexpr int kWarpHeaderBytes = 512;      // per warp: mbarrier (8 B) + 40-double scratch at +128 (reductions, cold-path I/O)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxRows = MRPNP_MAX_POINTS / 32;

// Scratch layout used by the out-of-line (cold) routines: no pointer-to-local crosses a call, because with
// the maximum shared-memory carve-out L1 is ~3 KB and local-memory traffic would go to L2.
constexpr int kScrSums = 0;    // [0..15] reduction scratch / the 16 sums returned by eval_pass_fp64
constexpr int kScrPt = 16;     // [16..19] evaluation point handed to eval_pass_fp64 / returned by linear_init
constexpr int kScrClip = 20;   // [20] clip flag returned by eval_pass_fp64

struct KParams {
    const float* c3d;
    const float* c2d;
    const float* wgt;
    const float* cam;
    const float* range;
    const float* init;
    const uint32_t* inl_in;   // packed: word k of object n = inlier bits of points 32k..32k+31
    float* result;
    uint32_t* inl_out;        // packed, same layout
    double* result64;
    int* counters;            // work counters of one launch (self-resetting; see pnp_kernel_fast.cuh for the fast kernel's five)
    int n_obj, n_pts, cam_stride, range_stride;
    int cov_mode, init_mode, inlier_opt_only, max_iter, adopt_ftol;
    int use_tma, slot_floats;
    float z_min, std_scale, istd_thres;
    // fused head -> PnP entry (mrpnp_solve_dense): c3d = noc_pred [N,3,P], wgt = proj_logstd [N,2,P], c2d = rois [N,4]
    int dense, roi_w;
    const float* dims;      // [N,3] dimensions (l,h,w): decoded, or encoded when dim_means is set
    const float* dims_var;  // [N,3] or NULL
    const float* dim_means; // [n_dim_classes,3] or NULL: MultiClassNormDimCoder.decode in the prologue
    const float* dim_stds;
    const long long* dim_labels;
    int n_dim_classes;
    float* dims_out;        // [N,3] or NULL: decoded dimensions
    float* dims_var_out;
    float noc_mean[3], noc_std[3];
    const float* distance;  // [N] or NULL
    const long long* labels;  // [N] class ids or NULL; used when pred_stride != 0 (c3d/wgt point into all_pred)
    long long pred_stride;    // floats between consecutive objects of all_pred, 0 = pre-sliced maps
    float proj_gain2;       // (ref_focal_y * epistemic_std_gain / scaling_denominator)^2
    float inv_scaling_denominator, distance_min;
    // MRPNP_PREC_FAST: objects the fp32 path must not decide (a point near a clip bound, a decision within the rounding
    // band of its threshold) are handed back through this list (entries object + 1, 0 = empty) and solved by the exact
    // fp64 routine inside the same launch
    int* redo_list;
    unsigned long long* stats;   // [0] running total of handed-back objects (mrpnp_handed_back_count)
    float band_first, band_rel, band_mix;   // half-widths of the decision bands (mrpnp_params)
    float band_ratio, band_rel_min;
    int band_ratio_from;
    int* hand_back_log;       // [N] or NULL: reason | evaluations << 8 of the objects handed back (diagnostic)
    int global_interleaved;   // the tensors in global memory are [N,P,C] although the slot is planar (staging transposes)
    // reprojection-threshold consensus after the start pose (replaces the inlier refinement of cv2.solvePnPRansac,
    // pnp_uncert_cpu.py:34-51): per-object threshold in pixels, or (fused head entry) ratio * RoI height
    const float* ransac_thres;
    float ransac_ratio;
    // exact kernels: optional work list (NULL = objects 0..n_obj-1)
    const int* work_list;
    int* work_count;
    int prefetch_distance;    // objects between the one a team starts and the one it prefetches into L2
    // fused all-gather of the result rows: with n_peers > 0 the row of local object i is stored into EVERY peer's
    // gathered buffer at row row_offset + i (peer-to-peer stores over NVLink), instead of into `result`
    float* peer[MRPNP_MAX_PEERS];
    int n_peers;
    long long row_offset;
    // completion flags of the fused gather (mrpnp_params.peer_flags / acks)
    unsigned int* peer_flag[MRPNP_MAX_PEERS];
    int flag_slot;
    unsigned int flag_value, ack_value;
    const unsigned int* acks;
};

// ------------------------------------------------------------------ cross-GPU flags (system scope)
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long kFlagTimeoutNs = 1000000000ull;

// Spin until *p >= value (wrap-around safe: the flags are step counters).  False after kFlagTimeoutNs.
__device__ __forceinline__ bool wait_flag(const unsigned int* p, unsigned int value) {
    if ((int)(ld_acquire_sys(p) - value) >= 0) return true;
    const unsigned long long t0 = global_timer_ns();
    while ((int)(ld_acquire_sys(p) - value) < 0) {
        __nanosleep(100);
        if (global_timer_ns() - t0 > kFlagTimeoutNs) return false;
    }
    return true;
}

// First thing a launch with peers does: the consumers of every rank have released the buffers it is about to rewrite.
__device__ __forceinline__ void wait_for_acks(const KParams& kp) {
    if (kp.n_peers == 0 || kp.acks == nullptr) return;
    if ((int)threadIdx.x < kp.n_peers) {
        if (!wait_flag(kp.acks + threadIdx.x, kp.ack_value) && kp.stats) atomicAdd(kp.stats + 1, 1ull);
    }
    __syncthreads();
}

// Last thing the LAST thread block does (its thread 0, after it observed every other block's arrival on the launch's
// counter): the rows of all blocks are performed system-wide, then every rank's flag slot of this rank is raised.
__device__ __forceinline__ void raise_peer_flags(const KParams& kp) {
    if (kp.n_peers == 0 || kp.peer_flag[0] == nullptr) return;
    __threadfence_system();
#pragma unroll 1
    for (int r = 0; r < kp.n_peers; ++r) st_release_sys(kp.peer_flag[r] + kp.flag_slot, kp.flag_value);
}

// Consumer side (mrpnp_gather_wait): one warp.
__global__ void gather_wait_kernel(const unsigned int* flags, int n, unsigned int value, KParams kp_acks, int n_acks,
                                   int ack_slot, unsigned int ack_value, unsigned long long* stats) {
    const int lane = threadIdx.x;
    if (flags && lane < n) {
        if (!wait_flag(flags + lane, value) && stats) atomicAdd(stats + 1, 1ull);
    }
    __syncwarp();
    if (lane < n_acks) {
        __threadfence_system();
        st_release_sys(kp_acks.peer_flag[lane] + ack_slot, ack_value);
    }
}

// The 96-byte result row of one object: lanes 0..23 hold its floats.
__device__ __forceinline__ void store_result_row(const KParams& kp, int obj, int lane, float v) {
    if (lane >= MRPNP_RESULT_STRIDE) return;
    if (kp.n_peers == 0) {
        kp.result[(size_t)obj * MRPNP_RESULT_STRIDE + lane] = v;
    } else {
        const size_t at = (size_t)(kp.row_offset + obj) * MRPNP_RESULT_STRIDE + lane;
#pragma unroll 1
        for (int r = 0; r < kp.n_peers; ++r) kp.peer[r][at] = v;
    }
}

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
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
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

// Transposed all-reduce of 16 per-lane partial sums: 8+4+2+1+1 = 16 shuffles instead of 16 x 5.
// After the exchange lane L holds the total of value L>>1; the totals go through the warp's shared
// scratch so that every lane ends with all 16 (bitwise identical on all lanes, fixed summation tree).
template <typename T>
__device__ __forceinline__ void warp_allreduce16(T v[16], T* scratch, int lane) {
#pragma unroll
    for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const T send = up ? v[i] : v[i + half];
            const T keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, bit);
        }
    }
    v[0] += __shfl_xor_sync(kFull, v[0], 1);
    __syncwarp();
    scratch[lane >> 1] = v[0];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = scratch[i];
}

// fp64 reciprocal / square root: MUFU seed (2^-23) + two Newton steps on the fp64 pipe.
// No slow-path branch; relative error ~1e-16 (not correctly rounded).
__device__ __forceinline__ double fast_rcp(double z) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(z));
    double e = fma(-z, x, 1.0);
    x = fma(x, e, x);
    e = fma(-z, x, 1.0);
    x = fma(x, e, x);
    return x;
}
// One Newton step: relative error (2^-23)^2 ~ 1.4e-14 -- used on the serial 4x4 solve where latency matters.
__device__ __forceinline__ double fast_rcp1(double z) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(z));
    const double e = fma(-z, x, 1.0);
    return fma(x, e, x);
}
__device__ __forceinline__ float fast_rcp(float z) {
    float x;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(z));
    return x;
}
__device__ __forceinline__ double fast_sqrt(double a) {  // a >= 0; returns 0 for a == 0
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double h = 0.5 * y;
    double e = fma(-a * y, h, 0.5);  // 0.5 - 0.5 a y^2
    y = fma(y, e, y);
    h = 0.5 * y;
    e = fma(-a * y, h, 0.5);
    y = fma(y, e, y);
    double s = a * y;
    s = fma(0.5 * y, fma(-s, s, a), s);  // one correction of sqrt itself
    return a > 0.0 ? s : 0.0;
}

// sin/cos of a small angle (|d| <= 0.5): Taylor to d^15 / d^14, remainder < 1e-17; two independent Horner
// chains of 7 fused multiply-adds instead of the ~80-instruction library sincos on the serial LM path.
__device__ __forceinline__ void sincos_small(double d, double* s, double* c) {
    const double d2 = d * d;
    double ps = -1.0 / 1307674368000.0;              // -1/15!
    ps = fma(ps, d2, 1.0 / 6227020800.0);             //  1/13!
    ps = fma(ps, d2, -1.0 / 39916800.0);              // -1/11!
    ps = fma(ps, d2, 1.0 / 362880.0);                 //  1/9!
    ps = fma(ps, d2, -1.0 / 5040.0);                  // -1/7!
    ps = fma(ps, d2, 1.0 / 120.0);                    //  1/5!
    ps = fma(ps, d2, -1.0 / 6.0);                     // -1/3!
    double pc = -1.0 / 87178291200.0;                 // -1/14!
    pc = fma(pc, d2, 1.0 / 479001600.0);              //  1/12!
    pc = fma(pc, d2, -1.0 / 3628800.0);               // -1/10!
    pc = fma(pc, d2, 1.0 / 40320.0);                  //  1/8!
    pc = fma(pc, d2, -1.0 / 720.0);                   // -1/6!
    pc = fma(pc, d2, 1.0 / 24.0);                     //  1/4!
    pc = fma(pc, d2, -0.5);                           // -1/2!
    *s = fma(ps * d2, d, d);
    *c = fma(pc, d2, 1.0);
}
// Same, also returning cos d - 1 with full relative precision.
__device__ __forceinline__ void sincos_small(double d, double* s, double* c, double* cm1) {
    const double d2 = d * d;
    double ps = -1.0 / 1307674368000.0;
    ps = fma(ps, d2, 1.0 / 6227020800.0);
    ps = fma(ps, d2, -1.0 / 39916800.0);
    ps = fma(ps, d2, 1.0 / 362880.0);
    ps = fma(ps, d2, -1.0 / 5040.0);
    ps = fma(ps, d2, 1.0 / 120.0);
    ps = fma(ps, d2, -1.0 / 6.0);
    double pc = -1.0 / 87178291200.0;
    pc = fma(pc, d2, 1.0 / 479001600.0);
    pc = fma(pc, d2, -1.0 / 3628800.0);
    pc = fma(pc, d2, 1.0 / 40320.0);
    pc = fma(pc, d2, -1.0 / 720.0);
    pc = fma(pc, d2, 1.0 / 24.0);
    pc = fma(pc, d2, -0.5);
    *s = fma(ps * d2, d, d);
    *cm1 = pc * d2;
    *c = 1.0 + *cm1;
}

// ------------------------------------------------------------------ which rows of the slot a warp works on
// After compaction the inliers of an object sit in up to two segments of the slot (one per warp that compacted):
// segment 0 = points [0, n0), segment 1 = points [base1, base1 + n1).  Their rows of 32 points are numbered through
// ("virtual rows": rows of segment 0, then rows of segment 1) and dealt round-robin to the `team` warps, so every
// warp always revisits the same points (it owns their tracked residuals) and the load is balanced to one row.
// The warp-per-object kernel uses the degenerate form TeamRows(n, 0, 0, 0, 1).
struct TeamRows {
    int n0, n1, base1, rows0, rows_total, t, team, safe;
    __device__ __forceinline__ TeamRows(int n0_, int n1_, int base1_, int t_, int team_)
        : n0(n0_), n1(n1_), base1(base1_), rows0((n0_ + 31) >> 5), rows_total(((n0_ + 31) >> 5) + ((n1_ + 31) >> 5)),
          t(t_), team(team_), safe(n0_ > 0 ? 0 : base1_) {}
    // number of R-row groups this warp runs
    __device__ __forceinline__ int groups(int R) const {
        const int mine = rows_total > t ? (rows_total - t + team - 1) / team : 0;
        return (mine + R - 1) / R;
    }
    // slot index of lane `lane` of this warp's row r of group g; invalid lanes get a valid index to read (the
    // caller zeroes their weights)
    __device__ __forceinline__ int point(int g, int r, int R, int lane, bool& valid) const {
        const int v = (g * R + r) * team + t;
        const bool s1 = v >= rows0;
        const int k = (s1 ? v - rows0 : v) * 32 + lane;
        valid = (v < rows_total) && (k < (s1 ? n1 : n0));
        return valid ? (s1 ? base1 + k : k) : safe;
    }
};

template <typename T>
struct Camera {
    T fx, fy, cx, cy;
    T z_min, u_min, u_max, v_min, v_max;
};

// ------------------------------------------------------------------ fused pass, fp64 arithmetic (MRPNP_PREC_FP64)
// Accumulates over this lane's share of the n active points and all-reduces over the warp:
//   acc[0]     sum |r|^2
//   acc[1..4]  J^T r        (order yaw, tx, ty, tz)
//   acc[5..14] J^T J upper triangle  (00 01 02 03 11 12 13 22 23 33)
// Every operation is fp64 in the reference's own order of evaluation, so decisions of the LM loop are
// reproduced exactly.
// clipsem 0: Ceres-Jet semantics (pnp_uncert_cpu.cpp:36-42): z clip drops only d/dz', u/v clamp drops that row.
// clipsem 1: jacobian.py:52-59 semantics: a z-clipped point loses both rows (used for the pipeline covariance).
// use_bits : skip points whose bit in `bits` (bit k <-> point 32k+lane) is 0 (outliers when not compacted).
// Evaluation point in: scratch[kScrPt..]; out: 16 sums in scratch[kScrSums..], clip flag in scratch[kScrClip].
// Deliberately not inlined: it is the main
// pass only in MRPNP_PREC_FP64 mode and the cold fallback of the mixed pass, and keeping one copy keeps the
// hot code of the kernel inside the instruction cache.
template <typename T>
__device__ __forceinline__ Camera<T> load_camera(const KParams& kp, int obj) {
    Camera<T> c;
    const float* K = kp.cam + (size_t)obj * kp.cam_stride;
    const float* R = kp.range + (size_t)obj * kp.range_stride;
    c.fx = (T)__ldg(K + 0); c.fy = (T)__ldg(K + 4);  // pnp_uncert_cpu.cpp:265
    c.cx = (T)__ldg(K + 2); c.cy = (T)__ldg(K + 5);
    c.z_min = (T)kp.z_min;
    c.u_min = (T)__ldg(R + 0); c.u_max = (T)__ldg(R + 1);
    c.v_min = (T)__ldg(R + 2); c.v_max = (T)__ldg(R + 3);
    return c;
}

// k0 / kstep: the rows (of 32 points) k0, k0 + kstep, ... only -- a team of warps splits an evaluation this way
// (RedoTeam below); 0 / 1 = every row.
template <int WMODE, int LAYOUT>
__device__ __noinline__ void eval_pass_fp64(const KParams& kp, int obj, const float* __restrict__ slot, int n, int lane,
                                            uint32_t bits, int clipsem, bool use_bits, double* scratch, int k0 = 0,
                                            int kstep = 1) {
    const int P = kp.n_pts;
    const float* __restrict__ s3 = slot;
    const float* __restrict__ s2 = slot + 3 * P;
    const float* __restrict__ sw = slot + 5 * P;
    const Camera<double> cam = load_camera<double>(kp, obj);
    const double x[4] = {scratch[kScrPt], scratch[kScrPt + 1], scratch[kScrPt + 2], scratch[kScrPt + 3]};
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    double sn, cs;
    sincos(x[0], &sn, &cs);
    const double tx = x[1], ty = x[2], tz = x[3];
    double acc[16];  // register accumulators; copied to acc_out after the reduction
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.0;
    bool any = false;
#pragma unroll 2
    for (int p = lane + 32 * k0, k = k0; p < n; p += 32 * kstep, k += kstep) {
        if (use_bits && !((bits >> k) & 1u)) continue;
        const double X = (double)s3[sidx<LAYOUT, 3>(p, 0, P)];
        const double Y = (double)s3[sidx<LAYOUT, 3>(p, 1, P)];
        const double Z = (double)s3[sidx<LAYOUT, 3>(p, 2, P)];
        const double uo = (double)s2[sidx<LAYOUT, 2>(p, 0, P)];
        const double vo = (double)s2[sidx<LAYOUT, 2>(p, 1, P)];
        const double qx = fma(cs, X, sn * Z);
        const double qz = fma(cs, Z, -sn * X);
        const double xc = qx + tx, yc = Y + ty, zc = qz + tz;
        const bool zfree = !(zc < cam.z_min);
        const double z = zfree ? zc : cam.z_min;
        const double iz = fast_rcp(z);
        const double xn = xc * iz, yn = yc * iz;
        double pu = fma(cam.fx, xn, cam.cx);
        double pv = fma(cam.fy, yn, cam.cy);
        bool ufree = true, vfree = true;
        if (pu < cam.u_min) { pu = cam.u_min; ufree = false; } else if (pu > cam.u_max) { pu = cam.u_max; ufree = false; }
        if (pv < cam.v_min) { pv = cam.v_min; vfree = false; } else if (pv > cam.v_max) { pv = cam.v_max; vfree = false; }
        any = any || !zfree || !ufree || !vfree;
        const double du = pu - uo, dv = pv - vo;
        const double mz = zfree ? 1.0 : 0.0;
        if (clipsem == 1 && !zfree) { ufree = false; vfree = false; }
        // unweighted projection Jacobian rows: Ju = (ju0, a_u, 0, b_u), Jv = (jv0, 0, a_v, b_v)
        const double au = ufree ? cam.fx * iz : 0.0;
        const double av = vfree ? cam.fy * iz : 0.0;
        const double bu = -au * xn * mz, bv = -av * yn * mz;
        const double ju0 = fma(au, qz, -bu * qx);
        const double jv0 = -bv * qx;
        if (WMODE != MRPNP_W_FULL) {
            const double wu = (double)sw[sidx<LAYOUT, WC>(p, 0, P)];
            const double wv = (double)sw[sidx<LAYOUT, WC>(p, 1, P)];
            const double ru = wu * du, rv = wv * dv;
            const double a0 = wu * ju0, a1 = wu * au, a3 = wu * bu;   // row u: (a0, a1, 0, a3)
            const double b0 = wv * jv0, b2 = wv * av, b3 = wv * bv;   // row v: (b0, 0, b2, b3)
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
            const double wxx = (double)sw[sidx<LAYOUT, WC>(p, 0, P)];
            const double wxy = (double)sw[sidx<LAYOUT, WC>(p, 1, P)];
            const double wyy = (double)sw[sidx<LAYOUT, WC>(p, 2, P)];
            const double r0 = fma(wxx, du, wxy * dv), r1 = fma(wxy, du, wyy * dv);
            // whitened rows: r0-row = wxx*Ju + wxy*Jv, r1-row = wxy*Ju + wyy*Jv
            const double a0 = fma(wxx, ju0, wxy * jv0), a1 = wxx * au, a2 = wxy * av, a3 = fma(wxx, bu, wxy * bv);
            const double b0 = fma(wxy, ju0, wyy * jv0), b1 = wxy * au, b2 = wyy * av, b3 = fma(wxy, bu, wyy * bv);
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
    any = __any_sync(kFull, any);
    warp_allreduce16<double>(acc, scratch + kScrSums, lane);  // leaves the 16 totals in scratch[0..15]
    if (lane == 0) scratch[kScrClip] = any ? 1.0 : 0.0;
    __syncwarp();
}

// ------------------------------------------------------------------ a CTA's warps evaluating ONE object together
// Used by the MRPNP_PREC_FAST kernel for the objects it hands back: when all warps of a CTA have run out of fresh objects
// they solve handed-back objects together -- warp 0 runs solve_object_exact, and each of its fp64 evaluations is split by
// rows over all warps (warp w takes rows w, w + W, ...; partial totals through shared memory, added in warp order, so
// the result does not depend on timing).  Two named-barrier rounds per evaluation.
struct RedoTeam {
    double part[kMaxWarpsPerCta][18];   // per warp: 16 totals, [16] clip flag
    double pt[4];
    uint64_t bar;                       // mbarrier, one arrival per warp
    float* slot;
    int cmd, n, clipsem, obj;
};
enum { kTeamEval = 0, kTeamDone = 1 };
// All warps of the CTA meet here, leader and workers from DIFFERENT code: a named barrier in its non-aligned form
// (`bar.sync` = barrier.sync.aligned promises that every thread executes the same barrier instruction, which synccheck
// rightly rejects here).  MRPNP_TEAM_MBARRIER builds the equivalent on an mbarrier with one arrival per warp -- `phase` is
// each thread's own count of the rounds -- which racecheck does not model (it reports the hand-overs as hazards).
__device__ __forceinline__ void cta_round(int id, uint64_t* mbar, uint32_t& phase, int lane) {
#ifdef MRPNP_TEAM_MBARRIER
    __syncwarp();
    if (lane == 0) {
        asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(mbar)) : "memory");
    }
    mbar_wait(mbar, phase);
    phase ^= 1u;
#else
    __syncwarp();
    asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"((int)blockDim.x) : "memory");
#endif
}
__device__ __forceinline__ void team_barrier(RedoTeam* team, uint32_t& phase, int lane) { cta_round(1, &team->bar, phase, lane); }

// Leader side of one evaluation (same contract as eval_pass_fp64 with use_bits = false).
template <int WMODE, int LAYOUT>
__device__ __forceinline__ void team_eval_fp64(const KParams& kp, RedoTeam* team, uint32_t& phase, int obj, float* slot, int n,
                                               int lane, int clipsem, double* scratch) {
    const int nw = blockDim.x >> 5;
    if (lane < 4) team->pt[lane] = scratch[kScrPt + lane];
    if (lane == 0) { team->cmd = kTeamEval; team->n = n; team->clipsem = clipsem; team->slot = slot; }   // team->obj: set by the fetch
    team_barrier(team, phase, lane);                                      // workers start
    eval_pass_fp64<WMODE, LAYOUT>(kp, obj, slot, n, lane, 0u, clipsem, false, scratch, 0, nw);
    if (lane < 16) team->part[0][lane] = scratch[kScrSums + lane];
    if (lane == 16) team->part[0][16] = scratch[kScrClip];
    team_barrier(team, phase, lane);                                      // all partial totals are written
    if (lane < 17) {
        double t = team->part[0][lane];
        for (int w = 1; w < nw; ++w) t += team->part[w][lane];
        if (lane < 16) scratch[kScrSums + lane] = t; else scratch[kScrClip] = t != 0.0 ? 1.0 : 0.0;
    }
    __syncwarp();
}

// Worker side: warps 1.. of the CTA while warp 0 solves an object; returns when the leader posts kTeamDone.
template <int WMODE, int LAYOUT>
__device__ __noinline__ uint32_t team_worker(const KParams& kp, RedoTeam* team, uint32_t phase, double* scratch, int warp,
                                             int lane) {
    const int nw = blockDim.x >> 5;
    while (true) {
        team_barrier(team, phase, lane);
        if (team->cmd == kTeamDone) break;
        if (lane < 4) scratch[kScrPt + lane] = team->pt[lane];
        __syncwarp();
        eval_pass_fp64<WMODE, LAYOUT>(kp, team->obj, team->slot, team->n, lane, 0u, team->clipsem, false, scratch, warp, nw);
        if (lane < 16) team->part[warp][lane] = scratch[kScrSums + lane];
        if (lane == 16) team->part[warp][16] = scratch[kScrClip];
        team_barrier(team, phase, lane);
    }
    return phase;
}

// ------------------------------------------------------------------ fused pass, mixed precision (MRPNP_PREC_MIXED)
// The residual chain (rotation, projection, pixel difference, weighting) and the cost sum run in fp64 on
// the fp64 pipe, so cost and cost differences -- what the trust-region accept / function-tolerance tests
// read -- agree with the fp64 reference to ~1e-15.  The Jacobian, J^T r and J^T J are built from an
// independent fp32 projection on the FMA pipe (errors ~1e-7 relative, which only perturb the step).
// If any point comes within a small margin of a clip bound (never on in-range data) `flagged_out` is set
// and the caller redoes the whole pass with the exact fp64 routine.  Output layout as eval_pass_fp64.
// (An invalid padding lane re-reads point n-1 with zero weights, so it can only raise the flag if that
// point itself is near a clip.)
template <int WMODE, int LAYOUT>
__device__ __forceinline__ void eval_pass_mixed(const float* __restrict__ s3, const float* __restrict__ s2,
                                                const float* __restrict__ sw, int P, int n, int lane,
                                                const double x[4], double sn, double cs,
                                                const Camera<double>& cam, const Camera<float>& camf,
                                                double acc[16], double* scratch, bool& flagged_out,
                                                bool& jfinite_out) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    constexpr int R = 2;  // rows carried through the dependent chains together (ILP; one row alone issues ~1/8 cycles)
    const double tx = x[1], ty = x[2], tz = x[3];
    const float snf = (float)sn, csf = (float)cs, txf = (float)tx, tyf = (float)ty, tzf = (float)tz;
    // clip detection on the fp32 projection with a safety margin far above its rounding error (~1e-4 px)
    const float zlo = camf.z_min * 1.001f + 1e-3f;
    const float ulo = camf.u_min + 0.05f, uhi = camf.u_max - 0.05f, vlo = camf.v_min + 0.05f, vhi = camf.v_max - 0.05f;
    double cost[R];
#pragma unroll
    for (int r = 0; r < R; ++r) cost[r] = 0.0;
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 0.f;
    float margin = 1e30f;  // min over points of the distance to the nearest clip bound (negative = flagged)
    const int rows = (n + 31) >> 5;
    for (int k0 = 0; k0 < rows; k0 += R) {
        float Xf[R], Yf[R], Zf[R], uf[R], vf[R], w0[R], w1[R], w2[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int pr = (k0 + r) * 32 + lane;
            const bool valid = pr < n;
            const int p = valid ? pr : n - 1;
            Xf[r] = s3[sidx<LAYOUT, 3>(p, 0, P)]; Yf[r] = s3[sidx<LAYOUT, 3>(p, 1, P)]; Zf[r] = s3[sidx<LAYOUT, 3>(p, 2, P)];
            uf[r] = s2[sidx<LAYOUT, 2>(p, 0, P)]; vf[r] = s2[sidx<LAYOUT, 2>(p, 1, P)];
            w0[r] = sw[sidx<LAYOUT, WC>(p, 0, P)]; w1[r] = sw[sidx<LAYOUT, WC>(p, 1, P)];
            w2[r] = (WMODE == MRPNP_W_FULL) ? sw[sidx<LAYOUT, WC>(p, WC - 1, P)] : 0.f;
            if (!valid) { w0[r] = 0.f; w1[r] = 0.f; w2[r] = 0.f; }  // padding lanes / rows contribute nothing
        }
        // ---- fp64 residual chain ----
        double xc[R], yc[R], zc[R], iz[R], du[R], dv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double X = (double)Xf[r], Y = (double)Yf[r], Z = (double)Zf[r];
            xc[r] = fma(cs, X, fma(sn, Z, tx));
            zc[r] = fma(cs, Z, fma(-sn, X, tz));
            yc[r] = Y + ty;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) iz[r] = fast_rcp(zc[r]);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            du[r] = fma(cam.fx, xc[r] * iz[r], cam.cx) - (double)uf[r];
            dv[r] = fma(cam.fy, yc[r] * iz[r], cam.cy) - (double)vf[r];
        }
        // ---- fp32 projection: Jacobian + clip detection ----
        float qxf[R], qzf[R], izf[R], xnf[R], ynf[R], puf[R], pvf[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            qxf[r] = fmaf(csf, Xf[r], snf * Zf[r]);
            qzf[r] = fmaf(csf, Zf[r], -snf * Xf[r]);
            const float zcf = qzf[r] + tzf;
            margin = fminf(margin, zcf - zlo);
            izf[r] = fast_rcp(zcf);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            xnf[r] = (qxf[r] + txf) * izf[r];
            ynf[r] = (Yf[r] + tyf) * izf[r];
            puf[r] = fmaf(camf.fx, xnf[r], camf.cx);
            pvf[r] = fmaf(camf.fy, ynf[r], camf.cy);
            margin = fminf(margin, fminf(fminf(puf[r] - ulo, uhi - puf[r]), fminf(pvf[r] - vlo, vhi - pvf[r])));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float au = camf.fx * izf[r], av = camf.fy * izf[r];
            const float bu = -au * xnf[r], bv = -av * ynf[r];
            const float ju0 = fmaf(au, qzf[r], -bu * qxf[r]), jv0 = -bv * qxf[r];
            const float duf = puf[r] - uf[r], dvf = pvf[r] - vf[r];
            if (WMODE != MRPNP_W_FULL) {
                const double ru = (double)w0[r] * du[r], rv = (double)w1[r] * dv[r];
                cost[r] = fma(ru, ru, fma(rv, rv, cost[r]));
                const float ruf = w0[r] * duf, rvf = w1[r] * dvf;
                const float a0 = w0[r] * ju0, a1 = w0[r] * au, a3 = w0[r] * bu;
                const float b0 = w1[r] * jv0, b2 = w1[r] * av, b3 = w1[r] * bv;
                a[0] = fmaf(a0, ruf, fmaf(b0, rvf, a[0]));
                a[1] = fmaf(a1, ruf, a[1]);
                a[2] = fmaf(b2, rvf, a[2]);
                a[3] = fmaf(a3, ruf, fmaf(b3, rvf, a[3]));
                a[4] = fmaf(a0, a0, fmaf(b0, b0, a[4]));
                a[5] = fmaf(a0, a1, a[5]);
                a[6] = fmaf(b0, b2, a[6]);
                a[7] = fmaf(a0, a3, fmaf(b0, b3, a[7]));
                a[8] = fmaf(a1, a1, a[8]);
                a[10] = fmaf(a1, a3, a[10]);
                a[11] = fmaf(b2, b2, a[11]);
                a[12] = fmaf(b2, b3, a[12]);
                a[13] = fmaf(a3, a3, fmaf(b3, b3, a[13]));
            } else {
                const double r0 = fma((double)w0[r], du[r], (double)w1[r] * dv[r]);
                const double r1 = fma((double)w1[r], du[r], (double)w2[r] * dv[r]);
                cost[r] = fma(r0, r0, fma(r1, r1, cost[r]));
                const float r0f = fmaf(w0[r], duf, w1[r] * dvf), r1f = fmaf(w1[r], duf, w2[r] * dvf);
                const float a0 = fmaf(w0[r], ju0, w1[r] * jv0), a1 = w0[r] * au, a2 = w1[r] * av, a3 = fmaf(w0[r], bu, w1[r] * bv);
                const float b0 = fmaf(w1[r], ju0, w2[r] * jv0), b1 = w1[r] * au, b2 = w2[r] * av, b3 = fmaf(w1[r], bu, w2[r] * bv);
                a[0] = fmaf(a0, r0f, fmaf(b0, r1f, a[0]));
                a[1] = fmaf(a1, r0f, fmaf(b1, r1f, a[1]));
                a[2] = fmaf(a2, r0f, fmaf(b2, r1f, a[2]));
                a[3] = fmaf(a3, r0f, fmaf(b3, r1f, a[3]));
                a[4] = fmaf(a0, a0, fmaf(b0, b0, a[4]));
                a[5] = fmaf(a0, a1, fmaf(b0, b1, a[5]));
                a[6] = fmaf(a0, a2, fmaf(b0, b2, a[6]));
                a[7] = fmaf(a0, a3, fmaf(b0, b3, a[7]));
                a[8] = fmaf(a1, a1, fmaf(b1, b1, a[8]));
                a[9] = fmaf(a1, a2, fmaf(b1, b2, a[9]));
                a[10] = fmaf(a1, a3, fmaf(b1, b3, a[10]));
                a[11] = fmaf(a2, a2, fmaf(b2, b2, a[11]));
                a[12] = fmaf(a2, a3, fmaf(b2, b3, a[12]));
                a[13] = fmaf(a3, a3, fmaf(b3, b3, a[13]));
            }
        }
    }
    flagged_out = __any_sync(kFull, !(margin >= 0.f));
    double cost2 = cost[0];
#pragma unroll
    for (int r = 1; r < R; ++r) cost2 += cost[r];
    // cost: 5-step fp64 butterfly; the 14 fp32 sums: transposed reduction through the scratch
    cost2 = warp_sum(cost2);
    warp_allreduce16<float>(a, reinterpret_cast<float*>(scratch), lane);
    // finite iff every Jacobian sum is finite (tree of |.| in fp32 instead of a 14-deep fp64 chain)
    const float s01 = (fabsf(a[0]) + fabsf(a[1])) + (fabsf(a[2]) + fabsf(a[3]));
    const float s23 = (fabsf(a[4]) + fabsf(a[5])) + (fabsf(a[6]) + fabsf(a[7]));
    const float s45 = (fabsf(a[8]) + fabsf(a[9])) + (fabsf(a[10]) + fabsf(a[11]));
    const float s67 = fabsf(a[12]) + fabsf(a[13]);
    jfinite_out = ((s01 + s23) + (s45 + s67)) < 3.0e38f;
    acc[0] = cost2;
#pragma unroll
    for (int i = 0; i < 14; ++i) acc[1 + i] = (double)a[i];
    acc[15] = 0.0;
}

// ------------------------------------------------------------------ 4x4 SPD helpers (fp64, fully unrolled)
// Upper-triangle packing index: (0,0)=0 (0,1)=1 (0,2)=2 (0,3)=3 (1,1)=4 (1,2)=5 (1,3)=6 (2,2)=7 (2,3)=8 (3,3)=9.
__device__ __forceinline__ constexpr int tri(int i, int j) {
    return i <= j ? (i * 4 - (i * (i - 1)) / 2 + (j - i)) : (j * 4 - (j * (j - 1)) / 2 + (i - j));
}

// A = L D L^T of the packed symmetric A (unit lower L, reciprocal pivots), without square roots and
// with fast reciprocals.  ok == false if a pivot is not positive/finite (A not SPD).
struct Ldl4 {
    double l10, l20, l30, l21, l31, l32;
    double i0, i1, i2, i3;
    bool ok;
};
__device__ __forceinline__ Ldl4 ldl4_factor(const double A[10]) {
    Ldl4 f;
    const double d0 = A[0];
    f.i0 = fast_rcp1(d0);
    f.l10 = A[1] * f.i0; f.l20 = A[2] * f.i0; f.l30 = A[3] * f.i0;
    const double d1 = fma(-f.l10, A[1], A[4]);
    f.i1 = fast_rcp1(d1);
    const double t21 = fma(-f.l20, A[1], A[5]), t31 = fma(-f.l30, A[1], A[6]);
    f.l21 = t21 * f.i1; f.l31 = t31 * f.i1;
    const double d2 = fma(-f.l21, t21, fma(-f.l20, A[2], A[7]));
    f.i2 = fast_rcp1(d2);
    const double t32 = fma(-f.l31, t21, fma(-f.l30, A[2], A[8]));
    f.l32 = t32 * f.i2;
    const double d3 = fma(-f.l32, t32, fma(-f.l31, t31, fma(-f.l30, A[3], A[9])));
    f.i3 = fast_rcp1(d3);
    f.ok = (d0 > 0.0) && (d1 > 0.0) && (d2 > 0.0) && (d3 > 0.0) && (d0 < 1.7e308) && (d1 < 1.7e308) &&
           (d2 < 1.7e308) && (d3 < 1.7e308);
    return f;
}
__device__ __forceinline__ void ldl4_solve(const Ldl4& f, const double b[4], double y[4]) {
    const double z0 = b[0];
    const double z1 = fma(-f.l10, z0, b[1]);
    const double z2 = fma(-f.l21, z1, fma(-f.l20, z0, b[2]));
    const double z3 = fma(-f.l32, z2, fma(-f.l31, z1, fma(-f.l30, z0, b[3])));
    y[3] = z3 * f.i3;
    y[2] = fma(-f.l32, y[3], z2 * f.i2);
    y[1] = fma(-f.l31, y[3], fma(-f.l21, y[2], z1 * f.i1));
    y[0] = fma(-f.l30, y[3], fma(-f.l20, y[2], fma(-f.l10, y[1], z0 * f.i0)));
}

// Full inverse of the packed SPD matrix H into row-major inv[16]; false if not SPD.
__device__ __forceinline__ bool spd_inverse4(const double H[10], double inv[16]) {
    const Ldl4 f = ldl4_factor(H);
    if (!f.ok) return false;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double e[4] = {0.0, 0.0, 0.0, 0.0}, y[4];
        e[c] = 1.0;
        ldl4_solve(f, e, y);
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
    double inv[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double d = A[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d = fma(-A[j][k], A[j][k], d);
        ok = ok && (d > 0.0) && (d < 1.7e308);
        const double ld = fast_sqrt(d);
        inv[j] = fast_rcp(ld);
        A[j][j] = ld;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            double v = A[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) v = fma(-A[i][k], A[j][k], v);
            A[i][j] = v * inv[j];
        }
    }
    double w[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double v = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v = fma(-A[i][k], w[k], v);
        w[i] = v * inv[i];
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
        double v = w[i];
#pragma unroll
        for (int k = i + 1; k < N; ++k) v = fma(-A[k][i], x[k], v);
        x[i] = v * inv[i];
    }
    return ok;
}

// fp32 accumulation of the (well-scaled, normalised-coordinate) normal equations, fp64 solves.
// With a `prior` pose (yaw, t) and gate2 > 0 only the points whose reprojection error at the prior is within sqrt(gate2)
// pixels take part: the trimmed fits of the consensus stage (see consensus_prune).
template <int WMODE, int LAYOUT>
__device__ __forceinline__ bool linear_init_impl(const float* __restrict__ s3, const float* __restrict__ s2,
                                            const float* __restrict__ sw, int P, const TeamRows& rows, int lane,
                                            const Camera<float>& cam, float* scratch, double* x,
                                            const float* prior = nullptr, float gate2 = 0.f) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    const float ifx = 1.f / cam.fx, ify = 1.f / cam.fy;
    const bool gated = prior != nullptr && gate2 > 0.f;
    float psn = 0.f, pcs = 1.f, ptx = 0.f, pty = 0.f, ptz = 1.f;
    if (gated) {
        sincosf(prior[0], &psn, &pcs);
        ptx = prior[1]; pty = prior[2]; ptz = prior[3];
    }
    __syncwarp();
    // reprojection error^2 of point (X, Y, Z) -> (u, v) at the prior
    auto outside = [&](float X, float Y, float Z, float u, float v) {
        const float iz = 1.f / fmaxf(fmaf(pcs, Z, fmaf(-psn, X, ptz)), cam.z_min);
        const float du = fmaf(cam.fx, fmaf(pcs, X, fmaf(psn, Z, ptx)) * iz, cam.cx) - u;
        const float dv = fmaf(cam.fy, (Y + pty) * iz, cam.cy) - v;
        return !(fmaf(du, du, dv * dv) <= gate2);
    };
    // ---- stage A: 5 unknowns: 15 matrix entries + 5 rhs = 20 sums -> two transposed reductions ----
    float m[16], r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { m[i] = 0.f; r[i] = 0.f; }
    const int ngroups = rows.groups(1);
    for (int gi = 0; gi < ngroups; ++gi) {
        bool valid;
        const int p = rows.point(gi, 0, 1, lane, valid);
        if (!valid) continue;
        const float X = s3[sidx<LAYOUT, 3>(p, 0, P)], Y = s3[sidx<LAYOUT, 3>(p, 1, P)], Z = s3[sidx<LAYOUT, 3>(p, 2, P)];
        if (gated && outside(X, Y, Z, s2[sidx<LAYOUT, 2>(p, 0, P)], s2[sidx<LAYOUT, 2>(p, 1, P)])) continue;
        const float un = (s2[sidx<LAYOUT, 2>(p, 0, P)] - cam.cx) * ifx;
        const float vn = (s2[sidx<LAYOUT, 2>(p, 1, P)] - cam.cy) * ify;
        const float wu = sw[sidx<LAYOUT, WC>(p, 0, P)] * cam.fx, wv = sw[sidx<LAYOUT, WC>(p, WC - 1, P)] * cam.fy;
        const float a[5] = {wu * (X - un * Z), wu * (Z + un * X), wu, 0.f, -wu * un};
        const float b[5] = {-wv * vn * Z, wv * vn * X, 0.f, wv, -wv * vn};
        const float rb = -wv * Y;
        int q = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
            for (int j = i; j < 5; ++j, ++q) m[q] = fmaf(a[i], a[j], fmaf(b[i], b[j], m[q]));
            r[i] = fmaf(b[i], rb, r[i]);
        }
    }
    warp_allreduce16<float>(m, scratch, lane);
    warp_allreduce16<float>(r, scratch, lane);
    double A[5][5], rhs[5], sol[5];
    {
        int q = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
            for (int j = i; j < 5; ++j, ++q) {
                A[i][j] = (double)m[q];
                A[j][i] = (double)m[q];
            }
            rhs[i] = (double)r[i];
        }
    }
    if (!chol_solve_dense<5>(A, rhs, sol)) return false;
    const double yaw = atan2(sol[1], sol[0]);
    // ---- stage B: translation with yaw fixed ----
    double snd, csd;
    sincos(yaw, &snd, &csd);
    const float sn = (float)snd, cs = (float)csd;
    float m3[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m3[i] = 0.f;
    for (int gi = 0; gi < ngroups; ++gi) {
        bool valid;
        const int p = rows.point(gi, 0, 1, lane, valid);
        if (!valid) continue;
        const float X = s3[sidx<LAYOUT, 3>(p, 0, P)], Y = s3[sidx<LAYOUT, 3>(p, 1, P)], Z = s3[sidx<LAYOUT, 3>(p, 2, P)];
        if (gated && outside(X, Y, Z, s2[sidx<LAYOUT, 2>(p, 0, P)], s2[sidx<LAYOUT, 2>(p, 1, P)])) continue;
        const float un = (s2[sidx<LAYOUT, 2>(p, 0, P)] - cam.cx) * ifx;
        const float vn = (s2[sidx<LAYOUT, 2>(p, 1, P)] - cam.cy) * ify;
        const float wu = sw[sidx<LAYOUT, WC>(p, 0, P)] * cam.fx, wv = sw[sidx<LAYOUT, WC>(p, WC - 1, P)] * cam.fy;
        const float qx = fmaf(cs, X, sn * Z), qz = fmaf(cs, Z, -sn * X);
        // rows: wu [1 0 -un] t = -wu (qx - un qz);  wv [0 1 -vn] t = -wv (Y - vn qz)
        const float ra = -wu * (qx - un * qz), rb = -wv * (Y - vn * qz);
        const float a2 = -wu * un, b2 = -wv * vn;
        m3[0] = fmaf(wu, wu, m3[0]);                  // (0,0)
        m3[1] = fmaf(wu, a2, m3[1]);                  // (0,2)
        m3[2] = fmaf(wv, wv, m3[2]);                  // (1,1)
        m3[3] = fmaf(wv, b2, m3[3]);                  // (1,2)
        m3[4] = fmaf(a2, a2, fmaf(b2, b2, m3[4]));    // (2,2)
        m3[5] = fmaf(wu, ra, m3[5]);
        m3[6] = fmaf(wv, rb, m3[6]);
        m3[7] = fmaf(a2, ra, fmaf(b2, rb, m3[7]));
    }
    warp_allreduce16<float>(m3, scratch, lane);
    double B[3][3], rh3[3], t[3];
    B[0][0] = (double)m3[0];
    B[0][1] = B[1][0] = 0.0;
    B[0][2] = B[2][0] = (double)m3[1];
    B[1][1] = (double)m3[2];
    B[1][2] = B[2][1] = (double)m3[3];
    B[2][2] = (double)m3[4];
    rh3[0] = (double)m3[5]; rh3[1] = (double)m3[6]; rh3[2] = (double)m3[7];
    if (!chol_solve_dense<3>(B, rh3, t)) return false;
    x[0] = yaw; x[1] = t[0]; x[2] = t[1]; x[3] = t[2];
    return (fabs(t[0]) + fabs(t[1]) + fabs(t[2])) < 1.7e308;
}

// Out-of-line wrapper used by the warp-per-object kernel (keeps its hot code small); pose -> scratch[kScrPt..].
// gate > 0: trimmed fit around the pose currently in scratch[kScrPt..] (see linear_init_impl).
template <int WMODE, int LAYOUT>
__device__ __noinline__ bool linear_init(const KParams& kp, int obj, const float* __restrict__ slot, int n, int lane,
                                         double* scratch, float gate = 0.f) {
    const int P = kp.n_pts;
    const Camera<float> cam = load_camera<float>(kp, obj);
    double x[4];
    float prior[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) prior[i] = gate > 0.f ? (float)scratch[kScrPt + i] : 0.f;
    __syncwarp();
    const bool ok = linear_init_impl<WMODE, LAYOUT>(slot, slot + 3 * P, slot + 5 * P, P, TeamRows(n, 0, 0, 0, 1), lane,
                                                    cam, reinterpret_cast<float*>(scratch), x, gate > 0.f ? prior : nullptr,
                                                    gate * gate);
    __syncwarp();
    if (lane == 0 && (ok || !(gate > 0.f))) {   // a failed TRIMMED fit leaves the pose it started from in place
#pragma unroll
        for (int i = 0; i < 4; ++i) scratch[kScrPt + i] = ok ? (double)(float)x[i] : 0.0;  // fp32 hand-over; .py:119-125
    }
    __syncwarp();
    return ok;
}

// ------------------------------------------------------------------ reprojection-threshold consensus
// The reference starts LM from cv2.solvePnPRansac(EPNP, 30 iterations, reprojectionError = thr) and, when more than four
// points agree with the returned model, keeps only those: LM, the covariance and the returned mask then see the
// RANSAC survivors of the istd inliers (pnp_uncert_cpu.py:34-51; thr = 0.2 x RoI height,
// uncert_prop_pnp_optimizer.py:86-88).  OpenCV's random sampling is not reproduced; the deterministic counterpart here
// takes the START POSE as the model (the on-device linear initialiser, or the caller's init_pose): a compacted inlier
// whose reprojection error exceeds thr pixels is dropped, provided more than four survive.  The slot is compacted again
// in place (order preserved) and the packed mask in kp.inl_out is narrowed accordingly.
// pose: 4 floats (yaw, t); returns the new inlier count (== n if nothing changes).
template <int WMODE, int LAYOUT>
__device__ __noinline__ int consensus_prune(const KParams& kp, int obj, float* slot, int n, int lane, const float* pose, float thr) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    const int P = kp.n_pts;
    float* s3 = slot;
    float* s2 = slot + 3 * P;
    float* sw = slot + 5 * P;
    const Camera<float> cam = load_camera<float>(kp, obj);
    float sn, cs;
    sincosf(pose[0], &sn, &cs);
    const float tx = pose[1], ty = pose[2], tz = pose[3], thr2 = thr * thr;
    const int rows = (n + 31) >> 5;   // <= 32 because n <= MRPNP_MAX_POINTS
    uint32_t keep_word = 0u;          // lane r: survivors of compacted row r
    int count = 0;
    for (int r = 0; r < rows; ++r) {
        const int j = r * 32 + lane;
        const bool live = j < n;
        const int jc = live ? j : 0;
        const float X = s3[sidx<LAYOUT, 3>(jc, 0, P)], Y = s3[sidx<LAYOUT, 3>(jc, 1, P)], Z = s3[sidx<LAYOUT, 3>(jc, 2, P)];
        const float zc = fmaf(cs, Z, fmaf(-sn, X, tz));
        const float iz = 1.f / fmaxf(zc, cam.z_min);   // pnp_uncert_cpu.cpp:36 projection rule
        const float du = fmaf(cam.fx, fmaf(cs, X, fmaf(sn, Z, tx)) * iz, cam.cx) - s2[sidx<LAYOUT, 2>(jc, 0, P)];
        const float dv = fmaf(cam.fy, (Y + ty) * iz, cam.cy) - s2[sidx<LAYOUT, 2>(jc, 1, P)];
        const bool keep = live && (fmaf(du, du, dv * dv) <= thr2);   // false for NaN
        const unsigned m = __ballot_sync(kFull, keep);
        if (lane == r) keep_word = m;
        count += __popc(m);
    }
    if (count <= 4 || count == n) return n;   // pnp_uncert_cpu.py:43: the refinement applies only if > 4 survive
    int base = 0;
    for (int r = 0; r < rows; ++r) {
        const int j = r * 32 + lane;
        const int jc = j < n ? j : 0;
        float v3[3], v2[2], w[WC];
#pragma unroll
        for (int c = 0; c < 3; ++c) v3[c] = s3[sidx<LAYOUT, 3>(jc, c, P)];
#pragma unroll
        for (int c = 0; c < 2; ++c) v2[c] = s2[sidx<LAYOUT, 2>(jc, c, P)];
#pragma unroll
        for (int c = 0; c < WC; ++c) w[c] = sw[sidx<LAYOUT, WC>(jc, c, P)];
        const unsigned m = __shfl_sync(kFull, keep_word, r);
        __syncwarp();   // the reads of this row happen before any lane's compacted writes
        if ((m >> lane) & 1u) {
            const int d = base + __popc(m & ((1u << lane) - 1u));
#pragma unroll
            for (int c = 0; c < 3; ++c) s3[sidx<LAYOUT, 3>(d, c, P)] = v3[c];
#pragma unroll
            for (int c = 0; c < 2; ++c) s2[sidx<LAYOUT, 2>(d, c, P)] = v2[c];
#pragma unroll
            for (int c = 0; c < WC; ++c) sw[sidx<LAYOUT, WC>(d, c, P)] = w[c];
        }
        base += __popc(m);
    }
    __syncwarp();
    if (kp.inl_out) {
        // narrow the packed mask: the i-th set bit of the old mask keeps its bit iff compacted point i survived
        const int rows_p = (P + 31) >> 5;
        uint32_t* words = kp.inl_out + (size_t)obj * rows_p;
        const uint32_t mine = lane < rows_p ? words[lane] : 0u;   // written by this very lane during compaction
        uint32_t out = 0u;
        int off = 0;
        for (int k = 0; k < rows_p; ++k) {
            const uint32_t w = __shfl_sync(kFull, mine, k);
            const uint32_t lo = __shfl_sync(kFull, keep_word, (off >> 5) & 31), hi = __shfl_sync(kFull, keep_word, ((off >> 5) + 1) & 31);
            const uint32_t bits = __funnelshift_r(lo, hi, off & 31);   // survivors of compacted points off .. off + 31
            const int rank = __popc(w & ((1u << lane) - 1u));
            const unsigned nw = __ballot_sync(kFull, ((w >> lane) & 1u) && ((bits >> rank) & 1u));
            if (lane == k) out = nw;
            off += __popc(w);
        }
        if (lane < rows_p) words[lane] = out;
    }
    return count;
}

}  // namespace mrpnp
