// pnp_capi.cu -- C ABI (include/monorun_pnp.h) over the sm_100a solver kernels.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include <new>

#include "monorun_pnp.h"
#include "pnp_kernel.cuh"
#include "pnp_kernel_fast.cuh"
#include "pnp_score.cuh"
#include "pnp_nms.cuh"
#include "pnp_noc.cuh"
#include "pnp_exact_hessian.cuh"
#include "pnp_6dof.cuh"
#include "pnp_6dof_fast.cuh"
#include <stdlib.h>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}

#define MR_CUDA(call)                                                                    \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) return fail(MRPNP_ERR_CUDA, #call ": %s", cudaGetErrorString(e_)); \
    } while (0)

constexpr int kHostStreams = 2;
constexpr int kCounterSets = 8;
constexpr int kCounterInts = 8;  // per set: the five counters of pnp_kernel_fast.cuh (the exact kernels use the first two)

}  // namespace

struct mrpnp_ctx {
    int device = 0;
    int num_sms = 0;
    int max_smem_optin = 0;
    // Work-counter sets (and hand-back lists) rotate between launches so that solves on different streams can overlap;
    // a set is re-used only after the launch that used it last has finished: the next user's stream waits on that
    // launch's event (no host block).  `mu` serialises the context's bookkeeping between host threads.
    std::mutex mu;
    int* counters = nullptr;    // kCounterSets x kCounterInts ints, all zero between launches
    unsigned long long* stats = nullptr;   // device: [0] objects handed back to the exact routine since creation
    int* redo_lists = nullptr;  // kCounterSets x redo_cap entries (MRPNP_PREC_FAST hand-back lists, all zero between launches)
    int redo_cap = 0;
    int next_counter = 0;
    cudaEvent_t set_done[kCounterSets] = {};
    bool set_used[kCounterSets] = {};
    int64_t launches = 0;
    // host-path staging
    cudaStream_t streams[kHostStreams] = {nullptr, nullptr};
    void* chunk_buf[kHostStreams] = {nullptr, nullptr};
    size_t chunk_bytes = 0;
    float* small_buf = nullptr;  // cam / range
    size_t small_bytes = 0;
};

namespace {

using mrpnp::KParams;

struct LaunchPlan {
    int warps, groups, ctas, smem, use_tma, slot_floats;
};

int plan_launch(const mrpnp_ctx* ctx, const mrpnp_params* p, int precision, const void* c3d, const void* c2d,
                const void* wgt, LaunchPlan* plan) {
    const int wc = p->weight_mode == MRPNP_W_FULL ? 3 : 2;
    const size_t slot_bytes = ((size_t)(5 + wc) * p->n_pts * sizeof(float) + 15) & ~size_t(15);
    const size_t header = precision == MRPNP_PREC_FAST ? mrpnp::kFastHeaderBytes : mrpnp::kWarpHeaderBytes;
    const int max_warps = precision == MRPNP_PREC_FAST ? mrpnp::kFastMaxWarps : mrpnp::kMaxWarpsPerCta;
    int groups = (int)std::min<size_t>(max_warps, (size_t)ctx->max_smem_optin / (slot_bytes + header));
    if (groups < 1) return fail(MRPNP_ERR_ARG, "n_pts too large for shared memory%s");
    // keep every SM busy before stacking objects on one SM: at small N spread objects over CTAs
    const int per_sm = (p->n_obj + ctx->num_sms - 1) / ctx->num_sms;
    groups = std::max(1, std::min(groups, per_sm));
    plan->warps = groups;
    plan->groups = groups;
    plan->ctas = std::min(ctx->num_sms, (p->n_obj + groups - 1) / groups);
    plan->smem = (int)(groups * (slot_bytes + header));
    plan->slot_floats = (int)(slot_bytes / sizeof(float));
    const bool aligned = (p->n_pts % 4 == 0) && (((uintptr_t)c3d | (uintptr_t)c2d | (uintptr_t)wgt) % 16 == 0);
    // the fast kernel keeps a planar slot: [N,P,C] tensors are transposed by plain loads instead of bulk copies
    plan->use_tma = (aligned && !(precision == MRPNP_PREC_FAST && p->layout == MRPNP_LAYOUT_INTERLEAVED)) ? 1 : 0;
    return MRPNP_OK;
}

// MRPNP_PREC_FAST needs compacted inliers (the observations are overwritten by tracked residuals) and an even number
// of points per object (two points per 64-bit shared-memory access); other problems run as MRPNP_PREC_MIXED.
int effective_precision(const mrpnp_params* p) {
    if (p->precision == MRPNP_PREC_FAST && (!p->inlier_opt_only || (p->n_pts & 1))) return MRPNP_PREC_MIXED;
    return p->precision;
}

int check_params(const mrpnp_params* p) {
    if (!p) return fail(MRPNP_ERR_ARG, "params is NULL%s");
    if (p->n_obj < 0) return fail(MRPNP_ERR_ARG, "n_obj < 0%s");
    if (p->n_pts < 4 || p->n_pts > MRPNP_MAX_POINTS) return fail(MRPNP_ERR_ARG, "n_pts outside [4, 1024]%s");
    if (p->layout != MRPNP_LAYOUT_PLANAR && p->layout != MRPNP_LAYOUT_INTERLEAVED)
        return fail(MRPNP_ERR_ARG, "bad layout%s");
    if (p->weight_mode < 0 || p->weight_mode > 2) return fail(MRPNP_ERR_ARG, "bad weight_mode%s");
    if (p->precision < MRPNP_PREC_FP64 || p->precision > MRPNP_PREC_FAST) return fail(MRPNP_ERR_ARG, "bad precision%s");
    if (p->cov_mode < 0 || p->cov_mode > 2) return fail(MRPNP_ERR_ARG, "bad cov_mode%s");
    if (p->init_mode != MRPNP_INIT_GIVEN && p->init_mode != MRPNP_INIT_LINEAR) return fail(MRPNP_ERR_ARG, "bad init_mode%s");
    if (p->cam_stride != 0 && p->cam_stride != 9) return fail(MRPNP_ERR_ARG, "cam_stride must be 0 or 9%s");
    if (p->range_stride != 0 && p->range_stride != 4) return fail(MRPNP_ERR_ARG, "range_stride must be 0 or 4%s");
    if (p->weight_mode == MRPNP_W_LOGSTD && !(p->std_scale > 0.f)) return fail(MRPNP_ERR_ARG, "std_scale must be > 0%s");
    return MRPNP_OK;
}

template <int WMODE, int LAYOUT>
cudaError_t launch_exact(int precision, const KParams& kp, const LaunchPlan& plan, cudaStream_t stream) {
    cudaError_t e;
    if (precision == MRPNP_PREC_FP64) {
        auto k = mrpnp::pnp_lm_kernel<false, WMODE, LAYOUT>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem);
        if (e != cudaSuccess) return e;
        k<<<plan.ctas, plan.warps * 32, plan.smem, stream>>>(kp);
    } else {
        auto k = mrpnp::pnp_lm_kernel<true, WMODE, LAYOUT>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem);
        if (e != cudaSuccess) return e;
        k<<<plan.ctas, plan.warps * 32, plan.smem, stream>>>(kp);
    }
    return cudaGetLastError();
}

template <int WMODE>
cudaError_t launch_fast(const KParams& kp, const LaunchPlan& plan, cudaStream_t stream) {
    void (*k)(const KParams) = kp.n_pts == 784 ? mrpnp::pnp_lm_fast_kernel<WMODE, 784> : mrpnp::pnp_lm_fast_kernel<WMODE, 0>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem);
    if (e != cudaSuccess) return e;
    k<<<plan.ctas, plan.warps * 32, plan.smem, stream>>>(kp);
    return cudaGetLastError();
}

cudaError_t dispatch(const mrpnp_params* p, int precision, const KParams& kp, const LaunchPlan& plan, cudaStream_t stream) {
    if (precision == MRPNP_PREC_FAST) {
        if (p->weight_mode == MRPNP_W_LOGSTD) return launch_fast<MRPNP_W_LOGSTD>(kp, plan, stream);
        if (p->weight_mode == MRPNP_W_ISTD) return launch_fast<MRPNP_W_ISTD>(kp, plan, stream);
        return launch_fast<MRPNP_W_FULL>(kp, plan, stream);
    }
#define MR_CASE(W, L) \
    if (p->weight_mode == W && p->layout == L) return launch_exact<W, L>(precision, kp, plan, stream);
    MR_CASE(MRPNP_W_LOGSTD, MRPNP_LAYOUT_PLANAR)
    MR_CASE(MRPNP_W_ISTD, MRPNP_LAYOUT_PLANAR)
    MR_CASE(MRPNP_W_FULL, MRPNP_LAYOUT_PLANAR)
    MR_CASE(MRPNP_W_LOGSTD, MRPNP_LAYOUT_INTERLEAVED)
    MR_CASE(MRPNP_W_ISTD, MRPNP_LAYOUT_INTERLEAVED)
    MR_CASE(MRPNP_W_FULL, MRPNP_LAYOUT_INTERLEAVED)
#undef MR_CASE
    return cudaErrorInvalidValue;
}

struct DenseArgs {
    const mrpnp_dense_params* dp;
    const float* dims;
    const float* dims_var;
    const float* distance;
    const int64_t* labels;       // channel selection (NULL for pre-sliced maps)
    const int64_t* dim_labels;   // class of each object for the dimension coder (NULL without dim_means)
};

int solve_device(mrpnp_ctx* ctx, const mrpnp_params* p, const float* c3d, const float* c2d, const float* wgt,
                 const float* cam, const float* range, const float* init, const uint32_t* inl_in, float* result,
                 uint32_t* inl_out, double* result64, cudaStream_t stream, const DenseArgs* dense = nullptr) {
    if (p->n_obj == 0) return MRPNP_OK;
    if (!c3d || !c2d || !wgt || !cam || !range || (!result && p->n_peers == 0)) return fail(MRPNP_ERR_ARG, "NULL tensor pointer%s");
    if (p->n_peers < 0 || p->n_peers > MRPNP_MAX_PEERS) return fail(MRPNP_ERR_ARG, "n_peers outside [0, 8]%s");
    for (int r = 0; r < p->n_peers; ++r)
        if (!p->peer_results[r]) return fail(MRPNP_ERR_ARG, "NULL peer result buffer%s");
    if (p->init_mode == MRPNP_INIT_GIVEN && !init) return fail(MRPNP_ERR_ARG, "init_pose is NULL with MRPNP_INIT_GIVEN%s");
    const int precision = effective_precision(p);
    LaunchPlan plan;
    int rc = plan_launch(ctx, p, precision, c3d, dense ? c3d : c2d, wgt, &plan);
    if (rc != MRPNP_OK) return rc;
    KParams kp;
    kp.c3d = c3d; kp.c2d = c2d; kp.wgt = wgt; kp.cam = cam; kp.range = range; kp.init = init;
    kp.inl_in = inl_in; kp.result = result; kp.inl_out = inl_out; kp.result64 = result64;
    kp.n_peers = p->n_peers; kp.row_offset = p->row_offset;
    for (int r = 0; r < MRPNP_MAX_PEERS; ++r) kp.peer[r] = r < p->n_peers ? p->peer_results[r] : nullptr;
    const bool flags = p->n_peers > 0 && p->peer_flags[0] != nullptr;
    for (int r = 0; r < MRPNP_MAX_PEERS; ++r) kp.peer_flag[r] = (flags && r < p->n_peers) ? p->peer_flags[r] : nullptr;
    if (flags) {
        for (int r = 0; r < p->n_peers; ++r)
            if (!p->peer_flags[r]) return fail(MRPNP_ERR_ARG, "peer_flags[r] is NULL for r < n_peers%s");
        if (p->flag_slot < 0 || p->flag_slot >= MRPNP_MAX_PEERS) return fail(MRPNP_ERR_ARG, "flag_slot outside [0, 8)%s");
    }
    kp.flag_slot = p->flag_slot; kp.flag_value = p->flag_value;
    kp.acks = p->n_peers > 0 ? p->acks : nullptr; kp.ack_value = p->ack_value;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (precision == MRPNP_PREC_FAST && p->n_obj > ctx->redo_cap) {
        // grow the hand-back lists (cudaFree waits for the launches still using them)
        if (ctx->redo_lists) MR_CUDA(cudaFree(ctx->redo_lists));
        ctx->redo_lists = nullptr;
        ctx->redo_cap = 0;
        const int cap = std::max(p->n_obj, 1024);
        MR_CUDA(cudaMalloc(&ctx->redo_lists, sizeof(int) * (size_t)cap * kCounterSets));
        MR_CUDA(cudaMemset(ctx->redo_lists, 0, sizeof(int) * (size_t)cap * kCounterSets));
        ctx->redo_cap = cap;
    }
    // take the next counter set; if the launch that used it last may still be running, this stream waits for it
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &capturing);
    int set = kCounterSets - 1;   // launches captured into a CUDA graph share the last set (replays are stream-ordered)
    if (capturing == cudaStreamCaptureStatusNone) {
        set = ctx->next_counter;
        ctx->next_counter = (ctx->next_counter + 1) % (kCounterSets - 1);
        if (ctx->set_used[set] && cudaEventQuery(ctx->set_done[set]) != cudaSuccess)
            MR_CUDA(cudaStreamWaitEvent(stream, ctx->set_done[set], 0));
        (void)cudaGetLastError();  // cudaEventQuery's cudaErrorNotReady is not an error
    }
    kp.counters = ctx->counters + kCounterInts * set;
    kp.redo_list = precision == MRPNP_PREC_FAST ? ctx->redo_lists + (size_t)ctx->redo_cap * set : nullptr;
    kp.work_list = nullptr; kp.work_count = nullptr;
    kp.stats = ctx->stats;
    kp.band_first = p->band_first; kp.band_rel = p->band_rel; kp.band_mix = p->band_mix;
    kp.band_ratio = p->band_ratio; kp.band_rel_min = p->band_rel_min; kp.band_ratio_from = p->band_ratio_from;
    kp.hand_back_log = p->hand_back_log;
    kp.global_interleaved = (precision == MRPNP_PREC_FAST && p->layout == MRPNP_LAYOUT_INTERLEAVED) ? 1 : 0;
    kp.ransac_thres = (!dense && p->inlier_opt_only) ? p->ransac_thres : nullptr;
    kp.ransac_ratio = (dense && p->inlier_opt_only) ? p->ransac_ratio : 0.f;
    kp.n_obj = p->n_obj; kp.n_pts = p->n_pts; kp.cam_stride = p->cam_stride; kp.range_stride = p->range_stride;
    kp.cov_mode = p->cov_mode; kp.init_mode = p->init_mode; kp.inlier_opt_only = p->inlier_opt_only;
    kp.max_iter = p->max_iterations; kp.adopt_ftol = p->adopt_candidate_on_ftol;
    kp.use_tma = plan.use_tma;
    kp.slot_floats = plan.slot_floats;
    kp.z_min = p->z_min; kp.std_scale = p->std_scale; kp.istd_thres = p->istd_thres;
    kp.dense = dense ? 1 : 0;
    kp.roi_w = dense ? dense->dp->roi_w : 0;
    kp.dims = dense ? dense->dims : nullptr;
    kp.dims_var = dense ? dense->dims_var : nullptr;
    kp.dim_means = dense ? dense->dp->dim_means : nullptr;
    kp.dim_stds = dense ? dense->dp->dim_stds : nullptr;
    kp.dim_labels = dense ? reinterpret_cast<const long long*>(dense->dim_labels) : nullptr;
    kp.n_dim_classes = dense ? dense->dp->n_dim_classes : 0;
    kp.dims_out = dense ? dense->dp->dims_out : nullptr;
    kp.dims_var_out = dense ? dense->dp->dims_var_out : nullptr;
    for (int i = 0; i < 3; ++i) {
        kp.noc_mean[i] = dense ? dense->dp->noc_mean[i] : 0.f;
        kp.noc_std[i] = dense ? dense->dp->noc_std[i] : 1.f;
    }
    kp.distance = dense ? dense->distance : nullptr;
    kp.labels = dense ? reinterpret_cast<const long long*>(dense->labels) : nullptr;
    kp.pred_stride = (dense && dense->dp->num_classes > 0) ? dense->dp->pred_stride : 0;
    if (dense) {
        const float g = dense->dp->focal_gain / dense->dp->scaling_denominator;
        kp.proj_gain2 = g * g;
        kp.inv_scaling_denominator = 1.f / dense->dp->scaling_denominator;
        kp.distance_min = dense->dp->distance_min;
    } else {
        kp.proj_gain2 = 0.f; kp.inv_scaling_denominator = 1.f; kp.distance_min = 0.f;
    }
    kp.prefetch_distance = plan.ctas * plan.groups;
    cudaError_t e = dispatch(p, precision, kp, plan, stream);
    if (e != cudaSuccess) return fail(MRPNP_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e));
    ctx->launches += 1;
    if (capturing == cudaStreamCaptureStatusNone) {
        MR_CUDA(cudaEventRecord(ctx->set_done[set], stream));
        ctx->set_used[set] = true;
    }
    return MRPNP_OK;
}

}  // namespace

extern "C" {

int mrpnp_version(void) { return MRPNP_VERSION; }
const char* mrpnp_last_error(void) { return g_err; }

void mrpnp_default_params(mrpnp_params* p, int32_t n_obj, int32_t n_pts) {
    memset(p, 0, sizeof(*p));
    p->n_obj = n_obj;
    p->n_pts = n_pts;
    p->layout = MRPNP_LAYOUT_PLANAR;
    p->weight_mode = MRPNP_W_LOGSTD;
    p->cam_stride = 0;
    p->range_stride = 0;
    p->precision = MRPNP_PREC_FAST;
    p->cov_mode = MRPNP_COV_PIPELINE;
    p->init_mode = MRPNP_INIT_GIVEN;
    p->inlier_opt_only = 1;      // configs/kitti_multiclass.py:127
    p->max_iterations = 50;      // ceres::Solver::Options default
    p->adopt_candidate_on_ftol = 0;
    p->z_min = 0.5f;             // configs/kitti_multiclass.py:125
    p->std_scale = 10.f;         // uncert_prop_pnp_optimizer.py:28
    p->istd_thres = 0.6f;        // configs/kitti_multiclass.py:126
    p->band_first = 8e-6f;
    p->band_rel = 4e-3f;
    p->band_mix = 2e-6f;
    p->band_ratio = 2e-5f;      // slowly converging objects (>= 10 evaluations): band scaled by 2e-5 |previous step| / |step|,
    p->band_rel_min = 1e-4f;    // not below 1e-4 / band_rel of its width (profiles/r02_band_sweep.txt)
    p->band_ratio_from = 10;
}

int mrpnp_create(mrpnp_ctx** out, int device) {
    if (!out) return fail(MRPNP_ERR_ARG, "ctx out pointer is NULL%s");
    *out = nullptr;
    MR_CUDA(cudaSetDevice(device));
    mrpnp_ctx* c = new (std::nothrow) mrpnp_ctx();
    if (!c) return fail(MRPNP_ERR_ARG, "out of host memory%s");
    c->device = device;
    cudaDeviceProp prop;
    MR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        delete c;
        return fail(MRPNP_ERR_CUDA, "libmonorun_pnp is built for sm_100a only; device is %s", prop.name);
    }
    c->num_sms = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    MR_CUDA(cudaMalloc(&c->counters, sizeof(int) * kCounterInts * kCounterSets));
    MR_CUDA(cudaMemset(c->counters, 0, sizeof(int) * kCounterInts * kCounterSets));
    MR_CUDA(cudaMalloc(&c->stats, sizeof(unsigned long long) * 4));
    MR_CUDA(cudaMemset(c->stats, 0, sizeof(unsigned long long) * 4));
    for (int i = 0; i < kCounterSets; ++i) MR_CUDA(cudaEventCreateWithFlags(&c->set_done[i], cudaEventDisableTiming));
    *out = c;
    return MRPNP_OK;
}

void mrpnp_destroy(mrpnp_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < kHostStreams; ++i) {
        if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
        if (c->chunk_buf[i]) cudaFree(c->chunk_buf[i]);
    }
    for (int i = 0; i < kCounterSets; ++i)
        if (c->set_done[i]) cudaEventDestroy(c->set_done[i]);
    if (c->small_buf) cudaFree(c->small_buf);
    if (c->counters) cudaFree(c->counters);
    if (c->stats) cudaFree(c->stats);
    if (c->redo_lists) cudaFree(c->redo_lists);
    delete c;
}

int64_t mrpnp_launch_count(const mrpnp_ctx* c) { return c ? c->launches : 0; }

int mrpnp_gather_wait(mrpnp_ctx* ctx, const uint32_t* flags, int32_t n, uint32_t value, uint32_t* const* peer_acks,
                      int32_t ack_slot, uint32_t ack_value, void* stream) {
    if (!ctx) return fail(MRPNP_ERR_ARG, "ctx is NULL%s");
    if (n < 1 || n > MRPNP_MAX_PEERS) return fail(MRPNP_ERR_ARG, "n outside [1, 8]%s");
    if (!flags && !peer_acks) return fail(MRPNP_ERR_ARG, "neither flags nor peer_acks given%s");
    if (peer_acks && (ack_slot < 0 || ack_slot >= MRPNP_MAX_PEERS)) return fail(MRPNP_ERR_ARG, "ack_slot outside [0, 8)%s");
    MR_CUDA(cudaSetDevice(ctx->device));
    g_err[0] = 0;
    mrpnp::KParams kp{};   // carries the peer pointers by value
    if (peer_acks)
        for (int r = 0; r < n; ++r) {
            if (!peer_acks[r]) return fail(MRPNP_ERR_ARG, "peer_acks[r] is NULL for r < n%s");
            kp.peer_flag[r] = peer_acks[r];
        }
    mrpnp::gather_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(flags, n, value, kp, peer_acks ? n : 0, ack_slot,
                                                                             ack_value, ctx->stats);
    MR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return MRPNP_OK;
}

int64_t mrpnp_gather_timeouts(mrpnp_ctx* c) {
    if (!c || !c->stats) return 0;
    unsigned long long v = 0;
    cudaSetDevice(c->device);
    if (cudaMemcpy(&v, c->stats + 1, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;   // synchronises the device
    return (int64_t)v;
}

int64_t mrpnp_handed_back_count(mrpnp_ctx* c) {
    if (!c || !c->stats) return 0;
    unsigned long long v = 0;
    cudaSetDevice(c->device);
    if (cudaMemcpy(&v, c->stats, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;   // synchronises the device
    return (int64_t)v;
}

int mrpnp_pose_features(mrpnp_ctx* ctx, const float* rows, const float* dims, const float* cov_calib_logscale,
                        float cov_correction_sd, int32_t distance_z_depth, int32_t use_calib,
                        const float* norm_mean, const float* norm_var, const float* norm_weight, const float* norm_bias,
                        float norm_eps, float* feat, float* cov_calib, int32_t n, void* stream) {
    if (!ctx) return fail(MRPNP_ERR_ARG, "ctx is NULL%s");
    if (n < 0) return fail(MRPNP_ERR_ARG, "n < 0%s");
    if (n == 0) return MRPNP_OK;
    if (!rows || !dims || !feat) return fail(MRPNP_ERR_ARG, "NULL tensor pointer%s");
    const bool any_norm = norm_mean || norm_var || norm_weight || norm_bias;
    if (any_norm && !(norm_mean && norm_var && norm_weight && norm_bias))
        return fail(MRPNP_ERR_ARG, "pose_norm needs mean, var, weight and bias%s");
    MR_CUDA(cudaSetDevice(ctx->device));
    g_err[0] = 0;
    mrpnp::ScoreParams sp;
    sp.rows = rows; sp.dims = dims;
    sp.calib_logscale = cov_calib_logscale;
    sp.corr_sd = cov_correction_sd; sp.corr_z_depth = distance_z_depth; sp.use_calib = use_calib;
    sp.norm_mean = norm_mean; sp.norm_var = norm_var; sp.norm_weight = norm_weight; sp.norm_bias = norm_bias;
    sp.norm_eps = norm_eps; sp.feat = feat; sp.cov_calib = cov_calib; sp.n = n;
    mrpnp::pose_features_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(sp);
    MR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return MRPNP_OK;
}

int mrpnp_score_stage(mrpnp_ctx* ctx, const float* rows, const float* dims, const float* cov_calib_logscale,
                      float cov_correction_sd, int32_t distance_z_depth, int32_t use_calib,
                      const float* norm_mean, const float* norm_var, const float* norm_weight, const float* norm_bias,
                      float norm_eps, const float* reg_fc_out, const float* w1, const float* b1, const float* w2t,
                      const float* b2, const float* w3, const float* b3, int32_t h1, int32_t h2,
                      const float* det_scores, int32_t pre_sigmoid, float* scores, float* bbox_3d, float* cov_calib,
                      float* logits, int32_t n, void* stream) {
    if (!ctx) return fail(MRPNP_ERR_ARG, "ctx is NULL%s");
    if (n < 0) return fail(MRPNP_ERR_ARG, "n < 0%s");
    if (n == 0) return MRPNP_OK;
    if (!rows || !dims || !w1 || !b1 || !w2t || !b2 || !w3 || !b3 || (!scores && !bbox_3d && !logits))
        return fail(MRPNP_ERR_ARG, "NULL tensor pointer%s");
    if (h1 < 1 || h2 < 1 || h1 > 4096 || h2 > 4096) return fail(MRPNP_ERR_ARG, "layer widths outside [1, 4096]%s");
    const bool any_norm = norm_mean || norm_var || norm_weight || norm_bias;
    if (any_norm && !(norm_mean && norm_var && norm_weight && norm_bias))
        return fail(MRPNP_ERR_ARG, "pose_norm needs mean, var, weight and bias%s");
    MR_CUDA(cudaSetDevice(ctx->device));
    g_err[0] = 0;
    mrpnp::ScoreParams sp;
    sp.rows = rows; sp.dims = dims; sp.calib_logscale = cov_calib_logscale;
    sp.corr_sd = cov_correction_sd; sp.corr_z_depth = distance_z_depth; sp.use_calib = use_calib;
    sp.norm_mean = norm_mean; sp.norm_var = norm_var; sp.norm_weight = norm_weight; sp.norm_bias = norm_bias;
    sp.norm_eps = norm_eps; sp.feat = nullptr; sp.cov_calib = cov_calib; sp.n = n;
    mrpnp::MlpParams mp;
    mp.reg_fc_out = reg_fc_out; mp.w1 = w1; mp.b1 = b1; mp.w2t = w2t; mp.b2 = b2; mp.w3 = w3; mp.b3 = b3;
    mp.h1 = h1; mp.h2 = h2; mp.det_scores = det_scores; mp.pre_sigmoid = pre_sigmoid;
    mp.logits = logits; mp.scores = scores; mp.bbox3d = bbox_3d;
    const size_t smem = sizeof(float) * ((size_t)mrpnp::kScoreTile * (20 + h1 + mrpnp::kScoreThreads / 32));
    if (smem > (size_t)ctx->max_smem_optin) return fail(MRPNP_ERR_ARG, "pose layer too wide for shared memory%s");
    MR_CUDA(cudaFuncSetAttribute(mrpnp::score_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ctas = (n + mrpnp::kScoreTile - 1) / mrpnp::kScoreTile;
    mrpnp::score_stage_kernel<<<ctas, mrpnp::kScoreThreads, smem, static_cast<cudaStream_t>(stream)>>>(sp, mp);
    MR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return MRPNP_OK;
}

int mrpnp_nms_bev(mrpnp_ctx* ctx, const float* bbox_3d, const int64_t* labels, const int32_t* group_offsets,
                  int32_t n_groups, int32_t max_group, float iou_thr, uint8_t* keep, void* stream) {
    if (!ctx) return fail(MRPNP_ERR_ARG, "ctx is NULL%s");
    if (n_groups < 0 || max_group < 0) return fail(MRPNP_ERR_ARG, "negative size%s");
    if (n_groups == 0 || max_group == 0) return MRPNP_OK;
    if (!bbox_3d || !group_offsets || !keep) return fail(MRPNP_ERR_ARG, "NULL tensor pointer%s");
    if (max_group > 4096) return fail(MRPNP_ERR_ARG, "more than 4096 objects in one image%s");
    int cap = 1;
    while (cap < max_group) cap <<= 1;
    const size_t smem = (size_t)cap * (4 * sizeof(int) + sizeof(mrpnp::BevBox));
    MR_CUDA(cudaSetDevice(ctx->device));
    g_err[0] = 0;
    MR_CUDA(cudaFuncSetAttribute(mrpnp::nms_bev_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mrpnp::nms_bev_kernel<<<n_groups, mrpnp::kNmsThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        bbox_3d, reinterpret_cast<const long long*>(labels), group_offsets, iou_thr, cap, keep);
    MR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return MRPNP_OK;
}

int mrpnp_exact_hessian(mrpnp_ctx* ctx, const mrpnp_params* p,
                        const float* coords_3d, const float* coords_2d, const float* weights,
                        const float* cam_mats, const float* uv_range,
                        const float* pose, int32_t pose_stride, const uint32_t* inlier_in,
                        float* hessian, float* rows, void* stream) {
    if (!ctx || !p) return fail(MRPNP_ERR_ARG, "NULL argument%s");
    if (p->n_obj < 0) return fail(MRPNP_ERR_ARG, "n_obj < 0%s");
    if (p->n_pts < 1 || p->n_pts > MRPNP_MAX_POINTS) return fail(MRPNP_ERR_ARG, "n_pts out of range%s");
    if (p->layout != MRPNP_LAYOUT_PLANAR && p->layout != MRPNP_LAYOUT_INTERLEAVED) return fail(MRPNP_ERR_ARG, "bad layout%s");
    if (p->weight_mode != MRPNP_W_LOGSTD && p->weight_mode != MRPNP_W_ISTD)
        return fail(MRPNP_ERR_ARG, "the exact Hessian is defined for per-axis weights only (MRPNP_W_LOGSTD / MRPNP_W_ISTD)%s");
    if ((p->cam_stride != 0 && p->cam_stride != 9) || (p->range_stride != 0 && p->range_stride != 4))
        return fail(MRPNP_ERR_ARG, "bad cam_stride / range_stride%s");
    if (pose_stride < 4) return fail(MRPNP_ERR_ARG, "pose_stride < 4%s");
    if (p->n_obj == 0) return MRPNP_OK;
    if (!coords_3d || !coords_2d || !weights || !cam_mats || !uv_range || !pose || (!hessian && !rows))
        return fail(MRPNP_ERR_ARG, "NULL tensor pointer%s");
    MR_CUDA(cudaSetDevice(ctx->device));
    g_err[0] = 0;
    mrxh::KParams kp;
    kp.coords_3d = coords_3d; kp.coords_2d = coords_2d; kp.weights = weights; kp.cam_mats = cam_mats;
    kp.uv_range = uv_range; kp.pose = pose; kp.inlier = inlier_in; kp.hessian = hessian; kp.rows = rows;
    kp.n_obj = p->n_obj; kp.n_pts = p->n_pts; kp.planar = p->layout == MRPNP_LAYOUT_PLANAR;
    kp.logstd = p->weight_mode == MRPNP_W_LOGSTD; kp.cam_stride = p->cam_stride; kp.range_stride = p->range_stride;
    kp.pose_stride = pose_stride; kp.z_min = p->z_min; kp.std_scale = p->std_scale;
    const int ctas = std::min((p->n_obj + mrxh::kWarpsPerCta - 1) / mrxh::kWarpsPerCta, ctx->num_sms * 16);
    mrxh::exact_hessian_kernel<<<ctas, mrxh::kWarpsPerCta * 32, 0, static_cast<cudaStream_t>(stream)>>>(kp);
    MR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return MRPNP_OK;
}

int mrpnp_solve_6dof(mrpnp_ctx* ctx, const mrpnp_params* p,
                     const float* coords_3d, const float* coords_2d, const float* weights,
                     const float* cam_mats, const float* uv_range, const float* init_pose6,
                     const uint32_t* inlier_in, double* result, void* stream) {
    if (!ctx || !p) return fail(MRPNP_ERR_ARG, "NULL argument%s");
    if (p->n_obj < 0) return fail(MRPNP_ERR_ARG, "n_obj < 0%s");
    if (p->n_pts < 1 || p->n_pts > MRPNP_MAX_POINTS) return fail(MRPNP_ERR_ARG, "n_pts out of range%s");
    if (p->layout != MRPNP_LAYOUT_PLANAR && p->layout != MRPNP_LAYOUT_INTERLEAVED) return fail(MRPNP_ERR_ARG, "bad layout%s");
    if (p->weight_mode < MRPNP_W_LOGSTD || p->weight_mode > MRPNP_W_FULL) return fail(MRPNP_ERR_ARG, "bad weight_mode%s");
    if ((p->cam_stride != 0 && p->cam_stride != 9) || (p->range_stride != 0 && p->range_stride != 4))
        return fail(MRPNP_ERR_ARG, "bad cam_stride / range_stride%s");
    if (p->max_iterations < 0) return fail(MRPNP_ERR_ARG, "max_iterations < 0%s");
    if (p->n_obj == 0) return MRPNP_OK;
    if (!coords_3d || !coords_2d || !weights || !cam_mats || !uv_range || !init_pose6 || !result)
        return fail(MRPNP_ERR_ARG, "NULL tensor pointer%s");
    MR_CUDA(cudaSetDevice(ctx->device));
    g_err[0] = 0;
    mr6::KParams kp;
    kp.coords_3d = coords_3d; kp.coords_2d = coords_2d; kp.weights = weights; kp.cam_mats = cam_mats;
    kp.uv_range = uv_range; kp.init = init_pose6; kp.inlier = inlier_in; kp.result = result;
    kp.n_obj = p->n_obj; kp.n_pts = p->n_pts; kp.planar = p->layout == MRPNP_LAYOUT_PLANAR;
    kp.wmode = p->weight_mode; kp.cam_stride = p->cam_stride; kp.range_stride = p->range_stride;
    kp.max_iterations = p->max_iterations; kp.z_min = p->z_min; kp.std_scale = p->std_scale;
    const bool full = p->weight_mode == MRPNP_W_FULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p->precision != MRPNP_PREC_FP64) {
        // MRPNP_PREC_MIXED / _FAST: pnp_6dof_fast.cuh -- inliers staged once in shared memory, fp64 cost chain, fp32 normal
        // equations, one CTA per SM, objects handed out by a counter (the sets of the 4-DoF launches, same rotation)
        const int slot = mr6::mixed_slot_bytes(p->n_pts, p->weight_mode);
        const int warps = std::min(mr6::kMixMaxWarps, (ctx->max_smem_optin - 1024) / slot);
        if (warps < 1) return fail(MRPNP_ERR_ARG, "n_pts too large for the shared-memory slot%s");
        const int smem = warps * slot;
        void (*kernel)(const mr6::KParams, int*, int) =
            p->weight_mode == MRPNP_W_FULL ? mr6::pnp_6dof_mixed_kernel<mr6::kWFull>
            : p->weight_mode == MRPNP_W_ISTD ? mr6::pnp_6dof_mixed_kernel<mr6::kWIstd> : mr6::pnp_6dof_mixed_kernel<mr6::kWLogstd>;
        MR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        std::lock_guard<std::mutex> lock(ctx->mu);
        cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &capturing);
        int set = kCounterSets - 1;
        if (capturing == cudaStreamCaptureStatusNone) {
            set = ctx->next_counter;
            ctx->next_counter = (ctx->next_counter + 1) % (kCounterSets - 1);
            if (ctx->set_used[set] && cudaEventQuery(ctx->set_done[set]) != cudaSuccess)
                MR_CUDA(cudaStreamWaitEvent(st, ctx->set_done[set], 0));
            (void)cudaGetLastError();
        }
        int* counters = ctx->counters + kCounterInts * set;
        const int ctas = std::min(ctx->num_sms, (p->n_obj + warps - 1) / warps);
        kernel<<<ctas, warps * 32, smem, st>>>(kp, counters, slot);
        MR_CUDA(cudaGetLastError());
        ctx->launches += 1;
        if (capturing == cudaStreamCaptureStatusNone) {
            MR_CUDA(cudaEventRecord(ctx->set_done[set], st));
            ctx->set_used[set] = true;
        }
        return MRPNP_OK;
    }
    // MRPNP_PREC_FP64: one object per warp, CTAs handed out by the hardware scheduler: objects differ in LM iterations,
    // and a persistent grid with static striding measured 10 % slower (profiles/r01b_ncu_noc_summary.txt)
    const int ctas = (p->n_obj + mr6::kWarpsPerCta - 1) / mr6::kWarpsPerCta;
    if (full)
        mr6::pnp_6dof_kernel<true><<<ctas, mr6::kWarpsPerCta * 32, 0, st>>>(kp);
    else
        mr6::pnp_6dof_kernel<false><<<ctas, mr6::kWarpsPerCta * 32, 0, st>>>(kp);
    MR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return MRPNP_OK;
}

int mrpnp_solve_noc(mrpnp_ctx* ctx, const mrpnp_noc_params* p,
                    const float* coords_3d, const float* coords_2d, const float* weights,
                    const float* logdim, const float* logdim_wgt,
                    const float* cam_mats, const float* uv_range, const float* init_dimpose,
                    const uint32_t* inlier_in, double* result, void* stream) {
    if (!ctx || !p) return fail(MRPNP_ERR_ARG, "NULL argument%s");
    if (p->n_obj < 0) return fail(MRPNP_ERR_ARG, "n_obj < 0%s");
    if (p->n_pts < 1 || p->n_pts > MRPNP_MAX_POINTS) return fail(MRPNP_ERR_ARG, "n_pts out of range%s");
    if (p->layout != MRPNP_LAYOUT_PLANAR && p->layout != MRPNP_LAYOUT_INTERLEAVED) return fail(MRPNP_ERR_ARG, "bad layout%s");
    if (p->weight_mode != MRPNP_W_ISTD && p->weight_mode != MRPNP_W_FULL)
        return fail(MRPNP_ERR_ARG, "weight_mode must be MRPNP_W_ISTD or MRPNP_W_FULL%s");
    if ((p->cam_stride != 0 && p->cam_stride != 9) || (p->range_stride != 0 && p->range_stride != 4))
        return fail(MRPNP_ERR_ARG, "bad cam_stride / range_stride%s");
    if (!(p->huber_delta > 0.f)) return fail(MRPNP_ERR_ARG, "huber_delta must be positive%s");
    if (p->max_iterations < 0) return fail(MRPNP_ERR_ARG, "max_iterations < 0%s");
    if (p->n_obj == 0) return MRPNP_OK;
    if (!coords_3d || !coords_2d || !weights || !logdim || !logdim_wgt || !cam_mats || !uv_range || !init_dimpose || !result)
        return fail(MRPNP_ERR_ARG, "NULL tensor pointer%s");
    MR_CUDA(cudaSetDevice(ctx->device));
    g_err[0] = 0;
    mrnoc::KParams kp;
    kp.coords_3d = coords_3d; kp.coords_2d = coords_2d; kp.weights = weights;
    kp.logdim = logdim; kp.logdim_wgt = logdim_wgt; kp.cam_mats = cam_mats; kp.uv_range = uv_range;
    kp.init = init_dimpose; kp.inlier = inlier_in; kp.result = result;
    kp.n_obj = p->n_obj; kp.n_pts = p->n_pts; kp.planar = p->layout == MRPNP_LAYOUT_PLANAR;
    kp.cam_stride = p->cam_stride; kp.range_stride = p->range_stride; kp.max_iterations = p->max_iterations;
    kp.z_min = p->z_min; kp.delta = p->huber_delta;
    // one object per warp, CTAs handed out by the hardware scheduler (see mrpnp_solve_6dof)
    const int ctas = (p->n_obj + mrnoc::kWarpsPerCta - 1) / mrnoc::kWarpsPerCta;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p->weight_mode == MRPNP_W_FULL)
        mrnoc::pnp_noc_kernel<true><<<ctas, mrnoc::kWarpsPerCta * 32, 0, st>>>(kp);
    else
        mrnoc::pnp_noc_kernel<false><<<ctas, mrnoc::kWarpsPerCta * 32, 0, st>>>(kp);
    MR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return MRPNP_OK;
}

int mrpnp_finish_scores(mrpnp_ctx* ctx, const float* score_logits, const float* rows, const float* dims,
                        const float* det_scores, int32_t pre_sigmoid, float* scores, float* bbox_3d, int32_t n,
                        void* stream) {
    if (!ctx) return fail(MRPNP_ERR_ARG, "ctx is NULL%s");
    if (n < 0) return fail(MRPNP_ERR_ARG, "n < 0%s");
    if (n == 0) return MRPNP_OK;
    if (!score_logits || !rows || (bbox_3d && !dims) || (!scores && !bbox_3d)) return fail(MRPNP_ERR_ARG, "NULL tensor pointer%s");
    MR_CUDA(cudaSetDevice(ctx->device));
    g_err[0] = 0;
    mrpnp::finish_scores_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        score_logits, rows, dims, det_scores, pre_sigmoid, scores, bbox_3d, n);
    MR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return MRPNP_OK;
}

int mrpnp_kernel_info(mrpnp_ctx* ctx, const mrpnp_params* p, int32_t info[4]) {
    if (!ctx || !info) return fail(MRPNP_ERR_ARG, "NULL argument%s");
    int rc = check_params(p);
    if (rc != MRPNP_OK) return rc;
    LaunchPlan plan;
    const int precision = effective_precision(p);
    rc = plan_launch(ctx, p, precision, nullptr, nullptr, nullptr, &plan);
    if (rc != MRPNP_OK) return rc;
    info[0] = plan.warps; info[1] = plan.ctas; info[2] = plan.smem; info[3] = plan.use_tma;
    return MRPNP_OK;
}

int mrpnp_solve(mrpnp_ctx* ctx, const mrpnp_params* p, const float* coords_3d, const float* coords_2d,
                const float* weights, const float* cam_mats, const float* uv_range, const float* init_pose,
                const uint32_t* inlier_in, float* result, uint32_t* inlier_out, double* result64, void* stream) {
    if (!ctx) return fail(MRPNP_ERR_ARG, "ctx is NULL%s");
    int rc = check_params(p);
    if (rc != MRPNP_OK) return rc;
    MR_CUDA(cudaSetDevice(ctx->device));
    g_err[0] = 0;
    return solve_device(ctx, p, coords_3d, coords_2d, weights, cam_mats, uv_range, init_pose, inlier_in, result,
                        inlier_out, result64, static_cast<cudaStream_t>(stream));
}

int mrpnp_solve_dense(mrpnp_ctx* ctx, const mrpnp_params* p, const mrpnp_dense_params* dp, const float* noc_pred,
                      const float* proj_logstd, const float* rois, const int64_t* labels, const float* dims,
                      const float* dims_var, const float* distance, const float* cam_mats, const float* uv_range, const float* init_pose, float* result,
                      uint32_t* inlier_out, void* stream) {
    if (!ctx || !dp) return fail(MRPNP_ERR_ARG, "ctx or dense params is NULL%s");
    if (!p) return fail(MRPNP_ERR_ARG, "params is NULL%s");
    mrpnp_params q = *p;
    q.layout = MRPNP_LAYOUT_PLANAR;
    q.weight_mode = MRPNP_W_LOGSTD;
    int rc = check_params(&q);
    if (rc != MRPNP_OK) return rc;
    if (!dims || !rois) return fail(MRPNP_ERR_ARG, "rois / dims is NULL%s");
    if (dp->roi_w <= 0 || p->n_pts % dp->roi_w != 0) return fail(MRPNP_ERR_ARG, "n_pts is not a multiple of roi_w%s");
    if (!(dp->scaling_denominator > 0.f)) return fail(MRPNP_ERR_ARG, "scaling_denominator must be positive%s");
    MR_CUDA(cudaSetDevice(ctx->device));
    g_err[0] = 0;
    const int C = dp->num_classes;
    if (C < 0) return fail(MRPNP_ERR_ARG, "num_classes < 0%s");
    if (C > 0) {
        if (!labels) return fail(MRPNP_ERR_ARG, "labels is NULL with num_classes > 0%s");
        if (dp->pred_stride < (int64_t)5 * C * p->n_pts) return fail(MRPNP_ERR_ARG, "pred_stride smaller than 5*C*H*W%s");
        if (p->n_pts % 4 == 0 && dp->pred_stride % 4 != 0)  // keeps every object's slice 16-byte aligned for the bulk copies
            return fail(MRPNP_ERR_ARG, "pred_stride must be a multiple of 4 floats%s");
        proj_logstd = noc_pred + (size_t)3 * C * p->n_pts;
    }
    if (dp->dim_means) {
        if (!dp->dim_stds || !labels || dp->n_dim_classes < 1)
            return fail(MRPNP_ERR_ARG, "dim_means needs dim_stds, labels and n_dim_classes >= 1%s");
    }
    const DenseArgs da{dp, dims, dims_var, distance, C > 0 ? labels : nullptr, dp->dim_means ? labels : nullptr};
    // alignment for the TMA path is decided on the two streamed tensors; `rois` rides in the coords_2d slot
    return solve_device(ctx, &q, noc_pred, rois, proj_logstd, cam_mats, uv_range, init_pose, nullptr, result,
                        inlier_out, nullptr, static_cast<cudaStream_t>(stream), &da);
}

// The reference's own native entry point (monorun/ops/least_squares/src/ext.h:1-13), GPU-backed: one object, host fp64
// buffers, no return code -- success is *result_val != 0 (pnp_uncert_cpu.cpp:276, :287).  A process-wide context on the
// current CUDA device serves these calls (the reference op is re-entrant; calls here are serialised by a mutex).  The
// solve runs as MRPNP_PREC_FP64 on every point handed in (the reference passes the points it wants solved,
// pnp_uncert_cpu.py:62-66), Ceres-style covariance when result_cov is not NULL (pnp_uncert_cpu.cpp:279-291).
void pnp_uncert(double* pts2d, double* pts3d, double* wgt2d, double* K, double* init_pose, int* result_val,
                double* result_pose, double* result_cov, double* result_tr, int pn, double* clips) {
    static std::mutex mu;
    static mrpnp_ctx* ctx = nullptr;
    static float* dbuf = nullptr;   // device: [pts3d 3P | pts2d 2P | wgt 2P | cam 9 | range 4 | init 4 | result 24] + result64 after it
    static float* hbuf = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (result_val) *result_val = 0;
    for (int i = 0; i < 4; ++i) result_pose[i] = init_pose[i];   // pnp_uncert_cpu.cpp:259
    if (pn < 4 || pn > MRPNP_MAX_POINTS) { fail(MRPNP_ERR_ARG, "pnp_uncert: pn outside [4, 1024]%s"); return; }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    if (!ctx && mrpnp_create(&ctx, dev) != MRPNP_OK) return;
    const size_t cap = 7 * MRPNP_MAX_POINTS + 9 + 4 + 4 + MRPNP_RESULT_STRIDE + 16;
    if (!dbuf && (cudaMalloc(&dbuf, cap * sizeof(float) + 64) != cudaSuccess || cudaMallocHost(&hbuf, cap * sizeof(float) + 64) != cudaSuccess)) return;
    const size_t P = (size_t)pn;
    float* h3 = hbuf; float* h2 = h3 + 3 * P; float* hw = h2 + 2 * P; float* hc = hw + 2 * P; float* hr = hc + 9; float* hi = hr + 4;
    for (size_t i = 0; i < 3 * P; ++i) h3[i] = (float)pts3d[i];
    for (size_t i = 0; i < 2 * P; ++i) { h2[i] = (float)pts2d[i]; hw[i] = (float)wgt2d[i]; }
    for (int i = 0; i < 9; ++i) hc[i] = (float)K[i];
    for (int i = 0; i < 4; ++i) { hr[i] = (float)clips[1 + i]; hi[i] = (float)init_pose[i]; }
    const size_t n_in = 7 * P + 17;
    float* d_res = dbuf + ((n_in + 3) & ~size_t(3));
    double* d_res64 = reinterpret_cast<double*>(dbuf + ((n_in + 3) & ~size_t(3)) + MRPNP_RESULT_STRIDE);   // 8-byte aligned: offsets are multiples of 4 floats
    mrpnp_params p;
    mrpnp_default_params(&p, 1, pn);
    p.layout = MRPNP_LAYOUT_INTERLEAVED;
    p.weight_mode = MRPNP_W_ISTD;
    p.precision = MRPNP_PREC_FP64;
    p.cov_mode = result_cov ? MRPNP_COV_CERES : MRPNP_COV_NONE;
    p.init_mode = MRPNP_INIT_GIVEN;
    p.istd_thres = 0.f;   // every point handed in takes part
    p.z_min = (float)clips[0];
    cudaStream_t st = nullptr;
    if (cudaMemcpyAsync(dbuf, hbuf, n_in * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess) return;
    if (mrpnp_solve(ctx, &p, dbuf, dbuf + 3 * P, dbuf + 5 * P, dbuf + 7 * P, dbuf + 7 * P + 9, dbuf + 7 * P + 13, nullptr, d_res, nullptr,
                    d_res64, st) != MRPNP_OK) return;
    float row[MRPNP_RESULT_STRIDE];
    double r64[8];
    if (cudaMemcpyAsync(row, d_res, sizeof(row), cudaMemcpyDeviceToHost, st) != cudaSuccess) return;
    if (cudaMemcpyAsync(r64, d_res64, sizeof(r64), cudaMemcpyDeviceToHost, st) != cudaSuccess) return;
    if (cudaStreamSynchronize(st) != cudaSuccess) return;
    for (int i = 0; i < 4; ++i) result_pose[i] = r64[i];
    if (result_cov) for (int i = 0; i < 16; ++i) result_cov[i] = (double)row[4 + i];
    if (result_tr) *result_tr = r64[5];
    if (result_val) *result_val = row[20] > 0.5f ? 1 : 0;
}

// Host-buffer entry: objects are cut into chunks; chunk i+1's host->device copies run on the other
// stream while chunk i is being solved, and each chunk's result rows are copied back as soon as its
// kernel finishes.  Pinned host memory makes the copies truly asynchronous; pageable memory also works.
int mrpnp_solve_host(mrpnp_ctx* ctx, const mrpnp_params* p, const float* coords_3d, const float* coords_2d,
                     const float* weights, const float* cam_mats, const float* uv_range, const float* init_pose,
                     const uint32_t* inlier_in, float* result, uint32_t* inlier_out) {
    if (!ctx) return fail(MRPNP_ERR_ARG, "ctx is NULL%s");
    int rc = check_params(p);
    if (rc != MRPNP_OK) return rc;
    if (p->n_obj == 0) return MRPNP_OK;
    if (!coords_3d || !coords_2d || !weights || !cam_mats || !uv_range || !result)
        return fail(MRPNP_ERR_ARG, "NULL tensor pointer%s");
    if (p->init_mode == MRPNP_INIT_GIVEN && !init_pose) return fail(MRPNP_ERR_ARG, "init_pose is NULL with MRPNP_INIT_GIVEN%s");
    MR_CUDA(cudaSetDevice(ctx->device));
    g_err[0] = 0;
    const int wc = p->weight_mode == MRPNP_W_FULL ? 3 : 2;
    const size_t P = (size_t)p->n_pts;
    const int chunk = std::min(p->n_obj, 2048);
    // per-chunk device layout: c3d | c2d | wgt | init | result | inl_in | inl_out (each 256-B aligned)
    auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
    const size_t o3 = 0, o2 = o3 + al(chunk * 3 * P * 4), ow = o2 + al(chunk * 2 * P * 4);
    const size_t oi = ow + al(chunk * wc * P * 4), orr = oi + al((size_t)chunk * 16);
    const size_t W = (P + 31) / 32;  // packed mask words per object
    const size_t omi = orr + al((size_t)chunk * MRPNP_RESULT_STRIDE * 4), omo = omi + al(chunk * W * 4);
    const size_t total = omo + al(chunk * W * 4);
    if (total > ctx->chunk_bytes) {
        for (int i = 0; i < kHostStreams; ++i) {
            if (ctx->chunk_buf[i]) MR_CUDA(cudaFree(ctx->chunk_buf[i]));
            ctx->chunk_buf[i] = nullptr;
            MR_CUDA(cudaMalloc(&ctx->chunk_buf[i], total));
        }
        ctx->chunk_bytes = total;
    }
    for (int i = 0; i < kHostStreams; ++i)
        if (!ctx->streams[i]) MR_CUDA(cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking));
    const size_t cam_n = p->cam_stride ? (size_t)p->n_obj * 9 : 9, rng_n = p->range_stride ? (size_t)p->n_obj * 4 : 4;
    const size_t small = (cam_n + rng_n) * 4;
    if (small > ctx->small_bytes) {
        if (ctx->small_buf) MR_CUDA(cudaFree(ctx->small_buf));
        ctx->small_buf = nullptr;
        MR_CUDA(cudaMalloc(&ctx->small_buf, small));
        ctx->small_bytes = small;
    }
    float* d_cam = ctx->small_buf;
    float* d_rng = ctx->small_buf + cam_n;
    MR_CUDA(cudaMemcpyAsync(d_cam, cam_mats, cam_n * 4, cudaMemcpyHostToDevice, ctx->streams[0]));
    MR_CUDA(cudaMemcpyAsync(d_rng, uv_range, rng_n * 4, cudaMemcpyHostToDevice, ctx->streams[0]));
    MR_CUDA(cudaStreamSynchronize(ctx->streams[0]));

    int ci = 0;
    for (int start = 0; start < p->n_obj; start += chunk, ++ci) {
        const int n = std::min(chunk, p->n_obj - start);
        cudaStream_t st = ctx->streams[ci % kHostStreams];
        char* base = static_cast<char*>(ctx->chunk_buf[ci % kHostStreams]);
        float* d3 = (float*)(base + o3); float* d2 = (float*)(base + o2); float* dw = (float*)(base + ow);
        float* di = (float*)(base + oi); float* dr = (float*)(base + orr);
        uint32_t* dmi = (uint32_t*)(base + omi); uint32_t* dmo = (uint32_t*)(base + omo);
        MR_CUDA(cudaMemcpyAsync(d3, coords_3d + (size_t)start * 3 * P, (size_t)n * 3 * P * 4, cudaMemcpyHostToDevice, st));
        MR_CUDA(cudaMemcpyAsync(d2, coords_2d + (size_t)start * 2 * P, (size_t)n * 2 * P * 4, cudaMemcpyHostToDevice, st));
        MR_CUDA(cudaMemcpyAsync(dw, weights + (size_t)start * wc * P, (size_t)n * wc * P * 4, cudaMemcpyHostToDevice, st));
        if (init_pose) MR_CUDA(cudaMemcpyAsync(di, init_pose + (size_t)start * 4, (size_t)n * 16, cudaMemcpyHostToDevice, st));
        if (inlier_in) MR_CUDA(cudaMemcpyAsync(dmi, inlier_in + (size_t)start * W, (size_t)n * W * 4, cudaMemcpyHostToDevice, st));
        mrpnp_params q = *p;
        q.n_obj = n;
        rc = solve_device(ctx, &q, d3, d2, dw, d_cam + (p->cam_stride ? (size_t)start * 9 : 0),
                          d_rng + (p->range_stride ? (size_t)start * 4 : 0), init_pose ? di : nullptr,
                          inlier_in ? dmi : nullptr, dr, inlier_out ? dmo : nullptr, nullptr, st);
        if (rc != MRPNP_OK) return rc;
        MR_CUDA(cudaMemcpyAsync(result + (size_t)start * MRPNP_RESULT_STRIDE, dr, (size_t)n * MRPNP_RESULT_STRIDE * 4,
                                cudaMemcpyDeviceToHost, st));
        if (inlier_out) MR_CUDA(cudaMemcpyAsync(inlier_out + (size_t)start * W, dmo, (size_t)n * W * 4, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < kHostStreams; ++i) MR_CUDA(cudaStreamSynchronize(ctx->streams[i]));
    return MRPNP_OK;
}

}  // extern "C"
