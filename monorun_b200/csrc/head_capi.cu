// C ABI of libmonorun_head.so (include/monorun_head.h): launch planning, TMA tensor maps, layer sequencing.
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/monorun_head.h"
#include "head_kernels.cuh"
#include "head_carafe_tc.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define MH_CUDA(call)                                                                                 \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return fail(MRHEAD_ERR_CUDA, #call ": %s", cudaGetErrorString(e_));    \
    } while (0)

using EncodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr size_t kSmemBudget = 227 * 1024;
constexpr size_t kSmemTail = 1024 /* alignment slack */ + 512 /* barriers (up to 2 x 6 + 2 x 9 + 4) + TMEM slot */;

}  // namespace

struct mrhead_ctx {
    int device = 0;
    int sm_count = 0;
    EncodeTiled encode = nullptr;
    std::atomic<int64_t> launches{0};
};

namespace {

// bf16 [rows, cols] row-major, box = 64 columns (128 bytes, the swizzle span) x box_rows rows
int make_tmap(mrhead_ctx* ctx, CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    if (reinterpret_cast<uintptr_t>(ptr) % 16 != 0) return fail(MRHEAD_ERR_ARG, "tensor base must be 16-byte aligned");
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * sizeof(__nv_bfloat16)};
    const cuuint32_t box[2] = {(cuuint32_t)mrhead::kBlockK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = ctx->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MRHEAD_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MRHEAD_OK;
}

int check_layer(const mrhead_layer* l) {
    if (!l || !l->weight) return fail(MRHEAD_ERR_ARG, "layer or weight is NULL");
    if (l->taps != 1 && l->taps != 9) return fail(MRHEAD_ERR_ARG, "taps must be 1 or 9");
    if (l->cin <= 0 || l->cin % mrhead::kBlockK != 0) return fail(MRHEAD_ERR_ARG, "cin must be a multiple of 64");
    if (l->cout_pad < 16 || l->cout_pad > 256 || l->cout_pad % 16 != 0) return fail(MRHEAD_ERR_ARG, "cout_pad must be a multiple of 16 in [16, 256]");
    if (l->cout <= 0 || l->cout > l->cout_pad) return fail(MRHEAD_ERR_ARG, "cout outside (0, cout_pad]");
    return MRHEAD_OK;
}

uint32_t tmem_cols_for(int cout_pad) {
    uint32_t c = 32;
    while (c < (uint32_t)(2 * cout_pad)) c <<= 1;
    return c;
}

int conv_launch(mrhead_ctx* ctx, const mrhead_layer* l, const void* in, int n, int h, int w, const float* row_bias,
                int out_mode, void* out, cudaStream_t stream) {
    int rc = check_layer(l);
    if (rc) return rc;
    if (!in || !out) return fail(MRHEAD_ERR_ARG, "NULL activation pointer");
    if (n <= 0 || h <= 0 || w <= 0) return fail(MRHEAD_ERR_ARG, "bad shape");
    if (out_mode < 0 || out_mode > 2) return fail(MRHEAD_ERR_ARG, "bad out_mode");
    if (out_mode == MRHEAD_OUT_BF16_ROWS && l->cout % 16 != 0) return fail(MRHEAD_ERR_ARG, "bf16 row output needs cout % 16 == 0");
    if (reinterpret_cast<uintptr_t>(out) % 16 != 0) return fail(MRHEAD_ERR_ARG, "out must be 16-byte aligned");
    const int hp = h + 2, wp = w + 2;
    const long long rows = (long long)n * hp * wp;
    if (rows > 0x7fffff00LL) return fail(MRHEAD_ERR_ARG, "too many rows for one launch");

    mrhead::ConvParams cp;
    std::memset(&cp, 0, sizeof(cp));
    cp.rows_total = (int)rows;
    cp.hp = hp; cp.wp = wp; cp.h = h; cp.w = w;
    cp.cin = l->cin; cp.cout = l->cout; cp.cout_pad = l->cout_pad; cp.taps = l->taps;
    cp.num_tiles = (int)((rows + mrhead::kBlockM - 1) / mrhead::kBlockM);
    const size_t stage = (size_t)2 * 128 * 128 + (size_t)l->cout_pad * 128;
    int stages = (int)((kSmemBudget - kSmemTail) / stage);
    if (stages > 8) stages = 8;
    if (stages < 2) return fail(MRHEAD_ERR_ARG, "tile does not fit in shared memory");
    cp.stages = stages;
    cp.relu = l->relu; cp.out_mode = out_mode;
    cp.tmem_cols = tmem_cols_for(l->cout_pad);
    cp.bias = l->bias; cp.row_bias = row_bias; cp.out = out;

    // 3x3: one activation load per K chunk, the nine taps as shifted operand views (conv3x3_reuse_kernel); 128-row tiles
    // with two accumulator sets when the channel count allows it.  MRHEAD_CONV = pertap | reuse256 | reuse128 (A/B).
    static const int reuse_mode = [] {
        const char* e = std::getenv("MRHEAD_CONV");
        if (!e) return 128;
        if (!std::strcmp(e, "pertap")) return 0;
        if (!std::strcmp(e, "reuse256")) return 256;
        return 128;
    }();
    // (1x1 layers take the same kernel with one tap and no halo when their channel count allows two accumulator sets:
    // what they gain is the epilogue under the next tile's loads.)
    if (reuse_mode && (l->taps == 9 || (reuse_mode == 128 && l->cout_pad % 32 == 0))) {
        const size_t budget = kSmemBudget - kSmemTail;
        const int bm = (reuse_mode == 128 && l->cout_pad % 32 == 0) ? 128 : 256;
        const mrhead::ReuseLayout L = mrhead::reuse_layout(bm, wp, l->cout_pad, budget, l->taps);
        if (L.b_stages >= 3) {
            cp.stages = (int)budget;
            cp.num_tiles = (int)((rows + bm - 1) / bm);
            CUtensorMap ta, tw;
            rc = make_tmap(ctx, &ta, in, (uint64_t)rows, (uint64_t)l->cin, mrhead::kHaloBox);
            if (rc) return rc;
            rc = make_tmap(ctx, &tw, l->weight, (uint64_t)l->taps * l->cout_pad, (uint64_t)l->cin, (uint32_t)l->cout_pad);
            if (rc) return rc;
            const size_t smem = L.a_stages * L.a_bytes + L.b_stages * L.b_bytes + kSmemTail;
            static std::atomic<int> configured2{0};
            if (!configured2.exchange(1)) {
                MH_CUDA(cudaFuncSetAttribute(mrhead::conv3x3_reuse_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
                MH_CUDA(cudaFuncSetAttribute(mrhead::conv3x3_reuse_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
            }
            const int grid = cp.num_tiles < ctx->sm_count ? cp.num_tiles : ctx->sm_count;
            if (bm == 128) mrhead::conv3x3_reuse_kernel<128><<<grid, mrhead::kConvThreads, smem, stream>>>(ta, tw, cp);
            else mrhead::conv3x3_reuse_kernel<256><<<grid, mrhead::kConvThreads, smem, stream>>>(ta, tw, cp);
            MH_CUDA(cudaGetLastError());
            ctx->launches++;
            return MRHEAD_OK;
        }
    }

    CUtensorMap ta, tw;
    rc = make_tmap(ctx, &ta, in, (uint64_t)rows, (uint64_t)l->cin, 128);
    if (rc) return rc;
    rc = make_tmap(ctx, &tw, l->weight, (uint64_t)l->taps * l->cout_pad, (uint64_t)l->cin, (uint32_t)l->cout_pad);
    if (rc) return rc;

    const size_t smem = (size_t)stages * stage + kSmemTail;
    static std::atomic<size_t> configured{0};
    if (configured.load() < smem) {
        MH_CUDA(cudaFuncSetAttribute(mrhead::conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
        configured.store(kSmemBudget);
    }
    const int grid = cp.num_tiles < ctx->sm_count ? cp.num_tiles : ctx->sm_count;
    mrhead::conv_gemm_kernel<<<grid, mrhead::kConvThreads, smem, stream>>>(ta, tw, cp);
    MH_CUDA(cudaGetLastError());
    ctx->launches++;
    return MRHEAD_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
    size_t a_lo[2], comp, logits, a_hi[2], row_bias, total;
};

Workspace plan_workspace(const mrhead_weights* wts, int n, int h, int w) {
    Workspace ws;
    const size_t rows_lo = (size_t)n * (h + 2) * (w + 2), rows_hi = (size_t)n * (2 * h + 2) * (2 * w + 2);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 1024); return o; };
    ws.a_lo[0] = take(rows_lo * 256 * 2);
    ws.a_lo[1] = take(rows_lo * 256 * 2);
    ws.comp = take(rows_lo * (size_t)wts->compressor.cout * 2);
    ws.logits = take(rows_lo * (size_t)wts->encoder.cout_pad * 4);
    ws.a_hi[0] = take(rows_hi * 256 * 2);
    ws.a_hi[1] = take(rows_hi * 256 * 2);
    ws.row_bias = take((size_t)n * 256 * 4);
    ws.total = off;
    return ws;
}

}  // namespace

extern "C" {

int mrhead_version(void) { return MRHEAD_VERSION; }
const char* mrhead_last_error(void) { return g_err; }

int mrhead_create(mrhead_ctx** out, int device) {
    if (!out) return fail(MRHEAD_ERR_ARG, "out is NULL");
    int count = 0;
    MH_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(MRHEAD_ERR_ARG, "device %d out of range (%d CUDA devices)", device, count);
    MH_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MH_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(MRHEAD_ERR_ARG, "libmonorun_head is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    MH_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !fn) return fail(MRHEAD_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    mrhead_ctx* ctx = new mrhead_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->encode = reinterpret_cast<EncodeTiled>(fn);
    *out = ctx;
    return MRHEAD_OK;
}

void mrhead_destroy(mrhead_ctx* ctx) { delete ctx; }

int64_t mrhead_launch_count(const mrhead_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int mrhead_pack_input(mrhead_ctx* ctx, const float* x, int n, int c, int h, int w, void* out, void* stream) {
    if (!ctx || !x || !out) return fail(MRHEAD_ERR_ARG, "NULL argument");
    if (n <= 0 || c <= 0 || c % 2 != 0 || h <= 0 || w <= 0) return fail(MRHEAD_ERR_ARG, "bad shape");
    MH_CUDA(cudaSetDevice(ctx->device));
    const int hw = h * w;
    const size_t smem = (size_t)(hw < mrhead::kPackRows ? hw : mrhead::kPackRows) * (c + 2) * sizeof(__nv_bfloat16);
    if (smem > kSmemBudget) return fail(MRHEAD_ERR_ARG, "too many channels for the transpose tile");
    static std::atomic<int> configured{0};
    if (!configured.exchange(1))
        MH_CUDA(cudaFuncSetAttribute(mrhead::pack_input_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    const dim3 grid(n, ((h + 2) * (w + 2) + mrhead::kPackRows - 1) / mrhead::kPackRows);
    mrhead::pack_input_kernel<<<grid, 512, smem, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(out), c, h, w);
    MH_CUDA(cudaGetLastError());
    ctx->launches++;
    return MRHEAD_OK;
}

int mrhead_conv(mrhead_ctx* ctx, const mrhead_layer* layer, const void* in, int n, int h, int w, const float* row_bias,
                int out_mode, void* out, void* stream) {
    if (!ctx) return fail(MRHEAD_ERR_ARG, "ctx is NULL");
    MH_CUDA(cudaSetDevice(ctx->device));
    return conv_launch(ctx, layer, in, n, h, w, row_bias, out_mode, out, static_cast<cudaStream_t>(stream));
}

int mrhead_latent_bias(mrhead_ctx* ctx, const float* latent, const float* w, const float* b, int n, int k, int cout,
                       int activation, float* out, void* stream) {
    if (!ctx || !latent || !w || !out) return fail(MRHEAD_ERR_ARG, "NULL argument");
    if (n <= 0 || k <= 0 || cout <= 0) return fail(MRHEAD_ERR_ARG, "bad shape");
    MH_CUDA(cudaSetDevice(ctx->device));
    const int total = n * cout;
    mrhead::latent_bias_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(latent, w, b, out, n, k, cout, activation);
    MH_CUDA(cudaGetLastError());
    ctx->launches++;
    return MRHEAD_OK;
}

int mrhead_carafe(mrhead_ctx* ctx, const void* feat, const float* logits, int ld_logits, int n, int h, int w, void* out,
                  void* stream) {
    if (!ctx || !feat || !logits || !out) return fail(MRHEAD_ERR_ARG, "NULL argument");
    if (n <= 0 || h <= 0 || w <= 0 || ld_logits < 100) return fail(MRHEAD_ERR_ARG, "bad shape");
    MH_CUDA(cudaSetDevice(ctx->device));
    // 14 x 14 maps (every reference config): the reassembly as a banded GEMM on the tensor cores (head_carafe_tc.cuh);
    // MRHEAD_CARAFE=fma keeps the fp32 kernel (A/B, other map sizes)
    static const bool use_tc = [] { const char* e = std::getenv("MRHEAD_CARAFE"); return !(e && !std::strcmp(e, "fma")); }();
    if (use_tc && w == 14 && h <= 16 && h >= 9) {
        const long long rows = (long long)n * (h + 2) * (w + 2);
        CUtensorMap tf;
        int rc = make_tmap(ctx, &tf, feat, (uint64_t)rows, 256, (uint32_t)mrhead::kCtK);
        if (rc) return rc;
        mrhead::CarafeTcParams cp;
        cp.logits = logits; cp.out = static_cast<__nv_bfloat16*>(out);
        cp.n = n; cp.h = h; cp.w = w; cp.ld_logits = ld_logits; cp.num_tiles = 2 * n;
        const size_t smem = (size_t)mrhead::kCtABytes + mrhead::kCtBBytes + mrhead::kCtStageBytes + kSmemTail;
        static std::atomic<int> configured3{0};
        if (!configured3.exchange(1))
            MH_CUDA(cudaFuncSetAttribute(mrhead::carafe_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
        const int grid = cp.num_tiles < ctx->sm_count ? cp.num_tiles : ctx->sm_count;
        mrhead::carafe_tc_kernel<<<grid, mrhead::kCtThreads, smem, static_cast<cudaStream_t>(stream)>>>(tf, cp);
        MH_CUDA(cudaGetLastError());
        ctx->launches++;
        return MRHEAD_OK;
    }
    mrhead::carafe_kernel<5, 2><<<n, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(feat), logits, static_cast<__nv_bfloat16*>(out), h, w, ld_logits);
    MH_CUDA(cudaGetLastError());
    ctx->launches++;
    return MRHEAD_OK;
}

size_t mrhead_workspace_bytes(const mrhead_weights* wts, int n, int h, int w) {
    if (!wts || n <= 0) return 0;
    return plan_workspace(wts, n, h, w).total;
}

int mrhead_forward(mrhead_ctx* ctx, const mrhead_weights* wts, const float* x, const float* latent, int n, int h, int w,
                   void* workspace, size_t workspace_bytes, float* all_pred, void* stream) {
    if (!ctx || !wts || !x || !workspace || !all_pred) return fail(MRHEAD_ERR_ARG, "NULL argument");
    if (n <= 0) return fail(MRHEAD_ERR_ARG, "n must be positive");
    if (wts->num_convs < 1 || wts->num_convs > 4 || wts->num_convs_up < 0 || wts->num_convs_up > 2)
        return fail(MRHEAD_ERR_ARG, "unsupported layer counts");
    if (wts->convs[0].cin != 256 || wts->compressor.cin != 256 || wts->final.cin != 256)
        return fail(MRHEAD_ERR_ARG, "the head is built for 256 feature channels");
    if (wts->encoder.cout != 100 || wts->encoder.cin != wts->compressor.cout)
        return fail(MRHEAD_ERR_ARG, "CARAFE encoder must be 3x3 compressed -> 25*4");
    if (reinterpret_cast<uintptr_t>(workspace) % 1024 != 0) return fail(MRHEAD_ERR_ARG, "workspace must be 1024-byte aligned");
    const Workspace ws = plan_workspace(wts, n, h, w);
    if (workspace_bytes < ws.total) return fail(MRHEAD_ERR_ARG, "workspace too small: %zu < %zu", workspace_bytes, ws.total);
    MH_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* base = static_cast<uint8_t*>(workspace);
    void* lo[2] = {base + ws.a_lo[0], base + ws.a_lo[1]};
    void* hi[2] = {base + ws.a_hi[0], base + ws.a_hi[1]};
    float* row_bias = nullptr;
    int rc;
    if (wts->latent_w && latent) {
        row_bias = reinterpret_cast<float*>(base + ws.row_bias);
        rc = mrhead_latent_bias(ctx, latent, wts->latent_w, wts->latent_b, n, wts->latent_channels, 256, wts->latent_activation, row_bias, stream);
        if (rc) return rc;
    }
    rc = mrhead_pack_input(ctx, x, n, 256, h, w, lo[0], stream);
    if (rc) return rc;
    int cur = 0;
    for (int i = 0; i < wts->num_convs; ++i) {   // fcn_noc_decoder.py:192-204 (+ the latent add of :205-209 on the last one)
        rc = conv_launch(ctx, &wts->convs[i], lo[cur], n, h, w, i + 1 == wts->num_convs ? row_bias : nullptr,
                         MRHEAD_OUT_BF16_ROWS, lo[cur ^ 1], st);
        if (rc) return rc;
        cur ^= 1;
    }
    // CARAFEPack.forward: compressor -> encoder -> (pixel shuffle, softmax, reassembly)
    rc = conv_launch(ctx, &wts->compressor, lo[cur], n, h, w, nullptr, MRHEAD_OUT_BF16_ROWS, base + ws.comp, st);
    if (rc) return rc;
    rc = conv_launch(ctx, &wts->encoder, base + ws.comp, n, h, w, nullptr, MRHEAD_OUT_F32_ROWS, base + ws.logits, st);
    if (rc) return rc;
    rc = mrhead_carafe(ctx, lo[cur], reinterpret_cast<const float*>(base + ws.logits), wts->encoder.cout_pad, n, h, w, hi[0], stream);
    if (rc) return rc;
    int curh = 0;
    for (int i = 0; i < wts->num_convs_up; ++i) {   // :218-219
        rc = conv_launch(ctx, &wts->convs_up[i], hi[curh], n, 2 * h, 2 * w, nullptr, MRHEAD_OUT_BF16_ROWS, hi[curh ^ 1], st);
        if (rc) return rc;
        curh ^= 1;
    }
    return conv_launch(ctx, &wts->final, hi[curh], n, 2 * h, 2 * w, nullptr, MRHEAD_OUT_F32_PLANAR, all_pred, st);  // :220
}

}  // extern "C"
