// Dense correspondence head (FCNNOCDecoder, fcn_noc_decoder.py:189-240) as sm_100a kernels.
//
// Activation layout ("padded-flat NHWC"): every RoI map of h x w pixels is stored with a one-pixel zero halo as
// hp x wp = (h+2) x (w+2) rows of C contiguous bf16 channels, RoIs back to back:
//     row(n, y, x) = (n * hp + y) * wp + x,   interior = 1 <= y <= h, 1 <= x <= w.
// In this layout a 3x3 convolution is nine GEMMs that read the SAME [rows, C] matrix shifted by
// (ky - 1) * wp + (kx - 1) rows, so the A operand of every tap is a plain 2-D TMA tile (out-of-range rows are
// zero-filled by the TMA unit) and no im2col buffer exists.  Outputs are computed for halo rows too and zeroed in
// the epilogue, which re-creates the next layer's padding for free.
#pragma once
#include "head_tc.cuh"

namespace mrhead {

constexpr int kBlockM = 256;   // rows per CTA tile = two UMMA M=128 halves that share every weight tile
constexpr int kBlockK = 64;    // bf16 channels per pipeline stage = one 128-byte swizzled row
constexpr int kUmmaK = 16;
constexpr int kConvThreads = 384;  // warp 0: TMA producer, warp 1: MMA issuer, warp 2: TMEM allocator, warps 4-11: epilogue
constexpr int kEpilogueWarps = 8;

enum OutMode : int {
    kOutBf16Rows = 0,   // bf16 [rows, cout]         padded-flat, halo rows written as zeros
    kOutF32Rows = 1,    // fp32 [rows, cout_pad]     padded-flat (CARAFE kernel logits)
    kOutF32Planar = 2,  // fp32 [n, cout, h, w]      NCHW without halo (the head's all_pred)
};

struct ConvParams {
    int rows_total;          // n * hp * wp
    int hp, wp, h, w;
    int cin, cout, cout_pad; // cout_pad: multiple of 16, <= 256 (rows of each tap's weight tile)
    int taps;                // 1 (1x1) or 9 (3x3)
    int num_tiles, stages, relu, out_mode;
    uint32_t tmem_cols;      // power of two >= 2 * cout_pad
    const float* bias;       // [cout] or NULL
    const float* row_bias;   // [n, cout] added AFTER the activation (latent vector, fcn_noc_decoder.py:205-209) or NULL
    void* out;
};

__device__ __forceinline__ size_t conv_stage_bytes(int cout_pad) {
    return (size_t)2 * 128 * 128 + (size_t)cout_pad * 128;  // A0 + A1 (128 rows x 128 B each) + B (cout_pad rows x 128 B)
}

// Epilogue warps (4..11) of both convolution kernels: TMEM -> registers -> bias / ReLU / latent bias -> global, tile by
// tile; arrives on tmem_empty when a tile's accumulators have been drained.
// BM = 256: one accumulator set, warps 4-7 drain rows 0-127 (columns [0, cout_pad)), warps 8-11 rows 128-255
// (columns [cout_pad, 2 cout_pad)).  BM = 128: TWO accumulator sets of cout_pad columns used by alternate tiles (full /
// empty barrier pair per set), so that this epilogue overlaps the MMAs of the next tile; warps 4-7 drain the first half
// of the columns, warps 8-11 the second half.
template <int BM>
__device__ __forceinline__ void conv_epilogue(const ConvParams& cp, uint32_t tmem_base, uint64_t* tmem_full,
                                              uint64_t* tmem_empty, int warp, int lane) {
        const int quarter = warp & 3;            // TMEM lanes this warp may access: 32 * (warp % 4) ...
        const int half = (warp - 4) >> 2;
        const int roi_rows = cp.hp * cp.wp;
        const int c_begin = BM == 256 ? 0 : half * (cp.cout_pad / 2), c_end = BM == 256 ? cp.cout_pad : c_begin + cp.cout_pad / 2;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < cp.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = BM == 256 ? 0u : (it & 1u), tphase = BM == 256 ? (it & 1u) : ((it >> 1) & 1u);
            mbar_wait(tmem_full + buf, tphase);
            tc_fence_after();
            const int row = tile * BM + (BM == 256 ? half * 128 : 0) + quarter * 32 + lane;
            const int roi = row / roi_rows;
            const int pos = row - roi * roi_rows;
            const int y = pos / cp.wp, x = pos - y * cp.wp;
            const bool in_range = row < cp.rows_total;
            const bool interior = in_range && y >= 1 && y <= cp.h && x >= 1 && x <= cp.w;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((BM == 256 ? half : (int)buf) * cp.cout_pad);
            // 32 accumulator columns per TMEM round trip (two 16-column loads, one wait); bias and per-RoI latent bias
            // come in as float4 (every lane reads the same bias: broadcast loads)
            const float* rb = (cp.row_bias && in_range) ? cp.row_bias + (size_t)roi * cp.cout : nullptr;
            for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                uint32_t v[32];
                const bool two = c0 + 16 < c_end;
                tmem_ld16(taddr + (uint32_t)c0, v);   // whole warp, also for rows past the end
                if (two) tmem_ld16(taddr + (uint32_t)(c0 + 16), v + 16);
                tmem_ld_wait();
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    if (hh == 1 && !two) break;
                    const int cb = c0 + 16 * hh;
                    float f[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[16 * hh + j]);
                    if (cb + 16 <= cp.cout && (cp.cout & 3) == 0) {   // full chunk: vector loads
                        if (cp.bias) {
                            const float4* b4 = reinterpret_cast<const float4*>(cp.bias + cb);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float4 t = __ldg(b4 + q);
                                f[4 * q] += t.x; f[4 * q + 1] += t.y; f[4 * q + 2] += t.z; f[4 * q + 3] += t.w;
                            }
                        }
                        if (cp.relu) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
                        }
                        if (rb) {
                            const float4* r4 = reinterpret_cast<const float4*>(rb + cb);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float4 t = __ldg(r4 + q);
                                f[4 * q] += t.x; f[4 * q + 1] += t.y; f[4 * q + 2] += t.z; f[4 * q + 3] += t.w;
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int c = cb + j;
                            if (c < cp.cout) {
                                if (cp.bias) f[j] += __ldg(cp.bias + c);
                                if (cp.relu) f[j] = fmaxf(f[j], 0.f);
                                if (rb) f[j] += __ldg(rb + c);
                            }
                        }
                    }
                    if (!interior) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = 0.f;
                    }
                    if (cp.out_mode == kOutBf16Rows) {
                        if (in_range && cb < cp.cout) {
                            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(cp.out) + (size_t)row * cp.cout + cb);
                            dst[0] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                            dst[1] = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
                        }
                    } else if (cp.out_mode == kOutF32Rows) {
                        if (in_range) {
                            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(cp.out) + (size_t)row * cp.cout_pad + cb);
#pragma unroll
                            for (int j = 0; j < 4; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                        }
                    } else {
                        if (interior) {
                            float* dst = reinterpret_cast<float*>(cp.out) + ((size_t)roi * cp.cout * cp.h + (y - 1)) * cp.w + (x - 1);
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (cb + j < cp.cout) dst[(size_t)(cb + j) * cp.h * cp.w] = f[j];
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + buf);
        }
}

// D[rows, cout] = act( sum_taps A[rows + shift(tap), cin] * W_tap[cout, cin]^T + bias ) (+ row_bias), bf16 operands,
// fp32 accumulation in TMEM.  Persistent: CTA b handles tiles b, b + gridDim.x, ...
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_wgt,
                 const __grid_constant__ ConvParams cp) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment of every operand tile is what the 128-byte swizzle needs
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const size_t stage_bytes = conv_stage_bytes(cp.cout_pad);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)cp.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + cp.stages;
    uint64_t* tmem_full = empty_bar + cp.stages;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kchunks = cp.cin / kBlockK;
    const int k_iters = cp.taps * kchunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_act);
        tma_prefetch_desc(&tmap_wgt);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < cp.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, kEpilogueWarps);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, cp.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < cp.num_tiles; tile += gridDim.x) {
                const int m0 = tile * kBlockM;
                for (int it = 0; it < k_iters; ++it) {
                    const int tap = it / kchunks, kc = it - tap * kchunks;
                    const int shift = (cp.taps == 9) ? (tap / 3 - 1) * cp.wp + (tap % 3 - 1) : 0;
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    uint8_t* st = smem + (size_t)stage * stage_bytes;
                    mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
                    tma_load_2d(st, &tmap_act, kc * kBlockK, m0 + shift, &full_bar[stage]);
                    tma_load_2d(st + 128 * 128, &tmap_act, kc * kBlockK, m0 + 128 + shift, &full_bar[stage]);
                    tma_load_2d(st + 2 * 128 * 128, &tmap_wgt, kc * kBlockK, tap * cp.cout_pad, &full_bar[stage]);
                    if (++stage == cp.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, cp.cout_pad);
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            for (int tile = blockIdx.x; tile < cp.num_tiles; tile += gridDim.x) {
                mbar_wait(tmem_empty, tphase ^ 1u);  // epilogue has drained the previous tile's accumulators
                tc_fence_after();
                for (int it = 0; it < k_iters; ++it) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    uint8_t* st = smem + (size_t)stage * stage_bytes;
                    const uint64_t b0 = umma_desc_k128(st + 2 * 128 * 128);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint64_t a0 = umma_desc_k128(st + half * 128 * 128);
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            // advancing 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (>>4) address field
                            umma_bf16(tmem_base + (uint32_t)(half * cp.cout_pad), a0 + (uint64_t)(2 * k),
                                      b0 + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);  // frees the stage once these MMAs have read it
                    if (++stage == cp.stages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tmem_full);  // accumulators complete
                tphase ^= 1u;
            }
        }
    } else if (warp >= 4) {
        conv_epilogue<256>(cp, tmem_base, tmem_full, tmem_empty, warp, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, cp.tmem_cols);
}

// ------------------------------------------------------------------ 3x3 convolution with ONE activation load per K chunk
// conv_gemm_kernel reads the activation tile once per TAP: nine loads of rows that are the same rows shifted by
// (ky-1) wp + (kx-1).  Here the BM rows are loaded ONCE per 64-channel chunk together with their (wp + 1)-row halo on
// either side (64-row TMA boxes), and the nine taps are nine UMMA operand views INTO that tile: the matrix descriptor's
// start address moves by whole 128-byte rows.  (The 128-byte swizzle is a function of the absolute shared-memory address
// bits, for TMA writes and UMMA reads alike, so a view that starts mid-pattern needs nothing else: verified on B200 by
// tests/test_head_gpu.py; setting the descriptor's base-offset field to the row phase gives WRONG results.)
// Only the weight tiles still stream per tap: 328 KB instead of 576 KB of L2 reads per 256 rows and chunk.  That alone
// changed nothing (2.411 vs 2.416 ms per 1024 RoIs: the per-tap kernel was not L2-bound, profiles/r02_head_variants.txt);
// what it buys is BM = 128: with half the rows per tile the accumulator needs cout_pad <= 256 of the 512 TMEM columns,
// TWO sets fit, and the epilogue of tile i runs under the MMAs of tile i + 1 -- at +8 % L2 traffic instead of +50 %.
constexpr int kHaloBox = 64;   // rows per TMA box of the activation super-tile

struct ReuseLayout {
    int a_boxes, a_lead;       // boxes per super-tile, rows in front of the tile's first row
    int a_stages, b_stages;
    size_t a_bytes, b_bytes;
};

__host__ __device__ inline ReuseLayout reuse_layout(int bm, int wp, int cout_pad, size_t budget, int taps = 9) {
    ReuseLayout L;
    L.a_lead = taps == 9 ? wp + 1 : 0;   // 1x1: no halo, the tile itself
    L.a_boxes = (bm + 2 * L.a_lead + kHaloBox - 1) / kHaloBox;
    L.a_bytes = (size_t)L.a_boxes * kHaloBox * 128;
    L.b_bytes = (size_t)cout_pad * 128;
    L.a_stages = taps == 9 ? 2 : 6;   // 1x1 layers are HBM-bound on the activations: more of them in flight
    long long left = (long long)budget - (long long)(L.a_stages * L.a_bytes);
    L.b_stages = left > 0 ? (int)(left / (long long)L.b_bytes) : 0;
    if (L.b_stages > 9) L.b_stages = 9;
    return L;
}

template <int BM>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_reuse_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_wgt,
                     const __grid_constant__ ConvParams cp) {
    static_assert(BM == 128 || BM == 256, "tile height");
    constexpr int kSets = BM == 128 ? 2 : 1;     // accumulator sets
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const ReuseLayout L = reuse_layout(BM, cp.wp, cp.cout_pad, (size_t)cp.stages, cp.taps);   // cp.stages carries the operand budget in bytes
    uint8_t* a_tiles = smem;
    uint8_t* b_tiles = smem + (size_t)L.a_stages * L.a_bytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(b_tiles + (size_t)L.b_stages * L.b_bytes);
    uint64_t* a_empty = a_full + L.a_stages;
    uint64_t* b_full = a_empty + L.a_stages;
    uint64_t* b_empty = b_full + L.b_stages;
    uint64_t* tmem_full = b_empty + L.b_stages;
    uint64_t* tmem_empty = tmem_full + kSets;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + kSets);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kchunks = cp.cin / kBlockK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_act);
        tma_prefetch_desc(&tmap_wgt);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < L.a_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < L.b_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < kSets; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kEpilogueWarps); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, cp.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int as = 0, bs = 0;
            uint32_t aphase = 0, bphase = 0;
            for (int tile = blockIdx.x; tile < cp.num_tiles; tile += gridDim.x) {
                const int m0 = tile * BM;
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&a_empty[as], aphase ^ 1u);
                    uint8_t* at = a_tiles + (size_t)as * L.a_bytes;
                    mbar_expect_tx(&a_full[as], (uint32_t)L.a_bytes);
                    for (int b = 0; b < L.a_boxes; ++b)   // rows before the first / after the last are zero-filled
                        tma_load_2d(at + (size_t)b * kHaloBox * 128, &tmap_act, kc * kBlockK, m0 - L.a_lead + b * kHaloBox, &a_full[as]);
                    if (++as == L.a_stages) { as = 0; aphase ^= 1u; }
                    for (int tap = 0; tap < cp.taps; ++tap) {
                        mbar_wait(&b_empty[bs], bphase ^ 1u);
                        mbar_expect_tx(&b_full[bs], (uint32_t)L.b_bytes);
                        tma_load_2d(b_tiles + (size_t)bs * L.b_bytes, &tmap_wgt, kc * kBlockK, tap * cp.cout_pad, &b_full[bs]);
                        if (++bs == L.b_stages) { bs = 0; bphase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, cp.cout_pad);
            int as = 0, bs = 0;
            uint32_t aphase = 0, bphase = 0, it = 0;
            for (int tile = blockIdx.x; tile < cp.num_tiles; tile += gridDim.x, ++it) {
                const uint32_t set = kSets == 2 ? (it & 1u) : 0u, tphase = kSets == 2 ? ((it >> 1) & 1u) : (it & 1u);
                mbar_wait(&tmem_empty[set], tphase ^ 1u);   // the epilogue has drained this set (two tiles ago for BM = 128)
                tc_fence_after();
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&a_full[as], aphase);
                    tc_fence_after();
                    const uint8_t* at = a_tiles + (size_t)as * L.a_bytes;
                    for (int tap = 0; tap < cp.taps; ++tap) {
                        mbar_wait(&b_full[bs], bphase);
                        tc_fence_after();
                        const uint64_t b0 = umma_desc_k128(b_tiles + (size_t)bs * L.b_bytes);
                        const int row0 = cp.taps == 9 ? L.a_lead + (tap / 3 - 1) * cp.wp + (tap % 3 - 1) : 0;
#pragma unroll
                        for (int half = 0; half < BM / 128; ++half) {
                            const uint64_t a0 = umma_desc_k128(at + (size_t)(row0 + half * 128) * 128);
                            const uint32_t d = tmem_base + (uint32_t)((BM == 256 ? half : (int)set) * cp.cout_pad);
#pragma unroll
                            for (int k = 0; k < kBlockK / kUmmaK; ++k)
                                umma_bf16(d, a0 + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), idesc, (kc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit(&b_empty[bs]);
                        if (++bs == L.b_stages) { bs = 0; bphase ^= 1u; }
                    }
                    umma_commit(&a_empty[as]);   // all taps of this chunk have read the super-tile
                    if (++as == L.a_stages) { as = 0; aphase ^= 1u; }
                }
                umma_commit(&tmem_full[set]);
            }
        }
    } else if (warp >= 4) {
        conv_epilogue<BM>(cp, tmem_base, tmem_full, tmem_empty, warp, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, cp.tmem_cols);
}

// ------------------------------------------------------------------ NCHW fp32 -> padded-flat NHWC bf16
// CTA (n, chunk) produces padded rows [chunk * kPackRows, +kPackRows) of RoI n.  Their interior pixels are one
// contiguous range of source pixels, which is transposed through shared memory so that both sides are coalesced.
constexpr int kPackRows = 256;
__global__ void pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int c, int h, int w) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(smem_raw);   // [pixels of this chunk][c + 2]
    const int n = blockIdx.x, hw = h * w, ld = c + 2, hp = h + 2, wp = w + 2;
    const int r0 = blockIdx.y * kPackRows, r1 = min(r0 + kPackRows, hp * wp);
    // first / last interior pixel among the padded rows [r0, r1)
    int pmin = hw, pmax = -1;
    {
        const int y0 = r0 / wp, x0 = r0 - y0 * wp;            // first row: next interior position at or after it
        int yy = y0, xx = x0;
        if (xx > w) { ++yy; xx = 1; }
        if (xx < 1) xx = 1;
        if (yy < 1) { yy = 1; xx = 1; }
        if (yy <= h) pmin = (yy - 1) * w + (xx - 1);
        const int rl = r1 - 1, y1 = rl / wp, x1 = rl - y1 * wp;  // last row: previous interior position at or before it
        yy = y1; xx = x1;
        if (xx < 1) { --yy; xx = w; }
        if (xx > w) xx = w;
        if (yy > h) { yy = h; xx = w; }
        if (yy >= 1) pmax = (yy - 1) * w + (xx - 1);
    }
    const int np = pmax - pmin + 1;
    const float* src = x + (size_t)n * c * hw;
    // one channel per warp iteration (no per-element division), lanes along the contiguous pixels
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    if (np > 0) {   // a channel PAIR per warp iteration: two coalesced loads, one packed 32-bit store (c is even)
        for (int ch = 2 * warp; ch < c; ch += 2 * nwarps) {
            const float* sp = src + (size_t)ch * hw + pmin;
            for (int p = lane; p < np; p += 32)
                *reinterpret_cast<__nv_bfloat162*>(tile + p * ld + ch) = __floats2bfloat162_rn(__ldg(sp + p), __ldg(sp + hw + p));
        }
    }
    __syncthreads();
    __nv_bfloat16* dst = out + (size_t)n * hp * wp * c;
    const int pairs = c / 2;
    // one padded row per warp iteration, lanes along the channel pairs
    for (int r = r0 + warp; r < r1; r += nwarps) {
        const int y = r / wp, xx = r - y * wp;
        const bool interior = y >= 1 && y <= h && xx >= 1 && xx <= w;
        const __nv_bfloat16* t = tile + ((y - 1) * w + (xx - 1) - pmin) * ld;
        __nv_bfloat162* drow = reinterpret_cast<__nv_bfloat162*>(dst + (size_t)r * c);
        for (int q = lane; q < pairs; q += 32) {
            __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
            if (interior) v = *reinterpret_cast<const __nv_bfloat162*>(t + 2 * q);
            drow[q] = v;
        }
    }
}

// ------------------------------------------------------------------ latent vector -> per-RoI channel bias
// row_bias[n, co] = b[co] + sum_k act(latent[n, k]) * W[co, k]      (nn.Linear(16 -> 256), fcn_noc_decoder.py:205-209)
__global__ void latent_bias_kernel(const float* __restrict__ latent, const float* __restrict__ w, const float* __restrict__ b,
                                   float* __restrict__ out, int n, int k, int co, int activation) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * co) return;
    const int r = i / co, c = i - r * co;
    float acc = b ? __ldg(b + c) : 0.f;
    for (int j = 0; j < k; ++j) {
        float l = __ldg(latent + (size_t)r * k + j);
        if (activation == 1) l = fmaxf(l, 0.f);
        else if (activation == 2) l = l > 0.f ? l : 0.01f * l;
        acc = fmaf(l, __ldg(w + (size_t)c * k + j), acc);
    }
    out[i] = acc;
}

// ------------------------------------------------------------------ CARAFE: pixel shuffle + softmax + reassembly
// (mmcv.ops.carafe.CARAFEPack.forward after the two convolutions)
//   logits [rows_lo, ld_logits] fp32, channel k * s*s + sy * s + sx of low-res pixel (y, x) belongs to output pixel
//          (s y + sy, s x + sx) and tap k = ky * K + kx                           (F.pixel_shuffle)
//   w_k    = softmax_k(logits)                                                    (softmax over the K*K taps)
//   out[c] = sum_k w_k * feat[(y + ky - K/2, x + kx - K/2), c], zero outside the map  (carafe reassembly, group 1)
// feat: padded-flat bf16 [n, (h+2)(w+2), C]; out: padded-flat bf16 [n, (2h+2)(2w+2), C] with its halo zeroed here.
// One warp per low-res pixel (its s*s = 4 outputs share the 25 source rows); lane l owns channels 8l .. 8l+7 (C = 256).
template <int K, int S>
__global__ void __launch_bounds__(256) carafe_kernel(const __nv_bfloat16* __restrict__ feat, const float* __restrict__ logits,
                                                     __nv_bfloat16* __restrict__ out, int h, int w, int ld_logits) {
    static_assert(K * K <= 32 && S == 2, "one lane per tap, 2x upsampling");
    constexpr int C = 256, KK = K * K, SS = S * S;
    const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int hp = h + 2, wp = w + 2, ho = S * h, wo = S * w, hop = ho + 2, wop = wo + 2;
    const __nv_bfloat16* f = feat + (size_t)n * hp * wp * C;
    const float* lg = logits + (size_t)n * hp * wp * ld_logits;
    __nv_bfloat16* o = out + (size_t)n * hop * wop * C;

    // halo of the output map: 2 (wop + hop) - 4 rows of zeros
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (int i = warp; i < 2 * wop + 2 * (hop - 2); i += nwarps) {
        int y, x;
        if (i < wop) { y = 0; x = i; }
        else if (i < 2 * wop) { y = hop - 1; x = i - wop; }
        else { const int j = i - 2 * wop; y = 1 + (j >> 1); x = (j & 1) ? wop - 1 : 0; }
        reinterpret_cast<uint4*>(o + ((size_t)y * wop + x) * C)[lane] = zero;
    }

    for (int pix = warp; pix < h * w; pix += nwarps) {
        const int y = pix / w, x = pix - y * w;
        const float* lrow = lg + ((size_t)(y + 1) * wp + (x + 1)) * ld_logits;
        // lane k < 25 holds the weights of tap k for the four sub-pixels
        float wgt[SS];
#pragma unroll
        for (int s = 0; s < SS; ++s) {
            const float v = lane < KK ? __ldg(lrow + lane * SS + s) : -INFINITY;
            float m = v;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
            const float e = lane < KK ? __expf(v - m) : 0.f;
            float t = e;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
            wgt[s] = e / t;
        }
        // packed fp32: one FFMA2 (weight as the broadcast operand) per channel pair and sub-pixel
        float2 acc[SS][4];
#pragma unroll
        for (int s = 0; s < SS; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[s][j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < KK; ++k) {
            const int yy = y + k / K - K / 2, xx = x + k % K - K / 2;
            float ws[SS];
#pragma unroll
            for (int s = 0; s < SS; ++s) ws[s] = __shfl_sync(0xffffffffu, wgt[s], k);
            if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;   // warp-uniform
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(f + ((size_t)(yy + 1) * wp + (xx + 1)) * C) + lane);
            const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 v = make_float2(__uint_as_float(u[j] << 16), __uint_as_float(u[j] & 0xffff0000u));
#pragma unroll
                for (int s = 0; s < SS; ++s) acc[s][j] = __ffma2_rn(make_float2(ws[s], ws[s]), v, acc[s][j]);
            }
        }
#pragma unroll
        for (int s = 0; s < SS; ++s) {
            const int oy = S * y + s / S, ox = S * x + s % S;
            const uint4 pk = make_uint4(pack_bf16(acc[s][0].x, acc[s][0].y), pack_bf16(acc[s][1].x, acc[s][1].y),
                                        pack_bf16(acc[s][2].x, acc[s][2].y), pack_bf16(acc[s][3].x, acc[s][3].y));
            reinterpret_cast<uint4*>(o + ((size_t)(oy + 1) * wop + (ox + 1)) * C)[lane] = pk;
        }
    }
}

}  // namespace mrhead
