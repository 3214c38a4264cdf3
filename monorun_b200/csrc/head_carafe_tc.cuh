// CARAFE reassembly on the tensor cores (mmcv.ops.carafe.CARAFEPack.forward after the two convolutions).
//
//   out[(2y+sy, 2x+sx), c] = sum_{ky,kx} w[y,x,(ky,kx),(sy,sx)] * feat[(y+ky-2, x+kx-2), c],   w = softmax over the 25 taps
//
// is 5.1 G fp32 FMAs for 1024 RoIs -- 0.41 ms on the FMA pipes (carafe_kernel, head_kernels.cuh).  Here it is a GEMM with a
// BANDED left operand.  In the padded-flat layout (rows of 256 bf16 channels, (h+2) x (w+2) pixels with a zero halo, w + 2 =
// 16) the sources of the 8 x 16 low-res pixels of rows y0..y0+7 are 208 CONSECUTIVE flat rows: pixel (yy, x), tap (ky, kx)
// reads flat row start + (yy + ky) * 16 + (x + kx), start = (y0 - 1) * 16 - 1; taps that leave the map land on halo zeros
// (or on rows the TMA zero-fills).  So per tile (RoI, y0 in {0, 8}) and sub-pixel s:
//
//   D[128 outputs, 256 ch] = A_s[128, 208] * F[208, 256]
//
// with A_s holding 25 softmax weights per row at positions that are THE SAME for every tile and sub-pixel (they depend on
// (yy, x) only): the A tile is zeroed once per CTA and each build overwrites the same 25 entries of each row.
//   * F: 4 TMA boxes [208 rows x 64 ch] (128-byte swizzle) = an MN-major B operand as it stands (channels contiguous):
//     descriptor LBO = one box (next 64 channels), SBO = 8 rows; instruction descriptor bit 16 (b_major) set.
//   * A_s: K-major, 128-byte swizzle, 4 K blocks of 64 (208 used), written by the builder warps (generic proxy ->
//     fence.proxy.async -> UMMA).  13 tcgen05.mma (M128 N256 K16) per tile and sub-pixel; fp32 accumulation in TMEM, two
//     accumulator sets, so the epilogue of one sub-pixel runs under the MMAs of the next.
//   * warps 12-15 build A_s (one thread per row), warps 4-11 drain the accumulators: TMEM -> bf16 -> the 512-byte output rows.
// 10.4 GFLOP of useful work become 112 GFLOP of tensor work -- still 3x faster than the fp32 pipes.  Weights are rounded
// to bf16 (2^-9 relative), the products accumulate in fp32.
#pragma once
#include "head_tc.cuh"

namespace mrhead {

constexpr int kCtK = 208;                       // flat source rows per tile (13 x 16)
constexpr int kCtKSteps = kCtK / 16;            // 13
constexpr int kCtABytes = 4 * 128 * 128;        // 4 K blocks x 128 rows x 128 B
constexpr int kCtBBox = kCtK * 128;             // one [208 x 64 ch] box
constexpr int kCtBBytes = 4 * kCtBBox;
constexpr int kCtStageBytes = 8 * 32 * 144;     // epilogue staging (per warp 32 rows x 128 B, pitch 144 B)
constexpr int kCtThreads = 512;                 // warp 0: TMA, 1: MMA, 2: TMEM allocator, 3: output halo, 4-11: epilogue, 12-15: build A

struct CarafeTcParams {
    const float* logits;       // [n * hp * wp, ld_logits] fp32, channel k * 4 + s
    __nv_bfloat16* out;        // [n * (2h+2) * (2w+2), 256]
    int n, h, w, ld_logits;
    int num_tiles;             // n * 2
};

// MN-major SWIZZLE_128B descriptor of a [K rows x 64 elements] box stack: LBO = bytes to the next 64 elements along N.
__device__ __forceinline__ uint64_t umma_desc_mn128(const void* smem_tile, uint32_t lbo_bytes) {
    const uint64_t addr = (uint64_t)((smem_u32(smem_tile) & 0x3FFFFu) >> 4);
    return addr | (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(kCtThreads, 1)
carafe_tc_kernel(const __grid_constant__ CUtensorMap tmap_feat, const __grid_constant__ CarafeTcParams cp) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_tile = smem;
    uint8_t* b_tile = smem + kCtABytes;
    uint8_t* stage = b_tile + kCtBBytes;                                  // 8 epilogue warps x 32 rows x 144 B
    uint64_t* b_full = reinterpret_cast<uint64_t*>(stage + kCtStageBytes);
    uint64_t* b_empty = b_full + 1;
    uint64_t* a_full = b_empty + 1;
    uint64_t* a_empty = a_full + 1;
    uint64_t* tmem_full = a_empty + 1;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hp = cp.h + 2, wp = cp.w + 2;           // wp == 16
    const int hop = 2 * cp.h + 2, wop = 2 * cp.w + 2;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_feat);
    if (warp == 1 && lane == 0) {
        mbar_init(b_full, 1); mbar_init(b_empty, 1);
        mbar_init(a_full, 4); mbar_init(a_empty, 1);             // four builder warps
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kEpilogueWarps); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    // the A tile is zero except for 25 entries per row, and those sit at the same places in every build
    for (int i = threadIdx.x; i < kCtABytes / 16; i += kCtThreads) reinterpret_cast<uint4*>(a_tile)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_tiles = (cp.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles blockIdx.x, + gridDim.x, ...
    const int n_iters = my_tiles * 4;

    if (warp == 0) {
        // ===================== TMA producer: the 208 source rows of a tile, 4 boxes of 64 channels =====================
        if (lane == 0) {
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < cp.num_tiles; tile += gridDim.x) {
                const int n = tile >> 1, y0 = (tile & 1) * 8;
                const int start = n * hp * wp + (y0 - 1) * wp - 1;     // may be negative / past the end: zero-filled
                mbar_wait(b_empty, phase ^ 1u);
#ifndef CARAFE_DBG_NO_TMA
                mbar_expect_tx(b_full, (uint32_t)kCtBBytes);
                for (int cb = 0; cb < 4; ++cb) tma_load_2d(b_tile + (size_t)cb * kCtBBox, &tmap_feat, cb * 64, start, b_full);
#else
                (void)start;
                mbar_arrive(b_full);     // timing experiment: no loads
#endif
                phase ^= 1u;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, 256) | (1u << 16);   // B is MN-major
            uint32_t bphase = 0;
            for (int i = 0; i < n_iters; ++i) {
                const int s = i & 3;
                const uint32_t set = (uint32_t)i & 1u, tphase = ((uint32_t)i >> 1) & 1u;
                if (s == 0) { mbar_wait(b_full, bphase); bphase ^= 1u; }
                mbar_wait(a_full, (uint32_t)i & 1u);
                mbar_wait(&tmem_empty[set], tphase ^ 1u);
                tc_fence_after();
#pragma unroll 1
                for (int j = 0; j < kCtKSteps; ++j) {
                    const uint64_t adesc = umma_desc_k128(a_tile + (size_t)(j >> 2) * (128 * 128)) + (uint64_t)(2 * (j & 3));
                    const uint64_t bdesc = umma_desc_mn128(b_tile + (size_t)j * (16 * 128), (uint32_t)kCtBBox);
#ifndef CARAFE_DBG_NO_MMA
                    umma_bf16(tmem_base + set * 256u, adesc, bdesc, idesc, j > 0 ? 1u : 0u);
#endif
                }
                umma_commit(a_empty);              // the A tile may be rebuilt
                umma_commit(&tmem_full[set]);
                if (s == 3) umma_commit(b_empty);  // the source rows may be replaced
            }
        }
    } else if (warp >= 12) {
        // ===================== builders (warps 12-15): one thread per row of A_s =====================
        const int m = threadIdx.x - 384;                 // row of the tile, 0..127
        const int yy = m >> 4, x = m & 15;
        float v[25];     // the logits of the NEXT build: loaded before the wait for the A tile
        auto prefetch = [&](int i) {
            const int tile = (int)blockIdx.x + (i >> 2) * (int)gridDim.x, ss = i & 3;
            const int n = tile >> 1, y = (tile & 1) * 8 + yy;
            const bool live = y < cp.h && x < cp.w;
            const float* lrow = cp.logits + ((size_t)n * hp * wp + (size_t)(y + 1) * wp + (x + 1)) * cp.ld_logits;
#pragma unroll
            for (int k = 0; k < 25; ++k) v[k] = live ? __ldg(lrow + k * 4 + ss) : 0.f;
        };
        auto build = [&](int i) {
            const int tile = (int)blockIdx.x + (i >> 2) * (int)gridDim.x;
            const int y = (tile & 1) * 8 + yy;
            const bool live = y < cp.h && x < cp.w;
            float mx = v[0];
#pragma unroll
            for (int k = 1; k < 25; ++k) mx = fmaxf(mx, v[k]);
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 25; ++k) { v[k] = __expf(v[k] - mx); sum += v[k]; }
            const float inv = live ? 1.f / sum : 0.f;      // rows outside the map carry zero weights (and are not stored)
#pragma unroll
            for (int k = 0; k < 25; ++k) {
                const int kk = (yy + k / 5) * 16 + (x + k % 5);          // column of A
                const uint32_t off = (uint32_t)(kk >> 6) * (128 * 128) + (uint32_t)m * 128 +
                                     ((((uint32_t)(kk & 63) >> 3) ^ ((uint32_t)m & 7u)) << 4) + ((uint32_t)kk & 7u) * 2;
                *reinterpret_cast<__nv_bfloat16*>(a_tile + off) = __float2bfloat16_rn(v[k] * inv);
            }
            fence_proxy_async_smem();      // these generic-proxy writes are read by the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        };
        if (n_iters > 0) { prefetch(0); build(0); }
        for (int i = 0; i + 1 < n_iters; ++i) {
            prefetch(i + 1);
            mbar_wait(a_empty, (uint32_t)i & 1u);          // the MMAs of iteration i have read the A tile
            build(i + 1);
        }
    } else if (warp >= 4) {
        // ===================== epilogue (warps 4-11): TMEM -> bf16 -> the 512-byte output rows, 128 channels per warp =====================
        const int quarter = warp & 3, half = (warp - 4) >> 2;
        for (int i = 0; i < n_iters; ++i) {
            const uint32_t set = (uint32_t)i & 1u, tphase = ((uint32_t)i >> 1) & 1u;
            mbar_wait(&tmem_full[set], tphase);
            tc_fence_after();
            const int tile = (int)blockIdx.x + (i >> 2) * (int)gridDim.x, s = i & 3;
            const int n = tile >> 1;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + set * 256u + (uint32_t)(half * 128);
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 64) {
                uint32_t u[64];
#pragma unroll
                for (int q = 0; q < 4; ++q) tmem_ld16(taddr + (uint32_t)(c0 + 16 * q), u + 16 * q);
                tmem_ld_wait();
                // Through a per-warp staging tile (32 rows x 128 B, pitch 144 B) so that one store instruction writes four
                // rows x 128 contiguous bytes: with a row per lane every instruction wrote 32 half sectors, and the stores
                // were 140 of this kernel's 270 us (profiles/r02_head_variants.txt).
                uint8_t* stg = stage + (size_t)(warp - 4) * (32 * 144);
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<uint4*>(stg + lane * 144 + q * 16) =
                        make_uint4(pack_bf16(__uint_as_float(u[8 * q]), __uint_as_float(u[8 * q + 1])),
                                   pack_bf16(__uint_as_float(u[8 * q + 2]), __uint_as_float(u[8 * q + 3])),
                                   pack_bf16(__uint_as_float(u[8 * q + 4]), __uint_as_float(u[8 * q + 5])),
                                   pack_bf16(__uint_as_float(u[8 * q + 6]), __uint_as_float(u[8 * q + 7])));
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int rr = (lane >> 3) + 4 * j;              // row of this warp's 32
                    const int trow = quarter * 32 + rr;
                    const int ty = (tile & 1) * 8 + (trow >> 4), tx = trow & 15;
                    if (ty < cp.h && tx < cp.w) {
                        const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 144 + (lane & 7) * 16);
                        __nv_bfloat16* dr = cp.out + ((size_t)n * hop * wop + (size_t)(2 * ty + (s >> 1) + 1) * wop + (2 * tx + (s & 1) + 1)) * 256;
                        *reinterpret_cast<uint4*>(dr + half * 128 + c0 + (lane & 7) * 8) = val;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[set]);
        }
    } else if (warp == 3) {
        // ===================== the zero halo of the output maps of this CTA's RoIs =====================
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        for (int tile = blockIdx.x; tile < cp.num_tiles; tile += gridDim.x) {
            if (tile & 1) continue;                        // once per RoI
            __nv_bfloat16* o = cp.out + (size_t)(tile >> 1) * hop * wop * 256;
            for (int i = 0; i < 2 * wop + 2 * (hop - 2); ++i) {
                int y, x;
                if (i < wop) { y = 0; x = i; }
                else if (i < 2 * wop) { y = hop - 1; x = i - wop; }
                else { const int j = i - 2 * wop; y = 1 + (j >> 1); x = (j & 1) ? wop - 1 : 0; }
                reinterpret_cast<uint4*>(o + ((size_t)y * wop + x) * 256)[lane] = zero;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace mrhead
