// pnp_kernel_pair.cuh -- MRPNP_PREC_MIXED production kernel: TWO specialised warps per object.
//
//   warp A ("residual warp"): fp64 residual chain + cost for every point, the whole scalar Levenberg-
//                             Marquardt loop (Ceres 1.14 control flow), covariance, result row.
//   warp B ("Jacobian warp"): independent fp32 projection -> J^T r and J^T J sums + near-clip detection.
//
// Why: shared memory (one 22-25 KB slab per object) caps an SM at 10 objects in flight; with one warp per
// object that is 2.5 warps per scheduler and the kernel is bound by fixed-latency dependency stalls (ncu:
// `wait`).  Splitting the pass by *role* doubles the resident warps without duplicating a single instruction
// of the pass, halves the pass latency, and lets warp B's register budget stay small.  The two warps meet at
// a 64-thread named barrier twice per LM iteration; warp B sleeps at the barrier while A does the 4x4 solve.
// With the maximum shared-memory carve-out L1 is ~3 KB, so local-memory spills are expensive: A parks its LM
// state in the pair's shared header across the pass and the cold exact-fp64 routines take no pointer-to-local.
#pragma once
#include "pnp_kernel.cuh"

namespace mrpnp {

#ifdef MRPNP_TRACE
// debug build only: per-object phase timestamps (clock64) written by lane 0 of warp A into result64[obj*32+i]
#define MR_TRACE(i) do { if (isA && lane == 0 && kp.result64) kp.result64[(size_t)obj * 32 + (i)] = (double)clock64(); } while (0)
#else
#define MR_TRACE(i) do { } while (0)
#endif

constexpr int kPairHeaderBytes = 704;
constexpr int kMaxPairsPerCta = 10;
constexpr int kPairThreads = kMaxPairsPerCta * 64;

struct PairHeader {
    uint64_t bar;        // TMA mbarrier of the pair's slot
    int obj;             // current object
    int go;              // 1: run a pass at pt, 0: object finished
    int flagged;         // B: a point came within the margin of a clip bound in the last pass
    int pad0;
    double pt[4];        // evaluation point (yaw, t)
    double sn, cs;       // sin / cos of pt[0]
    float sums[16];      // B: g[4], upper-tri(J^T J)[10]; doubles as B's reduction scratch
    float wsum[4];       // weight sums: A's (u,v), B's (u,v)
    uint32_t masks[32];  // inlier ballot of each 32-point row
    // A's Levenberg-Marquardt state lives here, not in registers (20 warps/SM leave 96 registers per thread
    // and L1 is ~3 KB next to 225 KB of shared memory, so spills would go to L2):
    double x[4];         // current accepted point
    double gs[4];        // column-scaled gradient at x
    double scale[4];     // Jacobi scaling 1 / (1 + |J_i|) from the initial Jacobian
    double diag[4];      // clamped diag(Js^T Js) of the last accepted point (LM diagonal)
    double Hs[10];       // column-scaled Gauss-Newton matrix at x (upper triangle)
    double xres[18];     // exact-pass results (16 sums + clip flag) / reduction scratch / covariance staging
};
static_assert(sizeof(PairHeader) <= kPairHeaderBytes, "pair header too large");

__device__ __forceinline__ void pair_sync(int barid) {
    asm volatile("bar.sync %0, 64;" ::"r"(barid) : "memory");
}

// ------------------------------------------------------------------ cold paths of warp A (not inlined)
// Exact fp64 pass with full clip semantics at hdr->pt; results: 16 sums -> hdr->xres[0..15], clip -> xres[16].
template <int WMODE, int LAYOUT>
__device__ __noinline__ void pair_exact_pass(const KParams& kp, PairHeader* hdr, const float* slot, int n, int lane,
                                             int clipsem, int use_masks) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    const int P = kp.n_pts;
    const float* s3 = slot;
    const float* s2 = slot + 3 * P;
    const float* sw = slot + 5 * P;
    const Camera<float> cf = load_camera<float>(kp, hdr->obj);
    const double fx = cf.fx, fy = cf.fy, cx = cf.cx, cy = cf.cy, zmin = cf.z_min;
    const double umin = cf.u_min, umax = cf.u_max, vmin = cf.v_min, vmax = cf.v_max;
    double sn, cs;
    sincos(hdr->pt[0], &sn, &cs);
    const double tx = hdr->pt[1], ty = hdr->pt[2], tz = hdr->pt[3];
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.0;
    bool any = false;
    for (int p = lane, k = 0; p < n; p += 32, ++k) {
        if (use_masks && !((hdr->masks[k] >> lane) & 1u)) continue;
        const double X = (double)s3[sidx<LAYOUT, 3>(p, 0, P)], Y = (double)s3[sidx<LAYOUT, 3>(p, 1, P)];
        const double Z = (double)s3[sidx<LAYOUT, 3>(p, 2, P)];
        const double uo = (double)s2[sidx<LAYOUT, 2>(p, 0, P)], vo = (double)s2[sidx<LAYOUT, 2>(p, 1, P)];
        const double qx = fma(cs, X, sn * Z), qz = fma(cs, Z, -sn * X);
        const double xc = qx + tx, yc = Y + ty, zc = qz + tz;
        const bool zfree = !(zc < zmin);
        const double z = zfree ? zc : zmin;
        const double iz = fast_rcp(z);
        const double xn = xc * iz, yn = yc * iz;
        double pu = fma(fx, xn, cx), pv = fma(fy, yn, cy);
        bool ufree = true, vfree = true;
        if (pu < umin) { pu = umin; ufree = false; } else if (pu > umax) { pu = umax; ufree = false; }
        if (pv < vmin) { pv = vmin; vfree = false; } else if (pv > vmax) { pv = vmax; vfree = false; }
        any = any || !zfree || !ufree || !vfree;
        const double du = pu - uo, dv = pv - vo;
        const double mz = zfree ? 1.0 : 0.0;
        if (clipsem == 1 && !zfree) { ufree = false; vfree = false; }
        const double au = ufree ? fx * iz : 0.0, av = vfree ? fy * iz : 0.0;
        const double bu = -au * xn * mz, bv = -av * yn * mz;
        const double ju0 = fma(au, qz, -bu * qx), jv0 = -bv * qx;
        const double w0 = (double)sw[sidx<LAYOUT, WC>(p, 0, P)], w1 = (double)sw[sidx<LAYOUT, WC>(p, 1, P)];
        const double wxx = w0, wxy = (WMODE == MRPNP_W_FULL) ? w1 : 0.0;
        const double wyy = (WMODE == MRPNP_W_FULL) ? (double)sw[sidx<LAYOUT, WC>(p, WC - 1, P)] : w1;
        const double r0 = fma(wxx, du, wxy * dv), r1 = fma(wxy, du, wyy * dv);
        const double ja[4] = {fma(wxx, ju0, wxy * jv0), wxx * au, wxy * av, fma(wxx, bu, wxy * bv)};
        const double jb[4] = {fma(wxy, ju0, wyy * jv0), wxy * au, wyy * av, fma(wxy, bu, wyy * bv)};
        acc[0] = fma(r0, r0, fma(r1, r1, acc[0]));
        int q = 5;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[1 + i] = fma(ja[i], r0, fma(jb[i], r1, acc[1 + i]));
#pragma unroll
            for (int j = i; j < 4; ++j, ++q) acc[q] = fma(ja[i], ja[j], fma(jb[i], jb[j], acc[q]));
        }
    }
    any = __any_sync(kFull, any);
    warp_allreduce16<double>(acc, hdr->xres, lane);  // leaves the totals in xres[0..15]
    if (lane == 0) hdr->xres[16] = any ? 1.0 : 0.0;
    __syncwarp();
}

// On-device linear initialiser (see pnp_device.cuh::linear_init); result -> hdr->pt, returns validity.
template <int WMODE, int LAYOUT>
__device__ __noinline__ bool pair_linear_init(const KParams& kp, PairHeader* hdr, const float* slot, int n, int lane) {
    const int P = kp.n_pts;
    const Camera<float> cf = load_camera<float>(kp, hdr->obj);
    double x[4];
    const bool ok = linear_init_impl<WMODE, LAYOUT>(slot, slot + 3 * P, slot + 5 * P, P, n, lane, cf, reinterpret_cast<float*>(hdr->xres), x);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) hdr->pt[i] = ok ? (double)(float)x[i] : 0.0;  // fp32 hand-over; .py:119-125
    }
    __syncwarp();
    return ok;
}

// ------------------------------------------------------------------ warp B: fp32 Jacobian pass
template <int WMODE, int LAYOUT>
__device__ __forceinline__ void pair_pass_B(const float* __restrict__ s3, const float* __restrict__ s2,
                                            const float* __restrict__ sw, int P, int n, int lane,
                                            const Camera<float>& camf, PairHeader* hdr) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    const float snf = (float)hdr->sn, csf = (float)hdr->cs;
    const float txf = (float)hdr->pt[1], tyf = (float)hdr->pt[2], tzf = (float)hdr->pt[3];
    // clip detection with a safety margin far above the fp32 projection's rounding error (~1e-4 px)
    const float zlo = camf.z_min * 1.001f + 1e-3f;
    const float ulo = camf.u_min + 0.05f, uhi = camf.u_max - 0.05f, vlo = camf.v_min + 0.05f, vhi = camf.v_max - 0.05f;
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 0.f;
    bool flagged = false;
    const int rows = (n + 31) >> 5;
    constexpr int R = 2;  // rows carried through the dependent chain together (see pair_pass_A)
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
        float qxf[R], qzf[R], xnf[R], ynf[R], izf[R], puf[R], pvf[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            qxf[r] = fmaf(csf, Xf[r], snf * Zf[r]);
            qzf[r] = fmaf(csf, Zf[r], -snf * Xf[r]);
            const float zcf = qzf[r] + tzf;
            flagged = flagged || !(zcf >= zlo);
            izf[r] = fast_rcp(zcf);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            xnf[r] = (qxf[r] + txf) * izf[r];
            ynf[r] = (Yf[r] + tyf) * izf[r];
            puf[r] = fmaf(camf.fx, xnf[r], camf.cx);
            pvf[r] = fmaf(camf.fy, ynf[r], camf.cy);
            flagged = flagged || !(puf[r] >= ulo) || !(puf[r] <= uhi) || !(pvf[r] >= vlo) || !(pvf[r] <= vhi);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float au = camf.fx * izf[r], av = camf.fy * izf[r];
            const float bu = -au * xnf[r], bv = -av * ynf[r];
            const float ju0 = fmaf(au, qzf[r], -bu * qxf[r]), jv0 = -bv * qxf[r];
            const float duf = puf[r] - uf[r], dvf = pvf[r] - vf[r];
            if (WMODE != MRPNP_W_FULL) {
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
    flagged = __any_sync(kFull, flagged);
    // transposed reduction: lane L ends with the total of value L>>1, written straight to the header
#pragma unroll
    for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? a[i] : a[i + half];
            const float keep = up ? a[i + half] : a[i];
            a[i] = keep + __shfl_xor_sync(kFull, send, bit);
        }
    }
    a[0] += __shfl_xor_sync(kFull, a[0], 1);
    hdr->sums[lane >> 1] = a[0];
    if (lane == 0) hdr->flagged = flagged ? 1 : 0;
}

// ------------------------------------------------------------------ warp A: fp64 residual / cost pass
// The chain load -> convert -> rotate -> reciprocal -> project -> residual is ~14 dependent fp64 operations
// per point; one row at a time the warp issues an instruction every ~8 cycles (measured: 450 cycles per row).
// kRowsA rows are therefore carried through the chain together, stage by stage, so that their independent
// instructions interleave.
constexpr int kRowsA = 3;
template <int WMODE, int LAYOUT>
__device__ __forceinline__ double pair_pass_A(const float* __restrict__ s3, const float* __restrict__ s2,
                                              const float* __restrict__ sw, int P, int n, int lane, double sn,
                                              double cs, double tx, double ty, double tz, double fx, double fy,
                                              double cx, double cy) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    double cost[kRowsA];
#pragma unroll
    for (int r = 0; r < kRowsA; ++r) cost[r] = 0.0;
    const int rows = (n + 31) >> 5;
    for (int k0 = 0; k0 < rows; k0 += kRowsA) {
        float Xf[kRowsA], Yf[kRowsA], Zf[kRowsA], uf[kRowsA], vf[kRowsA], w0[kRowsA], w1[kRowsA], w2[kRowsA];
#pragma unroll
        for (int r = 0; r < kRowsA; ++r) {
            const int pr = (k0 + r) * 32 + lane;
            const bool valid = pr < n;
            const int p = valid ? pr : n - 1;
            Xf[r] = s3[sidx<LAYOUT, 3>(p, 0, P)]; Yf[r] = s3[sidx<LAYOUT, 3>(p, 1, P)]; Zf[r] = s3[sidx<LAYOUT, 3>(p, 2, P)];
            uf[r] = s2[sidx<LAYOUT, 2>(p, 0, P)]; vf[r] = s2[sidx<LAYOUT, 2>(p, 1, P)];
            w0[r] = sw[sidx<LAYOUT, WC>(p, 0, P)]; w1[r] = sw[sidx<LAYOUT, WC>(p, 1, P)];
            w2[r] = (WMODE == MRPNP_W_FULL) ? sw[sidx<LAYOUT, WC>(p, WC - 1, P)] : 0.f;
            if (!valid) { w0[r] = 0.f; w1[r] = 0.f; w2[r] = 0.f; }  // padding lanes / rows contribute nothing
        }
        double xc[kRowsA], yc[kRowsA], zc[kRowsA], iz[kRowsA];
#pragma unroll
        for (int r = 0; r < kRowsA; ++r) {
            const double X = (double)Xf[r], Y = (double)Yf[r], Z = (double)Zf[r];
            xc[r] = fma(cs, X, fma(sn, Z, tx));
            zc[r] = fma(cs, Z, fma(-sn, X, tz));
            yc[r] = Y + ty;
        }
#pragma unroll
        for (int r = 0; r < kRowsA; ++r) iz[r] = fast_rcp(zc[r]);
#pragma unroll
        for (int r = 0; r < kRowsA; ++r) {
            const double du = fma(fx, xc[r] * iz[r], cx) - (double)uf[r];
            const double dv = fma(fy, yc[r] * iz[r], cy) - (double)vf[r];
            if (WMODE != MRPNP_W_FULL) {
                const double ru = (double)w0[r] * du, rv = (double)w1[r] * dv;
                cost[r] = fma(ru, ru, fma(rv, rv, cost[r]));
            } else {
                const double r0 = fma((double)w0[r], du, (double)w1[r] * dv), r1 = fma((double)w1[r], du, (double)w2[r] * dv);
                cost[r] = fma(r0, r0, fma(r1, r1, cost[r]));
            }
        }
    }
    double cost2 = cost[0];
#pragma unroll
    for (int r = 1; r < kRowsA; ++r) cost2 += cost[r];
    return warp_sum(cost2);
}

// ------------------------------------------------------------------ the kernel
template <int WMODE, int LAYOUT>
__global__ void __launch_bounds__(kPairThreads, 1) pnp_lm_pair_kernel(const __grid_constant__ KParams kp) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    constexpr int CV = WC - 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npairs = blockDim.x >> 6;
    const bool isA = warp < npairs;                 // A warps 0..np-1 and B warps np..2np-1: both roles on every scheduler
    const int pair = isA ? warp : warp - npairs;
    PairHeader* hdr = reinterpret_cast<PairHeader*>(smem_raw + (size_t)pair * kPairHeaderBytes);
    float* slot = reinterpret_cast<float*>(smem_raw + (size_t)npairs * kPairHeaderBytes) + (size_t)pair * kp.slot_floats;
    const int barid = 1 + pair;
    const int P = kp.n_pts;
    float* s3 = slot;
    float* s2 = slot + 3 * P;
    float* sw = slot + 5 * P;
    const int rows_all = (P + 31) >> 5;
    const int max_iter = kp.max_iter > 0 ? kp.max_iter : (kp.max_iter < 0 ? 0 : 50);

    if (isA && lane == 0) {
        mbar_init(&hdr->bar, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    pair_sync(barid);
    uint32_t parity = 0;

    while (true) {
        // ---- S0: next object ----
        if (isA && lane == 0) hdr->obj = atomicAdd(kp.counters, 1);
        pair_sync(barid);
        const int obj = hdr->obj;
        if (obj >= kp.n_obj) break;
        MR_TRACE(0);

        // ---- stage the slab (A issues, both wait on the mbarrier) ----
        if (kp.use_tma) {
            if (isA && lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(&hdr->bar, (uint32_t)((5 + WC) * P * sizeof(float)));
                bulk_g2s(s3, kp.c3d + (size_t)obj * 3 * P, (uint32_t)(3 * P * sizeof(float)), &hdr->bar);
                bulk_g2s(s2, kp.c2d + (size_t)obj * 2 * P, (uint32_t)(2 * P * sizeof(float)), &hdr->bar);
                bulk_g2s(sw, kp.wgt + (size_t)obj * WC * P, (uint32_t)(WC * P * sizeof(float)), &hdr->bar);
            }
            mbar_wait(&hdr->bar, parity);
            parity ^= 1u;
        } else {
            const float* g3 = kp.c3d + (size_t)obj * 3 * P;
            const float* g2 = kp.c2d + (size_t)obj * 2 * P;
            const float* gw = kp.wgt + (size_t)obj * WC * P;
            const int t = (isA ? 0 : 32) + lane;
            for (int i = t; i < 3 * P; i += 64) s3[i] = __ldg(g3 + i);
            for (int i = t; i < 2 * P; i += 64) s2[i] = __ldg(g2 + i);
            for (int i = t; i < WC * P; i += 64) sw[i] = __ldg(gw + i);
            pair_sync(barid);
        }
        const Camera<float> camf = load_camera<float>(kp, obj);
        MR_TRACE(1);

        // ---- sweep A: weights -> istd in place + per-axis sums; A takes even rows, B odd rows ----
        {
            const float inv_scale = 1.f / kp.std_scale;
            float su = 0.f, sv = 0.f;
#pragma unroll 4
            for (int k = isA ? 0 : 1; k < rows_all; k += 2) {
                const int p = k * 32 + lane;
                if (p < P) {
                    float wu = sw[sidx<LAYOUT, WC>(p, 0, P)], wv = sw[sidx<LAYOUT, WC>(p, CV, P)];
                    if (WMODE == MRPNP_W_LOGSTD) {  // uncert_prop_pnp_optimizer.py:73
                        wu = __expf(-wu) * inv_scale;
                        wv = __expf(-wv) * inv_scale;
                        sw[sidx<LAYOUT, WC>(p, 0, P)] = wu;
                        sw[sidx<LAYOUT, WC>(p, CV, P)] = wv;
                    }
                    su += wu;
                    sv += wv;
                }
            }
            su = warp_sum(su);
            sv = warp_sum(sv);
            if (lane == 0) { hdr->wsum[isA ? 0 : 2] = su; hdr->wsum[isA ? 1 : 3] = sv; }
        }
        pair_sync(barid);  // S1
        MR_TRACE(2);
        // ---- inlier ballots of this warp's rows (pnp_uncert_cpu.py:164-168) ----
        {
            const float invP = 1.f / (float)P;
            const float thr_u = kp.istd_thres * ((hdr->wsum[0] + hdr->wsum[2]) * invP);
            const float thr_v = kp.istd_thres * ((hdr->wsum[1] + hdr->wsum[3]) * invP);
            const bool test = kp.istd_thres > 0.f;
#pragma unroll 4
            for (int k = isA ? 0 : 1; k < rows_all; k += 2) {
                const int p = k * 32 + lane;
                bool inl = p < P;
                if (kp.inl_in) {
                    const uint32_t word = __ldg(kp.inl_in + (size_t)obj * rows_all + k);
                    inl = inl && ((word >> lane) & 1u);
                } else if (test && inl) {
                    inl = (sw[sidx<LAYOUT, WC>(p, 0, P)] >= thr_u) && (sw[sidx<LAYOUT, WC>(p, CV, P)] >= thr_v);
                }
                const unsigned m = __ballot_sync(kFull, inl);
                if (lane == 0) hdr->masks[k] = m;
            }
        }
        pair_sync(barid);  // S2
        MR_TRACE(3);
        // ---- count, "<= 4 inliers -> all points" (pnp_uncert_cpu.py:23-32), packed inlier_out ----
        int n_inliers = 0;
        {
            uint32_t my = lane < rows_all ? hdr->masks[lane] : 0u;
            n_inliers = __reduce_add_sync(kFull, __popc(my));
            if (n_inliers <= 4) {
                n_inliers = P;
                const int rem = P - lane * 32;
                my = lane < rows_all ? (rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u)) : 0u;
                pair_sync(barid);  // both warps have read the old masks
                if (isA && lane < rows_all) hdr->masks[lane] = my;
                pair_sync(barid);
            }
            if (isA && kp.inl_out && lane < rows_all) kp.inl_out[(size_t)obj * rows_all + lane] = my;
        }
        // ---- order-preserving in-place compaction (pnp_uncert_cpu.py:24-27,62-66): A moves coords_3d, B moves
        //      coords_2d and the weights; the channels are disjoint, so the two warps never touch the same words ----
        const bool compacted = kp.inlier_opt_only != 0 && n_inliers < P;
        if (compacted) {
            // exclusive prefix of the row counts -> every row's destination base, so rows become independent
            const uint32_t mine = lane < rows_all ? hdr->masks[lane] : 0u;
            int incl = __popc(mine);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += t;
            }
            const int excl = incl - __popc(mine);
            constexpr int RC = 4;  // rows per batch: all loads, one __syncwarp, all stores
            for (int k0 = 0; k0 < rows_all; k0 += RC) {
                unsigned m[RC];
                int d[RC];
                bool inl[RC];
#pragma unroll
                for (int r = 0; r < RC; ++r) {
                    const int k = k0 + r;
                    m[r] = __shfl_sync(kFull, mine, k & 31);
                    const int base = __shfl_sync(kFull, excl, k & 31);
                    inl[r] = (k < rows_all) && ((m[r] >> lane) & 1u);
                    d[r] = base + __popc(m[r] & ((1u << lane) - 1u));
                }
                if (isA) {
                    float v0[RC], v1[RC], v2[RC];
#pragma unroll
                    for (int r = 0; r < RC; ++r) {
                        const int p = (k0 + r) * 32 + lane;
                        if (inl[r]) { v0[r] = s3[sidx<LAYOUT, 3>(p, 0, P)]; v1[r] = s3[sidx<LAYOUT, 3>(p, 1, P)]; v2[r] = s3[sidx<LAYOUT, 3>(p, 2, P)]; }
                    }
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < RC; ++r)
                        if (inl[r]) { s3[sidx<LAYOUT, 3>(d[r], 0, P)] = v0[r]; s3[sidx<LAYOUT, 3>(d[r], 1, P)] = v1[r]; s3[sidx<LAYOUT, 3>(d[r], 2, P)] = v2[r]; }
                } else {
                    float v0[RC], v1[RC], w[RC][WC];
#pragma unroll
                    for (int r = 0; r < RC; ++r) {
                        const int p = (k0 + r) * 32 + lane;
                        if (inl[r]) {
                            v0[r] = s2[sidx<LAYOUT, 2>(p, 0, P)]; v1[r] = s2[sidx<LAYOUT, 2>(p, 1, P)];
#pragma unroll
                            for (int c = 0; c < WC; ++c) w[r][c] = sw[sidx<LAYOUT, WC>(p, c, P)];
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < RC; ++r)
                        if (inl[r]) {
                            s2[sidx<LAYOUT, 2>(d[r], 0, P)] = v0[r]; s2[sidx<LAYOUT, 2>(d[r], 1, P)] = v1[r];
#pragma unroll
                            for (int c = 0; c < WC; ++c) sw[sidx<LAYOUT, WC>(d[r], c, P)] = w[r][c];
                        }
                }
                __syncwarp();
            }
        }
        const int n = compacted ? n_inliers : P;
        pair_sync(barid);  // S3: slot is in its final form
        MR_TRACE(4);

        if (!isA) {
            // =============================== warp B ===============================
            while (true) {
                pair_sync(barid);  // BAR1: A published pt / sn / cs / go
                if (!hdr->go) break;
                pair_pass_B<WMODE, LAYOUT>(s3, s2, sw, P, n, lane, camf, hdr);
                pair_sync(barid);  // BAR2: sums are in the header
            }
            continue;
        }

        // =============================== warp A ===============================
        const double fx = camf.fx, fy = camf.fy, cx = camf.cx, cy = camf.cy;
        bool init_ok = true;
        if (kp.init_mode == MRPNP_INIT_GIVEN) {
            if (lane < 4) hdr->pt[lane] = (double)__ldg(kp.init + (size_t)obj * 4 + lane);
        } else {
            init_ok = pair_linear_init<WMODE, LAYOUT>(kp, hdr, slot, n, lane);
        }
        if (lane < 4) hdr->x[lane] = 0.0;
        __syncwarp();
        if (lane < 4) hdr->x[lane] = hdr->pt[lane];
        double cost = 0.0, radius = kInitialRadius, decrease_factor = 2.0, x_norm = 0.0, model_change = 1.0;
        double step2 = 0.0, gmax = 0.0;
        int term = kNoConvergence, iteration = 0, cost_evals = 0, num_invalid = 0;
        bool reuse_diagonal = false, step_ok = true, clip_x = false, first = true;

        while (true) {
            // ---- publish the evaluation point and run the pass ----
            if (cost_evals > 0 && cost_evals < 7) MR_TRACE(4 + 4 * cost_evals);
            double sn, cs;
            sincos(hdr->pt[0], &sn, &cs);
            const double tx = hdr->pt[1], ty = hdr->pt[2], tz = hdr->pt[3];
            if (lane == 0) { hdr->sn = sn; hdr->cs = cs; hdr->go = 1; }
            pair_sync(barid);  // BAR1
            if (cost_evals < 6) MR_TRACE(5 + 4 * cost_evals);
            const double cost2 = pair_pass_A<WMODE, LAYOUT>(s3, s2, sw, P, n, lane, sn, cs, tx, ty, tz, fx, fy, cx, cy);
            if (cost_evals < 6) MR_TRACE(6 + 4 * cost_evals);
            pair_sync(barid);  // BAR2
            if (cost_evals < 6) MR_TRACE(7 + 4 * cost_evals);
            double acc[15];  // |r|^2, g[4], upper-tri(J^T J)[10] at pt
            acc[0] = cost2;
#pragma unroll
            for (int i = 0; i < 14; ++i) acc[1 + i] = (double)hdr->sums[i];
            bool clip_p = false;
            if (hdr->flagged) {  // a point near a clip bound: redo the pass with exact fp64 clip semantics
                pair_exact_pass<WMODE, LAYOUT>(kp, hdr, slot, n, lane, 0, 0);
#pragma unroll
                for (int i = 0; i < 15; ++i) acc[i] = hdr->xres[i];
                clip_p = hdr->xres[16] != 0.0;
                __syncwarp();
            }
            ++cost_evals;
            const bool cfinite = finite_value(acc[0]);
            double asum = 0.0;
#pragma unroll
            for (int i = 1; i < 15; ++i) asum += fabs(acc[i]);
            const bool jfinite = cfinite && finite_value(asum);
            bool accept = false, last = false;
            if (first) {  // IterationZero
                first = false;
                if (!jfinite || !init_ok) { term = kFailure; break; }
                accept = true;
                cost = 0.5 * acc[0];
                if (lane < 4) {  // jacobi_scaling from the initial Jacobian only
                    double d = 0.0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) d = (lane == i) ? acc[5 + tri(i, i)] : d;
                    hdr->scale[lane] = fast_rcp(1.0 + fast_sqrt(d));
                }
                __syncwarp();
            } else {
                const double cand_cost = cfinite ? 0.5 * acc[0] : kDblMax;
                const double ptol = kParameterTol * (x_norm + kParameterTol);
                if (step2 <= ptol * ptol) { term = kConvergence; break; }        // ParameterToleranceReached
                const double cost_change = cost - cand_cost;
                if (fabs(cost_change) <= kFunctionTol * cost) {                  // FunctionToleranceReached
                    term = kConvergence;
                    if (!(kp.adopt_ftol && cand_cost < cost)) break;            // Ceres 1.14: candidate dropped
                    last = true;
                }
                const double rho = cost_change * fast_rcp(model_change);
                if (last || rho > kMinRelDecrease) {  // HandleSuccessfulStep
                    if (!jfinite) { term = kFailure; break; }
                    accept = true;
                    cost = cand_cost;
                    const double q = 2.0 * rho - 1.0;
                    radius = fmin(kMaxRadius, radius * fast_rcp(fmax(1.0 / 3.0, 1.0 - q * q * q)));
                    decrease_factor = 2.0;
                    reuse_diagonal = false;
                } else {  // HandleUnsuccessfulStep
                    radius /= decrease_factor;
                    decrease_factor *= 2.0;
                }
            }
            if (accept) {  // x <- pt, scaled gradient and Gauss-Newton matrix at x
                const double s0 = hdr->scale[0], s1 = hdr->scale[1], s2s = hdr->scale[2], s3s = hdr->scale[3];
                const double sc[4] = {s0, s1, s2s, s3s};
                gmax = fmax(fmax(fabs(acc[1]), fabs(acc[2])), fmax(fabs(acc[3]), fabs(acc[4])));
                const double p0 = hdr->pt[0], p1 = hdr->pt[1], p2 = hdr->pt[2], p3 = hdr->pt[3];
                x_norm = fast_sqrt(p0 * p0 + p1 * p1 + p2 * p2 + p3 * p3);
                if (lane == 0) {
                    hdr->x[0] = p0; hdr->x[1] = p1; hdr->x[2] = p2; hdr->x[3] = p3;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        hdr->gs[i] = acc[1 + i] * sc[i];
#pragma unroll
                        for (int j = i; j < 4; ++j) hdr->Hs[tri(i, j)] = acc[5 + tri(i, j)] * (sc[i] * sc[j]);
                    }
                }
                clip_x = clip_p;
                step_ok = true;
                __syncwarp();
            }
            if (last) break;
            // ---- next trust-region step (invalid steps shrink the radius without a new evaluation) ----
            bool stop = false;
            while (true) {
                // FinalizeIterationAndCheckIfMinimizerCanContinue
                if (iteration >= max_iter) { term = kNoConvergence; stop = true; break; }
                if (step_ok && gmax <= kGradientTol) { term = kConvergence; stop = true; break; }
                if (radius <= kMinRadius) { term = kConvergence; stop = true; break; }
                ++iteration;
                step_ok = false;
                // LevenbergMarquardtStrategy::ComputeStep on the column-scaled system
                double A[10], gsr[4], D[4], y[4];
#pragma unroll
                for (int i = 0; i < 10; ++i) A[i] = hdr->Hs[i];
                if (!reuse_diagonal) {
                    if (lane < 4) {
                        double d = 0.0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) d = (lane == i) ? A[tri(i, i)] : d;
                        hdr->diag[lane] = fmin(fmax(d, kMinLmDiag), kMaxLmDiag);
                    }
                    __syncwarp();
                }
                reuse_diagonal = true;
                const double inv_radius = fast_rcp(radius);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    D[i] = hdr->diag[i] * inv_radius;
                    A[tri(i, i)] += D[i];
                    gsr[i] = hdr->gs[i];
                }
                const Ldl4 f = ldl4_factor(A);
                bool valid = f.ok;
                if (valid) {
                    ldl4_solve(f, gsr, y);  // step = -y
                    // model_cost_change = y^T gs - 1/2 y^T Hs y, and (Hs + D) y = gs  =>  1/2 (y^T gs + sum D_i y_i^2)
                    double yg = 0.0, ydy = 0.0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) { yg = fma(y[i], gsr[i], yg); ydy = fma(D[i] * y[i], y[i], ydy); }
                    model_change = 0.5 * (yg + ydy);
                    valid = (model_change > 0.0) && finite_value(fabs(y[0]) + fabs(y[1]) + fabs(y[2]) + fabs(y[3]));
                }
                if (valid) {
                    num_invalid = 0;
                    step2 = 0.0;
                    double mine = 0.0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double d = -y[i] * hdr->scale[i];
                        step2 = fma(d, d, step2);
                        mine = (lane == i) ? hdr->x[i] + d : mine;
                    }
                    if (lane < 4) hdr->pt[lane] = mine;
                    __syncwarp();
                    break;
                }
                // HandleInvalidStep
                if (++num_invalid >= kMaxInvalidSteps) { term = kFailure; stop = true; break; }
                radius /= decrease_factor;
                decrease_factor *= 2.0;
            }
            if (stop) break;
        }
        MR_TRACE(29);
        // release warp B
        if (lane == 0) hdr->go = 0;
        pair_sync(barid);  // BAR1 (final)
        bool usable = term != kFailure;  // Summary::IsSolutionUsable (pnp_uncert_cpu.cpp:276)

        // ---------------- pose covariance -> xres[0..15] ----------------
        {
            double cov[16];
            bool have = false;
            if (kp.cov_mode != MRPNP_COV_NONE && usable) {
                // Hs holds the scaled J^T J at the returned x with Ceres masks: cov = S Hs^-1 S.  The pipeline
                // covariance (hessian.py:67-87) differs only when a point is clipped at x or outliers were kept.
                const bool need_pass = kp.cov_mode == MRPNP_COV_PIPELINE && (clip_x || (!compacted && n_inliers < P));
                double Hm[10], sc[4];
                if (need_pass) {
                    if (lane < 4) hdr->pt[lane] = hdr->x[lane];
                    __syncwarp();
                    pair_exact_pass<WMODE, LAYOUT>(kp, hdr, slot, n, lane, 1, compacted ? 0 : 1);
#pragma unroll
                    for (int i = 0; i < 10; ++i) Hm[i] = hdr->xres[5 + i];
#pragma unroll
                    for (int i = 0; i < 4; ++i) sc[i] = 1.0;
                    __syncwarp();
                } else {
#pragma unroll
                    for (int i = 0; i < 10; ++i) Hm[i] = hdr->Hs[i];
#pragma unroll
                    for (int i = 0; i < 4; ++i) sc[i] = hdr->scale[i];
                }
                have = spd_inverse4(Hm, cov);
                if (have) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) cov[i * 4 + j] *= sc[i] * sc[j];
                } else {
                    usable = false;  // pnp_uncert.py:79-85 fallback: H := I, object invalid
                }
            }
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) hdr->xres[i] = have ? cov[i] : ((i % 5 == 0) ? 1.0 : 0.0);
            }
            __syncwarp();
        }
        // ---------------- result row: one coalesced 96-byte store ----------------
        {
            float v = 0.f;
            if (lane < 4) v = (float)hdr->x[lane];
            else if (lane < 20) v = (float)hdr->xres[lane - 4];
            v = (lane == 20) ? (usable ? 1.f : 0.f) : v;
            v = (lane == 21) ? (float)iteration : v;
            v = (lane == 22) ? (float)cost : v;
            v = (lane == 23) ? (float)radius : v;
            if (lane < MRPNP_RESULT_STRIDE) kp.result[(size_t)obj * MRPNP_RESULT_STRIDE + lane] = v;
#ifndef MRPNP_TRACE
            if (kp.result64) {
                double d = 0.0;
                if (lane < 4) d = hdr->x[lane];
                d = (lane == 4) ? cost : d;
                d = (lane == 5) ? radius : d;
                d = (lane == 6) ? (double)cost_evals : d;
                d = (lane == 7) ? (double)term : d;
                if (lane < 8) kp.result64[(size_t)obj * 8 + lane] = d;
            }
#endif
        }
        MR_TRACE(30);
        __syncwarp();
    }

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
