// pnp_fast.cuh -- MRPNP_PREC_FAST passes: candidate evaluations of the LM loop in fp32 with incrementally tracked
// residuals (device functions of pnp_kernel_fast.cuh).
//
// Why: Ceres' accept / function-tolerance tests (TrustRegionMinimizer, defaults of pnp_uncert_cpu.cpp:270-271) read
// cost(x) - cost(x + delta) and stop at |change| <= 1e-6 cost, so that difference has to be right to ~1e-4 of ITSELF.
// Recomputing residuals in fp32 at every point cannot deliver that (1e-4 px rounding on ~1 px residuals), which is why
// MRPNP_PREC_MIXED keeps the whole residual chain in fp64.  Here the residuals are evaluated ONCE per object in fp64
// (at the point reached by the first step, eval_pass_first) and kept in shared memory in place of the observations;
// every later evaluation computes only the CHANGE of each residual between the accepted point x and the candidate x + delta, from
// formulas in which every operand is small when the step is small:
//
//     q'  = R_y(yaw') X,  x' = q' + t'                      (candidate camera-frame point, plain fp32)
//     Dq  = q' - R_y(-dyaw) q' = (-(cos dyaw - 1) q'_x + sin dyaw q'_z , -sin dyaw q'_x - (cos dyaw - 1) q'_z)
//     Dx  = Dq_x + dt_x,  Dz = Dq_z + dt_z,  Dy = dt_y      (motion of the point, exact to fp32 RELATIVE precision)
//     D(x/z) = (Dx - (x'/z') Dz) / z_old,  z_old = z' - Dz  (change of the normalised projection, same for y)
//     dr  = w f D(x/z),   r' = r + dr,   cost' - cost = 1/2 sum dr (r + r')
//
// so cost differences carry a relative error of ~1e-6 however small the step, and the stored residuals pick up one
// fp32 rounding (~6e-8 of |r|) per accepted step.  No fp64 instruction and no conversion is left in the pass; the
// Jacobian, J^T r and J^T J at the candidate come out of the same quantities.  The stored residuals are updated
// speculatively in place (most steps are accepted); a rejected step is rolled back by undo_pass_delta.
//
// Objects for which any point comes near a clip bound (z_min or the u/v ranges of pnp_uncert_cpu.cpp:36-42) are not
// handled here: the kernel appends them to a redo list that the exact kernel (MRPNP_PREC_MIXED with its fp64 cold
// path) processes afterwards, so clip semantics stay those of the reference.
#pragma once
#include "pnp_device.cuh"

namespace mrpnp {

// Per-evaluation constants of the delta pass (all derived from fp64 scalars, then rounded once).
struct DeltaStep {
    float cp, sp, txp, typ, tzp;   // candidate pose: cos/sin yaw', t'
    float ncdm1, sd;               // -(cos dyaw - 1), sin dyaw          (dyaw = yaw' - yaw)
    float dtx, dty, dtz;           // t' - t
};

// Which rows of the slot a warp works on.  RowMap<1>: one warp owns all n compacted points.  RowMap<2>: two warps
// share an object; each compacted its own half of the rows, so the inliers sit in two segments [0, n0) and
// [base1, base1 + n1) whose rows of 32 points are numbered through and dealt alternately to the two warps (every
// warp always revisits the same points: it owns their tracked residuals; the load is balanced to one row).
template <int TEAM> struct RowMap;
template <> struct RowMap<1> {
    int n;
    __device__ __forceinline__ int groups(int R) const { return (((n + 31) >> 5) + R - 1) / R; }
    __device__ __forceinline__ int point(int g, int r, int R, int lane, bool& valid) const {
        const int pr = (g * R + r) * 32 + lane;
        valid = pr < n;
        return valid ? pr : 0;
    }
};
template <> struct RowMap<2> {
    int n0, n1, base1, rows0, rows_total, t, safe;
    __device__ __forceinline__ RowMap(int n0_, int n1_, int base1_, int t_)
        : n0(n0_), n1(n1_), base1(base1_), rows0((n0_ + 31) >> 5), rows_total(((n0_ + 31) >> 5) + ((n1_ + 31) >> 5)),
          t(t_), safe(n0_ > 0 ? 0 : base1_) {}
    __device__ __forceinline__ int groups(int R) const {
        const int mine = rows_total > t ? (rows_total - t + 1) >> 1 : 0;
        return (mine + R - 1) / R;
    }
    __device__ __forceinline__ int point(int g, int r, int R, int lane, bool& valid) const {
        const int v = (g * R + r) * 2 + t;
        const bool s1 = v >= rows0;
        const int k = (s1 ? v - rows0 : v) * 32 + lane;
        valid = (v < rows_total) && (k < (s1 ? n1 : n0));
        return valid ? (s1 ? base1 + k : k) : safe;
    }
};

// Clip window in normalised coordinates with the safety margins of eval_pass_mixed (0.05 px, z_min * 1.001 + 1e-3).
struct ClipWindow {
    float xmid, xhalf, ymid, yhalf, zlo;
};
__device__ __forceinline__ ClipWindow make_clip_window(const Camera<float>& c) {
    ClipWindow w;
    const float ifx = fast_rcp(c.fx), ify = fast_rcp(c.fy);
    const float xlo = (c.u_min + 0.05f - c.cx) * ifx, xhi = (c.u_max - 0.05f - c.cx) * ifx;
    const float ylo = (c.v_min + 0.05f - c.cy) * ify, yhi = (c.v_max - 0.05f - c.cy) * ify;
    w.xmid = 0.5f * (xlo + xhi); w.xhalf = 0.5f * (xhi - xlo);
    w.ymid = 0.5f * (ylo + yhi); w.yhalf = 0.5f * (yhi - ylo);
    w.zlo = c.z_min * 1.001f + 1e-3f;
    return w;
}

// Transposed reduction of 16 per-lane partial sums, WITHOUT the broadcast: lane L ends with the warp total of value
// L >> 1 (the two lanes of a pair hold the same number).  8+4+2+1+1 = 16 shuffles.
__device__ __forceinline__ float warp_reduce16_scatter(float v[16], int lane) {
#pragma unroll
    for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, bit);
        }
    }
    return v[0] + __shfl_xor_sync(kFull, v[0], 1);
}

// ------------------------------------------------------------------ evaluations from the observations
// The two evaluations that read the observations (u, v) instead of tracked residuals:
//   anchor == false  the very first one, at the initial point: plain fp32 (residuals of many pixels there: the 1e-4 px
//                    rounding of an fp32 projection is irrelevant, and the decision on the first step is never close);
//   anchor == true   the second one, at the point after the first step: the residual chain runs in fp64 (as in
//                    eval_pass_mixed) and the residuals REPLACE the observations in the slot -- the delta passes track
//                    them from here on.  Diagonal weights store r = w d and fold the focal lengths into the weights
//                    (w_u fx, w_v fy); full weights store the pixel differences d themselves.
// Anchoring after the first (large) step instead of at the initial point matters: a delta pass leaves ~1e-7 of the
// residual CHANGE behind as a fixed error of the tracked residuals, and only the first step changes them by many pixels.
// Out: per-lane partial sums a[0..13] (J^T r, J^T J; layout of eval_pass_fp64 minus the cost), a[14] = this lane's
// share of sum |r|^2, a[15] = 0; flagged = some point is within the margin of a clip bound.
#ifndef MRPNP_FIRST_R
#define MRPNP_FIRST_R 2
#endif
template <int WMODE, int LAYOUT, int R = MRPNP_FIRST_R, class ROWS = RowMap<1>>
__device__ __forceinline__ void eval_pass_first(const float* s3, float* s2, float* sw, int P, const ROWS rows, int lane,
                                                bool anchor, const float x[4], float snf, float csf,
                                                const Camera<float>& camf, float a[16], bool& flagged) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    const float txf = x[1], tyf = x[2], tzf = x[3];
    const double sn = (double)snf, cs = (double)csf, tx = (double)txf, ty = (double)tyf, tz = (double)tzf;
    const double fx = (double)camf.fx, fy = (double)camf.fy, cx = (double)camf.cx, cy = (double)camf.cy;
    const float zlo = camf.z_min * 1.001f + 1e-3f;
    const float ulo = camf.u_min + 0.05f, uhi = camf.u_max - 0.05f, vlo = camf.v_min + 0.05f, vhi = camf.v_max - 0.05f;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 0.f;
    float margin = 1e30f;
    const int ngroups = rows.groups(R);
#pragma unroll 1
    for (int g = 0; g < ngroups; ++g) {
        float Xf[R], Yf[R], Zf[R], uf[R], vf[R], w0[R], w1[R], w2[R];
        bool valid[R];
        int pidx[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = rows.point(g, r, R, lane, valid[r]);
            pidx[r] = p;
            // padding lanes touch no memory (another lane may be rewriting the slot): a point at the object's origin
            // with zero weights contributes nothing
            Xf[r] = Yf[r] = Zf[r] = uf[r] = vf[r] = w0[r] = w1[r] = w2[r] = 0.f;
            if (valid[r]) {
                Xf[r] = s3[sidx<LAYOUT, 3>(p, 0, P)]; Yf[r] = s3[sidx<LAYOUT, 3>(p, 1, P)]; Zf[r] = s3[sidx<LAYOUT, 3>(p, 2, P)];
                uf[r] = s2[sidx<LAYOUT, 2>(p, 0, P)]; vf[r] = s2[sidx<LAYOUT, 2>(p, 1, P)];
                w0[r] = sw[sidx<LAYOUT, WC>(p, 0, P)]; w1[r] = sw[sidx<LAYOUT, WC>(p, 1, P)];
                if (WMODE == MRPNP_W_FULL) w2[r] = sw[sidx<LAYOUT, WC>(p, WC - 1, P)];
            }
        }
        // ---- fp32 projection: Jacobian, clip detection, (first evaluation) residuals ----
        float qxf[R], qzf[R], izf[R], xnf[R], ynf[R], euf[R], evf[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            qxf[r] = fmaf(csf, Xf[r], snf * Zf[r]);
            qzf[r] = fmaf(csf, Zf[r], -snf * Xf[r]);
            const float zcf = qzf[r] + tzf;
            margin = fminf(margin, zcf - zlo);
            izf[r] = fast_rcp(zcf);
            xnf[r] = (qxf[r] + txf) * izf[r];
            ynf[r] = (Yf[r] + tyf) * izf[r];
            const float puf = fmaf(camf.fx, xnf[r], camf.cx), pvf = fmaf(camf.fy, ynf[r], camf.cy);
            margin = fminf(margin, fminf(fminf(puf - ulo, uhi - puf), fminf(pvf - vlo, vhi - pvf)));
            euf[r] = puf - uf[r];
            evf[r] = pvf - vf[r];
        }
        if (anchor) {
            // ---- fp64 residual chain: the pixel differences to ~1e-13 px before they are rounded to fp32 ----
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double X = (double)Xf[r], Y = (double)Yf[r], Z = (double)Zf[r];
                const double xc = fma(cs, X, fma(sn, Z, tx));
                const double zc = fma(cs, Z, fma(-sn, X, tz));
                const double yc = Y + ty;
                const double iz = fast_rcp(zc);
                euf[r] = (float)(fma(fx, xc * iz, cx) - (double)uf[r]);
                evf[r] = (float)(fma(fy, yc * iz, cy) - (double)vf[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float au = camf.fx * izf[r], av = camf.fy * izf[r];
            const float bu = -au * xnf[r], bv = -av * ynf[r];
            const float ju0 = fmaf(au, qzf[r], -bu * qxf[r]), jv0 = -bv * qxf[r];
            if (WMODE != MRPNP_W_FULL) {
                const float ruf = w0[r] * euf[r], rvf = w1[r] * evf[r];
                a[14] = fmaf(ruf, ruf, fmaf(rvf, rvf, a[14]));
                if (anchor && valid[r]) {
                    s2[sidx<LAYOUT, 2>(pidx[r], 0, P)] = ruf; s2[sidx<LAYOUT, 2>(pidx[r], 1, P)] = rvf;
                    sw[sidx<LAYOUT, WC>(pidx[r], 0, P)] = w0[r] * camf.fx;
                    sw[sidx<LAYOUT, WC>(pidx[r], 1, P)] = w1[r] * camf.fy;
                }
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
                if (anchor && valid[r]) { s2[sidx<LAYOUT, 2>(pidx[r], 0, P)] = euf[r]; s2[sidx<LAYOUT, 2>(pidx[r], 1, P)] = evf[r]; }
                const float r0f = fmaf(w0[r], euf[r], w1[r] * evf[r]), r1f = fmaf(w1[r], euf[r], w2[r] * evf[r]);
                a[14] = fmaf(r0f, r0f, fmaf(r1f, r1f, a[14]));
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
    flagged = __any_sync(kFull, !(margin >= 0.f));
}

// ------------------------------------------------------------------ candidate evaluation (fp32 delta pass)
// Candidate-frame quantities of one point and the change of its normalised projection since the accepted point.
struct PointDelta {
    float qx, qz, z1, izp, xnp, ynp, Du, Dv;
};
__device__ __forceinline__ PointDelta point_delta(float X, float Y, float Z, const DeltaStep& s) {
    PointDelta d;
    d.qx = fmaf(s.cp, X, s.sp * Z);
    d.qz = fmaf(s.cp, Z, -s.sp * X);
    const float x1 = d.qx + s.txp, y1 = Y + s.typ;
    d.z1 = d.qz + s.tzp;
    d.izp = fast_rcp(d.z1);
    d.xnp = x1 * d.izp;
    d.ynp = y1 * d.izp;
    const float Dx = fmaf(s.ncdm1, d.qx, fmaf(s.sd, d.qz, s.dtx));
    const float Dz = fmaf(s.ncdm1, d.qz, fmaf(-s.sd, d.qx, s.dtz));
    const float izo = fast_rcp(d.z1 - Dz);
    d.Du = fmaf(-d.xnp, Dz, Dx) * izo;
    d.Dv = fmaf(-d.ynp, Dz, s.dty) * izo;
    return d;
}

// In: tracked residuals at the accepted point in the s2 planes.  Out (per-lane partial sums): a[0..13] = J^T r' and
// J^T J at the candidate, a[14] = sum |r'|^2 - sum |r|^2 (twice the cost change), a[15] = 0; the s2 planes now hold r'.
template <int WMODE, int LAYOUT, int R = 2, class ROWS = RowMap<1>>
__device__ __forceinline__ void eval_pass_delta(const float* s3, float* s2, const float* sw, int P, const ROWS rows,
                                                int lane, const DeltaStep& st, const Camera<float>& camf,
                                                const ClipWindow& cw, float a[16], bool& flagged) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 0.f;
    float dc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) dc[r] = 0.f;
    float mx = 0.f, my = 0.f, mz = 1e30f;
    const int ngroups = rows.groups(R);
#pragma unroll 1
    for (int g = 0; g < ngroups; ++g) {
        float X[R], Y[R], Z[R], e0[R], e1[R], w0[R], w1[R], w2[R];
        bool valid[R];
        int pidx[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = rows.point(g, r, R, lane, valid[r]);
            pidx[r] = p;
            // padding lanes touch no memory: a point at the object's origin with zero weights contributes nothing
            X[r] = Y[r] = Z[r] = e0[r] = e1[r] = w0[r] = w1[r] = w2[r] = 0.f;
            if (valid[r]) {
                X[r] = s3[sidx<LAYOUT, 3>(p, 0, P)]; Y[r] = s3[sidx<LAYOUT, 3>(p, 1, P)]; Z[r] = s3[sidx<LAYOUT, 3>(p, 2, P)];
                e0[r] = s2[sidx<LAYOUT, 2>(p, 0, P)]; e1[r] = s2[sidx<LAYOUT, 2>(p, 1, P)];
                w0[r] = sw[sidx<LAYOUT, WC>(p, 0, P)]; w1[r] = sw[sidx<LAYOUT, WC>(p, 1, P)];
                if (WMODE == MRPNP_W_FULL) w2[r] = sw[sidx<LAYOUT, WC>(p, WC - 1, P)];
            }
        }
        PointDelta d[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            d[r] = point_delta(X[r], Y[r], Z[r], st);
            mz = fminf(mz, d[r].z1);
            mx = fmaxf(mx, fabsf(d[r].xnp - cw.xmid));
            my = fmaxf(my, fabsf(d[r].ynp - cw.ymid));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float qx = d[r].qx, qz = d[r].qz, izp = d[r].izp, xnp = d[r].xnp, ynp = d[r].ynp;
            if (WMODE != MRPNP_W_FULL) {
                // tracked r = (w_u fx) (x/z - un), (w_v fy) (y/z - vn); the weights already carry the focal lengths
                const float dru = w0[r] * d[r].Du, drv = w1[r] * d[r].Dv;
                const float ru = e0[r] + dru, rv = e1[r] + drv;
                dc[r] = fmaf(dru, e0[r] + ru, fmaf(drv, e1[r] + rv, dc[r]));
                if (valid[r]) { s2[sidx<LAYOUT, 2>(pidx[r], 0, P)] = ru; s2[sidx<LAYOUT, 2>(pidx[r], 1, P)] = rv; }
                const float a1 = w0[r] * izp, b2 = w1[r] * izp;
                const float a3 = -a1 * xnp, b3 = -b2 * ynp;
                const float a0 = fmaf(a1, qz, -a3 * qx), b0 = -b3 * qx;
                a[0] = fmaf(a0, ru, fmaf(b0, rv, a[0]));
                a[1] = fmaf(a1, ru, a[1]);
                a[2] = fmaf(b2, rv, a[2]);
                a[3] = fmaf(a3, ru, fmaf(b3, rv, a[3]));
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
                // tracked e = pixel differences; r = W e (pnp_uncert_cpu.cpp:214-215)
                const float deu = camf.fx * d[r].Du, dev = camf.fy * d[r].Dv;
                const float eu = e0[r] + deu, ev = e1[r] + dev;
                if (valid[r]) { s2[sidx<LAYOUT, 2>(pidx[r], 0, P)] = eu; s2[sidx<LAYOUT, 2>(pidx[r], 1, P)] = ev; }
                const float dr0 = fmaf(w0[r], deu, w1[r] * dev), dr1 = fmaf(w1[r], deu, w2[r] * dev);
                const float r0 = fmaf(w0[r], eu, w1[r] * ev), r1 = fmaf(w1[r], eu, w2[r] * ev);
                dc[r] = fmaf(dr0, fmaf(2.f, r0, -dr0), fmaf(dr1, fmaf(2.f, r1, -dr1), dc[r]));
                const float au = camf.fx * izp, av = camf.fy * izp;
                const float bu = -au * xnp, bv = -av * ynp;
                const float ju0 = fmaf(au, qz, -bu * qx), jv0 = -bv * qx;
                const float a0 = fmaf(w0[r], ju0, w1[r] * jv0), a1 = w0[r] * au, a2 = w1[r] * av, a3 = fmaf(w0[r], bu, w1[r] * bv);
                const float b0 = fmaf(w1[r], ju0, w2[r] * jv0), b1 = w1[r] * au, b2 = w2[r] * av, b3 = fmaf(w1[r], bu, w2[r] * bv);
                a[0] = fmaf(a0, r0, fmaf(b0, r1, a[0]));
                a[1] = fmaf(a1, r0, fmaf(b1, r1, a[1]));
                a[2] = fmaf(a2, r0, fmaf(b2, r1, a[2]));
                a[3] = fmaf(a3, r0, fmaf(b3, r1, a[3]));
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
    flagged = __any_sync(kFull, !(mz >= cw.zlo) || !(mx <= cw.xhalf) || !(my <= cw.yhalf));
    // cost change: fp32 per-lane partials and an fp32 cross-lane tree.  The partials cancel (the gradient is ~0 near the
    // optimum), which bounds the relative error of the total by ~1e-7 |r| / |dr| ~ 2e-4 at the function-tolerance
    // threshold -- a band in which ~1e-4 of all decisions fall.
    float dcs = dc[0];
#pragma unroll
    for (int r = 1; r < R; ++r) dcs += dc[r];
    a[14] = dcs;
}

// Roll the speculative residual update of a rejected candidate back: r = r' - dr with dr recomputed from the same
// inputs (at most one fp32 rounding away from the value before the candidate; a perturbation of ~6e-8 |r|).
template <int WMODE, int LAYOUT, class ROWS = RowMap<1>>
__device__ __noinline__ void undo_pass_delta(float* slot, int P, const ROWS rows, int lane, DeltaStep st, float fx, float fy) {
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    const float* s3 = slot;
    float* s2 = slot + 3 * P;
    const float* sw = slot + 5 * P;
    const int ngroups = rows.groups(1);
#pragma unroll 1
    for (int g = 0; g < ngroups; ++g) {
        bool valid;
        const int p = rows.point(g, 0, 1, lane, valid);
        if (!valid) continue;
        const PointDelta d = point_delta(s3[sidx<LAYOUT, 3>(p, 0, P)], s3[sidx<LAYOUT, 3>(p, 1, P)],
                                         s3[sidx<LAYOUT, 3>(p, 2, P)], st);
        float d0, d1;
        if (WMODE != MRPNP_W_FULL) {
            d0 = sw[sidx<LAYOUT, WC>(p, 0, P)] * d.Du;
            d1 = sw[sidx<LAYOUT, WC>(p, 1, P)] * d.Dv;
        } else {
            d0 = fx * d.Du;
            d1 = fy * d.Dv;
        }
        s2[sidx<LAYOUT, 2>(p, 0, P)] -= d0;
        s2[sidx<LAYOUT, 2>(p, 1, P)] -= d1;
    }
}

}  // namespace mrpnp
