// pnp_fast.cuh -- MRPNP_PREC_FAST passes (device functions of pnp_kernel_fast.cuh): the evaluations of the LM loop on
// incrementally TRACKED residuals, two points per lane in packed fp32 (sm_100 FFMA2 / FADD2 / FMUL2).
//
// Why tracked residuals: Ceres' accept / function-tolerance tests (TrustRegionMinimizer, defaults of
// pnp_uncert_cpu.cpp:270-271) read cost(x) - cost(x + delta) and stop at |change| <= 1e-6 cost, so that difference has
// to be right to ~1e-4 of ITSELF.  Recomputing residuals in fp32 at every point cannot deliver that (1e-4 px rounding on
// ~1 px residuals).  Here the residuals are evaluated ONCE per object in fp64 (at the point reached by the first step,
// the "anchor") and kept in shared memory in place of the observations; every later evaluation computes only the CHANGE
// of each residual between the accepted point x and the candidate x + delta, from formulas in which every operand is
// small when the step is small:
//
//     q'  = R_y(yaw') X,  x' = q' + t'                      (candidate camera-frame point, plain fp32)
//     Dq  = q' - R_y(-dyaw) q' = (-(cos dyaw - 1) q'_x + sin dyaw q'_z , -sin dyaw q'_x - (cos dyaw - 1) q'_z)
//     Dx  = Dq_x + dt_x,  Dz = Dq_z + dt_z,  Dy = dt_y      (motion of the point, exact to fp32 RELATIVE precision)
//     D(x/z) = (Dx - (x'/z') Dz) / z_old,  z_old = z' - Dz  (change of the normalised projection, same for y)
//
// Formulation (round 2).  The tracked quantity is the NORMALISED reprojection difference e = (x'/z' - un, y'/z' - vn),
// un = (u - cx) / fx, and each point carries M = F W^T W F (F = diag(fx, fy); W the whitening matrix of
// pnp_uncert_cpu.cpp:214-215, or diag(w_u, w_v) of :44-45), so that with the normalised Jacobian rows
// Ju = (ju, 1, 0, -xn) / z', Jv = (jv, 0, 1, -yn) / z'  (ju = qz + xn qx, jv = yn qx; order yaw, tx, ty, tz)
//
//     cost = 1/2 sum e^T M e        J^T r = sum [Ju; Jv]^T M e        J^T J = sum [Ju; Jv]^T M [Ju; Jv]
//
// -- identical to the whitened sums of the reference, but the unit / zero columns turn 9 of the 24 multiply-adds of
// J^T J and J^T r into plain adds, and W itself is never needed again.  The cost change of a candidate is
// 1/2 sum De^T (2 M e' - M De): relative accuracy ~1e-6 however small the step.
//
// Two points per lane: a lane owns the compacted points 64 g + 2 lane and + 1 of every group g of 64, loads them with
// one 64-bit shared-memory access per plane and runs every arithmetic step once for both in a packed fp32 instruction.
// The compacted arrays are padded to a multiple of 64 with null points (M = 0), so the loops carry no validity
// predicates; when the padding does not fit (more than floor64(P) inliers) the remainder goes through the same code
// instantiated for one point per lane with a validity flag, out of line.
//
// Objects for which a point comes near a clip bound (z_min or the u/v ranges of pnp_uncert_cpu.cpp:36-42) are not
// handled here: the kernel hands them to the exact fp64 routine (solve_object_exact), so clip semantics stay the
// reference's.  The per-point test is skipped when a bounding box of the object is inside the clip window.
#pragma once
#include "pnp_device.cuh"

namespace mrpnp {

// ------------------------------------------------------------------ packed / scalar arithmetic behind one set of names
__device__ __forceinline__ float2 vfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
// Three per-point operands.  In isolation (tools/microbench5.cu) an FFMA2 reading three register PAIRS issues every
// 3.9 cycles per scheduler -- slower than the two scalar FFMAs (2 x 1.56) it replaces -- while a packed instruction with
// a scalar-broadcast operand or only two operands (FMUL2 / FADD2) issues every 2.0 (against 2 x 1.19 scalar).  Inside
// the kernel the packed form still wins by 2-3 % at 8 warps (profiles/r02_ab_variants.txt, c_w8p3 against c_w8): the
// passes are short of instruction-cache and issue slots, not of FMA-pipe cycles.  MRPNP_EXP_SCALAR3 builds the other.
#ifndef MRPNP_EXP_SCALAR3
__device__ __forceinline__ float2 vfma3(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#else
__device__ __forceinline__ float2 vfma3(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
__device__ __forceinline__ float vfma3(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float2 vmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 vneg(float2 a) { return make_float2(-a.x, -a.y); }  // folds into operand modifiers
__device__ __forceinline__ float2 vsub(float2 a, float2 b) { return __fadd2_rn(a, vneg(b)); }
__device__ __forceinline__ float2 vrcp(float2 a) { return make_float2(fast_rcp(a.x), fast_rcp(a.y)); }
__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float vmul(float a, float b) { return a * b; }
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float vneg(float a) { return -a; }
__device__ __forceinline__ float vsub(float a, float b) { return a - b; }
__device__ __forceinline__ float vrcp(float a) { return fast_rcp(a); }

template <class V> struct Lanes;
template <> struct Lanes<float2> {
    static constexpr int kWidth = 2;
    static __device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }  // scalar-broadcast operand in SASS
    static __device__ __forceinline__ float2 ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
    static __device__ __forceinline__ void st(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
    static __device__ __forceinline__ float hsum(float2 v) { return v.x + v.y; }
    static __device__ __forceinline__ float hmin(float2 v) { return fminf(v.x, v.y); }
    static __device__ __forceinline__ float habsmax(float2 v) { return fmaxf(fabsf(v.x), fabsf(v.y)); }
    static __device__ __forceinline__ float get(float2 v, int k) { return k ? v.y : v.x; }
    static __device__ __forceinline__ float2 make(float a, float b) { return make_float2(a, b); }
};
template <> struct Lanes<float> {
    static constexpr int kWidth = 1;
    static __device__ __forceinline__ float bc(float s) { return s; }
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
    static __device__ __forceinline__ float hsum(float v) { return v; }
    static __device__ __forceinline__ float hmin(float v) { return v; }
    static __device__ __forceinline__ float habsmax(float v) { return fabsf(v); }
    static __device__ __forceinline__ float get(float v, int) { return v; }
    static __device__ __forceinline__ float make(float a, float) { return a; }
};

// ------------------------------------------------------------------ per-evaluation constants
// Per-evaluation constants of the delta pass.
struct DeltaStep {
    float cp, sp, txp, typ, tzp;   // candidate pose: cos/sin yaw', t'
    float ncdm1, sd;               // -(cos dyaw - 1), sin dyaw          (dyaw = yaw' - yaw)
    float dtx, dty, dtz;           // t' - t
};

// Camera constants in the normalised formulation.
struct CamN {
    float fx, fy, cx, cy, ifx, ify, ncxi, ncyi;  // 1/fx, 1/fy, -cx/fx, -cy/fy
};
__device__ __forceinline__ CamN make_camn(const Camera<float>& c) {
    CamN n;
    n.fx = c.fx; n.fy = c.fy; n.cx = c.cx; n.cy = c.cy;
    n.ifx = 1.f / c.fx; n.ify = 1.f / c.fy;
    n.ncxi = -c.cx * n.ifx; n.ncyi = -c.cy * n.ify;
    return n;
}

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

// max |X|, |Y|, |Z| over the inliers of the object (object frame), from the first evaluation
struct Extent {
    float xm, ym, zm;
};

// Is the object's bounding box, posed at (cos, sin, t), inside the clip window?  Then no point can be near a clip bound
// and the pass skips its per-point tests.  |x' - xmid z'| <= |tx - xmid tz| + ex + |xmid| ez and z' >= tz - ez for
// every point, ex = |c| Xm + |s| Zm, ez = |s| Xm + |c| Zm the half extents of the rotated box.
__device__ __forceinline__ bool box_inside_window(const ClipWindow& w, const Extent& e, float c, float s, float tx,
                                                  float ty, float tz) {
    const float ac = fabsf(c), as = fabsf(s);
    const float ex = fmaf(ac, e.xm, as * e.zm), ez = fmaf(as, e.xm, ac * e.zm);
    const float zmin = tz - ez;
    const float lx = (fabsf(fmaf(-w.xmid, tz, tx)) + fmaf(fabsf(w.xmid), ez, ex)) * 1.0001f;
    const float ly = (fabsf(fmaf(-w.ymid, tz, ty)) + fmaf(fabsf(w.ymid), ez, e.ym)) * 1.0001f;
    return (zmin >= w.zlo) && (lx <= w.xhalf * zmin) && (ly <= w.yhalf * zmin);  // false for NaN
}

// ------------------------------------------------------------------ the sums
// a[0..3]  J^T r (yaw, tx, ty, tz)          a[4..13] J^T J upper triangle (00 01 02 03 11 12 13 22 23 33)
// a[14]    sum |r|^2 (evaluations from the observations) or its CHANGE (delta pass)
// a[3], a[7], a[10], a[12] are accumulated with the opposite sign (see kNegatedSums).
constexpr unsigned kNegatedSums = (1u << 3) | (1u << 7) | (1u << 10) | (1u << 12);

// One point (pair): normal-equation terms from the candidate-frame quantities, v = M e, and M.
template <int WMODE, class V>
__device__ __forceinline__ void accumulate_normal(V a[15], V qx, V qz, V xn, V yn, V iz, V v0, V v1, V m00, V m01, V m11) {
    const V h0 = vmul(iz, v0), h1 = vmul(iz, v1);
    const V ju = vfma3(xn, qx, qz), jv = vmul(yn, qx);
    a[0] = vfma3(ju, h0, vfma3(jv, h1, a[0]));
    a[1] = vadd(a[1], h0);
    a[2] = vadd(a[2], h1);
    a[3] = vfma3(xn, h0, vfma3(yn, h1, a[3]));
    const V iz2 = vmul(iz, iz);
    const V n00 = vmul(iz2, m00), n11 = vmul(iz2, m11);
    V p00, p10, q03, q13;
    if (WMODE == MRPNP_W_FULL) {
        const V n01 = vmul(iz2, m01);
        p00 = vfma3(n00, ju, vmul(n01, jv));
        p10 = vfma3(n01, ju, vmul(n11, jv));
        q03 = vfma3(n00, xn, vmul(n01, yn));
        q13 = vfma3(n01, xn, vmul(n11, yn));
        a[9] = vadd(a[9], n01);
    } else {
        p00 = vmul(n00, ju);
        p10 = vmul(n11, jv);
        q03 = vmul(n00, xn);
        q13 = vmul(n11, yn);
    }
    a[4] = vfma3(ju, p00, vfma3(jv, p10, a[4]));
    a[5] = vadd(a[5], p00);
    a[6] = vadd(a[6], p10);
    a[7] = vfma3(ju, q03, vfma3(jv, q13, a[7]));
    a[8] = vadd(a[8], n00);
    a[10] = vadd(a[10], q03);
    a[11] = vadd(a[11], n11);
    a[12] = vadd(a[12], q13);
    a[13] = vfma3(xn, q03, vfma3(yn, q13, a[13]));
}

// Running clip / extent observations of a pass.
struct PassFlags {
    float mz, mx, my;      // min z', max |xn - xmid|, max |yn - ymid|
    float ex, ey, ez;      // max |X|, |Y|, |Z|  (first evaluation only)
};

// kPassFirst: an evaluation from the observations -- the initial point (plain fp32; also turns W into M) or, with
// PassArgs::anchor, the fp64 anchor.  One body for both (the remainder routine reads the switch at run time; the main
// loops instantiate it once per kind).
enum PassKind { kPassFirst = 0, kPassDelta = 2, kPassUndo = 3 };

// Uniform inputs of a pass (one struct for all kinds keeps the out-of-line remainder routine to one signature).
struct PassArgs {
    float cs, sn, tx, ty, tz;   // evaluation point (kPassFirst); for the other kinds see `step`
    DeltaStep step;
    const float* consts;        // shared memory: CamN (8 floats) then ClipWindow (5 floats) of the object -- read inside
                                // the passes, so that they are not live in registers across the scalar LM algebra
    bool check;                 // per-point clip tests (always on for kPassFirst)
    bool anchor;                // kPassFirst: the fp64 anchor evaluation instead of the initial one
};

__device__ __forceinline__ CamN args_cam(const PassArgs& u) {
    CamN c;
    c.fx = u.consts[0]; c.fy = u.consts[1]; c.cx = u.consts[2]; c.cy = u.consts[3];
    c.ifx = u.consts[4]; c.ify = u.consts[5]; c.ncxi = u.consts[6]; c.ncyi = u.consts[7];
    return c;
}
__device__ __forceinline__ ClipWindow args_win(const PassArgs& u) {
    ClipWindow w;
    w.xmid = u.consts[8]; w.xhalf = u.consts[9]; w.ymid = u.consts[10]; w.yhalf = u.consts[11]; w.zlo = u.consts[12];
    return w;
}
__device__ __forceinline__ void store_consts(float* consts, const CamN& c, const ClipWindow& w, int lane) {
    if (lane == 0) {
        consts[0] = c.fx; consts[1] = c.fy; consts[2] = c.cx; consts[3] = c.cy;
        consts[4] = c.ifx; consts[5] = c.ify; consts[6] = c.ncxi; consts[7] = c.ncyi;
        consts[8] = w.xmid; consts[9] = w.xhalf; consts[10] = w.ymid; consts[11] = w.yhalf; consts[12] = w.zlo;
    }
    __syncwarp();
}

// The body of every pass for one point (V = float) or one pair of points (V = float2) at slot index idx.
// `live`: only read for V = float (the remainder path): a dead lane computes on point 0 with M = 0 and stores nothing.
template <int WMODE, int KIND, class V, int AN = -1>   // AN: PassArgs::anchor known at compile time (0 / 1), -1 = read it
__device__ __forceinline__ void pass_point(float* __restrict__ slot, int P, int idx, bool live, const PassArgs& u,
                                           const CamN& cam, const ClipWindow& win, V a[15], PassFlags& f) {
    using L = Lanes<V>;
    const bool anchor = AN < 0 ? u.anchor : (AN != 0);
    constexpr int WC = (WMODE == MRPNP_W_FULL) ? 3 : 2;
    float* s3 = slot;
    float* s2 = slot + 3 * P;
    float* sw = slot + 5 * P;
    const V X = L::ld(s3 + idx), Y = L::ld(s3 + P + idx), Z = L::ld(s3 + 2 * P + idx);
    V o0 = L::ld(s2 + idx), o1 = L::ld(s2 + P + idx);            // observations (u, v) or tracked e
    V m00 = L::ld(sw + idx), m11 = L::ld(sw + (WC - 1) * P + idx), m01 = L::bc(0.f);
    if (WMODE == MRPNP_W_FULL) m01 = L::ld(sw + P + idx);
    if (L::kWidth == 1 && !live) { m00 = L::bc(0.f); m11 = L::bc(0.f); m01 = L::bc(0.f); }

    if (KIND == kPassFirst) {
        if (!anchor) {
            // weights -> M = F W^T W F, stored in place of W for the later passes
            if (WMODE == MRPNP_W_FULL) {
                const V wxx = m00, wxy = m01, wyy = m11;
                m00 = vmul(vfma3(wxx, wxx, vmul(wxy, wxy)), L::bc(cam.fx * cam.fx));
                m11 = vmul(vfma3(wyy, wyy, vmul(wxy, wxy)), L::bc(cam.fy * cam.fy));
                m01 = vmul(vmul(wxy, vadd(wxx, wyy)), L::bc(cam.fx * cam.fy));
            } else {
                const V t0 = vmul(m00, L::bc(cam.fx)), t1 = vmul(m11, L::bc(cam.fy));
                m00 = vmul(t0, t0);
                m11 = vmul(t1, t1);
            }
            if (L::kWidth == 2 || live) {
                L::st(sw + idx, m00);
                L::st(sw + (WC - 1) * P + idx, m11);
                if (WMODE == MRPNP_W_FULL) L::st(sw + P + idx, m01);
            }
            f.ex = fmaxf(f.ex, L::habsmax(X));
            f.ey = fmaxf(f.ey, L::habsmax(Y));
            f.ez = fmaxf(f.ez, L::habsmax(Z));
        }
        const V qx = vfma(L::bc(u.cs), X, vmul(L::bc(u.sn), Z));
        const V qz = vfma(L::bc(u.cs), Z, vmul(L::bc(-u.sn), X));
        const V x1 = vadd(qx, L::bc(u.tx)), y1 = vadd(Y, L::bc(u.ty)), z1 = vadd(qz, L::bc(u.tz));
        const V iz = vrcp(z1);
        const V xn = vmul(x1, iz), yn = vmul(y1, iz);
        V e0, e1;
        if (!anchor) {
            // plain fp32: the residuals are many pixels here, and the decision on the first step is checked against
            // its own rounding (see the LM loop)
            e0 = vsub(xn, vfma(o0, L::bc(cam.ifx), L::bc(cam.ncxi)));
            e1 = vsub(yn, vfma(o1, L::bc(cam.ify), L::bc(cam.ncyi)));
        } else {
            // the anchor: residual chain in fp64, rounded once to fp32, REPLACES the observations in the slot
            const double cs = (double)u.cs, sn = (double)u.sn, tx = (double)u.tx, ty = (double)u.ty, tz = (double)u.tz;
            const double ifx = 1.0 / (double)cam.fx, ify = 1.0 / (double)cam.fy;   // loop-invariant
            const double ncxi = -(double)cam.cx * ifx, ncyi = -(double)cam.cy * ify;
            float r0[2], r1[2];
#pragma unroll
            for (int k = 0; k < L::kWidth; ++k) {
                const double Xd = (double)L::get(X, k), Yd = (double)L::get(Y, k), Zd = (double)L::get(Z, k);
                const double xc = fma(cs, Xd, fma(sn, Zd, tx));
                const double zc = fma(cs, Zd, fma(-sn, Xd, tz));
                const double yc = Yd + ty;
                const double izd = fast_rcp(zc);
                const double un = fma((double)L::get(o0, k), ifx, ncxi);   // (u - cx) / fx
                const double vn = fma((double)L::get(o1, k), ify, ncyi);
                r0[k] = (float)fma(xc, izd, -un);
                r1[k] = (float)fma(yc, izd, -vn);
            }
            e0 = L::make(r0[0], r0[L::kWidth - 1]);
            e1 = L::make(r1[0], r1[L::kWidth - 1]);
            if (L::kWidth == 2 || live) { L::st(s2 + idx, e0); L::st(s2 + P + idx, e1); }
        }
        // clip proximity on the fp32 projection (margins far above its rounding)
        f.mz = fminf(f.mz, L::hmin(z1));
        f.mx = fmaxf(f.mx, L::habsmax(vsub(xn, L::bc(win.xmid))));
        f.my = fmaxf(f.my, L::habsmax(vsub(yn, L::bc(win.ymid))));
        V v0, v1;
        if (WMODE == MRPNP_W_FULL) {
            v0 = vfma3(m00, e0, vmul(m01, e1));
            v1 = vfma3(m01, e0, vmul(m11, e1));
        } else {
            v0 = vmul(m00, e0);
            v1 = vmul(m11, e1);
        }
        a[14] = vfma3(e0, v0, vfma3(e1, v1, a[14]));
        accumulate_normal<WMODE, V>(a, qx, qz, xn, yn, iz, v0, v1, m00, m01, m11);
    } else {
        const DeltaStep& s = u.step;
        const V qx = vfma(L::bc(s.cp), X, vmul(L::bc(s.sp), Z));
        const V qz = vfma(L::bc(s.cp), Z, vmul(L::bc(-s.sp), X));
        const V x1 = vadd(qx, L::bc(s.txp)), y1 = vadd(Y, L::bc(s.typ)), z1 = vadd(qz, L::bc(s.tzp));
        const V izp = vrcp(z1);
        const V xnp = vmul(x1, izp), ynp = vmul(y1, izp);
        const V Dx = vfma(L::bc(s.ncdm1), qx, vfma(L::bc(s.sd), qz, L::bc(s.dtx)));
        const V Dz = vfma(L::bc(s.ncdm1), qz, vfma(L::bc(-s.sd), qx, L::bc(s.dtz)));
        const V izo = vrcp(vsub(z1, Dz));
        const V Du = vmul(vfma3(vneg(xnp), Dz, Dx), izo);
        const V Dv = vmul(vfma(vneg(ynp), Dz, L::bc(s.dty)), izo);
        if (KIND == kPassUndo) {  // roll a rejected candidate back: e = e' - De (one fp32 rounding away from the old e)
            if (L::kWidth == 2 || live) { L::st(s2 + idx, vsub(o0, Du)); L::st(s2 + P + idx, vsub(o1, Dv)); }
            return;
        }
        const V e0 = vadd(o0, Du), e1 = vadd(o1, Dv);
        if (L::kWidth == 2 || live) { L::st(s2 + idx, e0); L::st(s2 + P + idx, e1); }   // speculative: most steps are accepted
        if (u.check) {
            f.mz = fminf(f.mz, L::hmin(z1));
            f.mx = fmaxf(f.mx, L::habsmax(vsub(xnp, L::bc(win.xmid))));
            f.my = fmaxf(f.my, L::habsmax(vsub(ynp, L::bc(win.ymid))));
        }
        V v0, v1, d0, d1;
        if (WMODE == MRPNP_W_FULL) {
            v0 = vfma3(m00, e0, vmul(m01, e1));
            v1 = vfma3(m01, e0, vmul(m11, e1));
            d0 = vfma3(m00, Du, vmul(m01, Dv));
            d1 = vfma3(m01, Du, vmul(m11, Dv));
        } else {
            v0 = vmul(m00, e0);
            v1 = vmul(m11, e1);
            d0 = vmul(m00, Du);
            d1 = vmul(m11, Dv);
        }
        // |r'|^2 - |r|^2 = De^T (2 M e' - M De)
        a[14] = vfma3(Du, vfma(L::bc(2.f), v0, vneg(d0)), vfma3(Dv, vfma(L::bc(2.f), v1, vneg(d1)), a[14]));
        accumulate_normal<WMODE, V>(a, qx, qz, xnp, ynp, izp, v0, v1, m00, m01, m11);
    }
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

// The same totals through shared memory: the 15 per-lane sums go to a [15][36] float tile (row stride 36 keeps the 128-bit
// reads below free of bank conflicts), lane (k, h) = (lane & 15, lane >> 4) adds 16 of the 32 entries of value k with four
// 128-bit loads and one shuffle joins the halves: lanes k and k + 16 end with the total of value k (k < 15).
// 15 + 15 + 4 + 15 + 2 instructions against ~120 for the shuffle network above.
constexpr int kReduceTileFloats = 15 * 36 + 4;
__device__ __forceinline__ float warp_reduce15_smem(const float2 a[15], float* __restrict__ red, int lane) {
#pragma unroll
    for (int i = 0; i < 15; ++i) red[i * 36 + lane] = a[i].x + a[i].y;
    __syncwarp();
    const int k = min(lane & 15, 14), h = lane >> 4;
    const float4* p = reinterpret_cast<const float4*>(red + k * 36 + h * 16);
    const float4 q0 = p[0], q1 = p[1], q2 = p[2], q3 = p[3];
    float t = (((q0.x + q0.y) + (q0.z + q0.w)) + ((q1.x + q1.y) + (q1.z + q1.w))) +
              (((q2.x + q2.y) + (q2.z + q2.w)) + ((q3.x + q3.y) + (q3.z + q3.w)));
    t += __shfl_xor_sync(kFull, t, 16);
    return t;
}

__device__ __forceinline__ bool flags_raised(const PassFlags& f, const ClipWindow& w) {
    return !(f.mz >= w.zlo) || !(f.mx <= w.xhalf) || !(f.my <= w.yhalf);
}

// The out-of-line routines (remainder, roll-back) take their PassArgs through the warp's shared-memory scratch: a struct
// passed by value or reference to a non-inlined function would be given a home in LOCAL memory, and with ~106 KB of
// local memory per CTA against ~28 KB of L1 every access to it from the hot loop would be an L2 round trip.
__device__ __forceinline__ void stash_args(float* smem, const PassArgs& u, int lane) {
    if (lane == 0) {
        smem[0] = u.cs; smem[1] = u.sn; smem[2] = u.tx; smem[3] = u.ty; smem[4] = u.tz;
        smem[5] = u.step.cp; smem[6] = u.step.sp; smem[7] = u.step.txp; smem[8] = u.step.typ; smem[9] = u.step.tzp;
        smem[10] = u.step.ncdm1; smem[11] = u.step.sd; smem[12] = u.step.dtx; smem[13] = u.step.dty; smem[14] = u.step.dtz;
        smem[15] = u.anchor ? 1.f : 0.f;
    }
    __syncwarp();
}
__device__ __forceinline__ PassArgs fetch_args(const float* smem, const float* consts) {
    PassArgs u;
    u.cs = smem[0]; u.sn = smem[1]; u.tx = smem[2]; u.ty = smem[3]; u.tz = smem[4];
    u.step.cp = smem[5]; u.step.sp = smem[6]; u.step.txp = smem[7]; u.step.typ = smem[8]; u.step.tzp = smem[9];
    u.step.ncdm1 = smem[10]; u.step.sd = smem[11]; u.step.dtx = smem[12]; u.step.dty = smem[13]; u.step.dtz = smem[14];
    u.anchor = smem[15] != 0.f;
    u.consts = consts;
    u.check = true;
    return u;
}

// Remainder of a pass (cold, out of line): the points [start, n) that the padded 64-point groups do not cover, one
// point per lane.  Adds its 15 totals to scratch[0..14] (which hold the totals of the main loop), merges the extents
// (initial evaluation) into scratch[16..18] and returns the clip flag.  kind: PassKind; args: stash_args().
template <int WMODE>
__device__ __noinline__ bool pass_remainder(float* slot, int P, int start, int n, int lane, int kind, const float* args,
                                            const float* consts, float* scratch) {
    const PassArgs u = fetch_args(args, consts);
    const CamN cam = args_cam(u);
    const ClipWindow win = args_win(u);
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 0.f;
    PassFlags f = {1e30f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int base = start; base < n; base += 32) {
        const int idx = base + lane;
        const bool live = idx < n;
        if (kind == kPassFirst) pass_point<WMODE, kPassFirst, float>(slot, P, live ? idx : 0, live, u, cam, win, a, f);
        else if (kind == kPassDelta) pass_point<WMODE, kPassDelta, float>(slot, P, live ? idx : 0, live, u, cam, win, a, f);
        else pass_point<WMODE, kPassUndo, float>(slot, P, live ? idx : 0, live, u, cam, win, a, f);
    }
    __syncwarp();
    if (kind == kPassUndo) return false;
    const float tot = warp_reduce16_scatter(a, lane);
    if ((lane & 1) == 0 && lane < 30) scratch[lane >> 1] += ((kNegatedSums >> (lane >> 1)) & 1u) ? -tot : tot;
    if (kind == kPassFirst && !u.anchor) {
        const unsigned ex = __reduce_max_sync(kFull, __float_as_uint(f.ex)), ey = __reduce_max_sync(kFull, __float_as_uint(f.ey)),
                       ez = __reduce_max_sync(kFull, __float_as_uint(f.ez));
        if (lane == 0) {
            scratch[16] = fmaxf(scratch[16], __uint_as_float(ex));
            scratch[17] = fmaxf(scratch[17], __uint_as_float(ey));
            scratch[18] = fmaxf(scratch[18], __uint_as_float(ez));
        }
    }
    __syncwarp();
    return __any_sync(kFull, flags_raised(f, win));
}

// One evaluation: the packed loop over the 64-point groups [0, n_main) -- from the observations (first == true: the
// initial point or the anchor, PassArgs::anchor) or as a delta pass -- then ONE copy of the transposed reduction, the
// totals through the warp's scratch (scratch[0..14]; [16..18] = extents after the initial evaluation), and the
// remainder [n_main, n) if any.  Returns false if a total is not finite.
template <int WMODE>
__device__ __forceinline__ bool run_pass(float* slot, int P, int n_main, int n, int lane, const PassArgs& u, bool first,
                                         float* scratch, float* arg_stash, float* red, bool& flagged) {
    float2 a[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) a[i] = make_float2(0.f, 0.f);
    PassFlags f = {1e30f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const CamN cam = args_cam(u);
    const ClipWindow win = args_win(u);
    const int end = n_main + 2 * lane;
    if (first) {
#ifndef MRPNP_EXP_MERGED_FIRST   // one loop per kind: 2-3 % faster than one loop with a runtime switch (r02_ab_variants.txt, call 16)
#ifndef MRPNP_EXP_FIRST_UNROLL
#define MRPNP_EXP_FIRST_UNROLL 1   // 2 measured 5 % slower on the full-covariance set (r02_ab_variants.txt, call 53)
#endif
        constexpr int kFirstUnroll = MRPNP_EXP_FIRST_UNROLL;
        if (u.anchor) {
#pragma unroll kFirstUnroll
            for (int idx = 2 * lane; idx < end; idx += 64) pass_point<WMODE, kPassFirst, float2, 1>(slot, P, idx, true, u, cam, win, a, f);
        } else {
#pragma unroll kFirstUnroll
            for (int idx = 2 * lane; idx < end; idx += 64) pass_point<WMODE, kPassFirst, float2, 0>(slot, P, idx, true, u, cam, win, a, f);
        }
#else
#pragma unroll 1
        for (int idx = 2 * lane; idx < end; idx += 64) pass_point<WMODE, kPassFirst, float2>(slot, P, idx, true, u, cam, win, a, f);
#endif
    } else {
#ifndef MRPNP_EXP_DELTA_UNROLL
#define MRPNP_EXP_DELTA_UNROLL 1   // 2 measured 4 % slower (instruction cache)
#endif
        constexpr int kDeltaUnroll = MRPNP_EXP_DELTA_UNROLL;
#pragma unroll kDeltaUnroll
        for (int idx = 2 * lane; idx < end; idx += 64) pass_point<WMODE, kPassDelta, float2>(slot, P, idx, true, u, cam, win, a, f);
    }
#ifdef MRPNP_EXP_SHFL_REDUCE
    float s[16];
#pragma unroll
    for (int i = 0; i < 15; ++i) s[i] = a[i].x + a[i].y;
    s[15] = 0.f;
    const float tot = warp_reduce16_scatter(s, lane);
    __syncwarp();
    if ((lane & 1) == 0) scratch[lane >> 1] = ((kNegatedSums >> (lane >> 1)) & 1u) ? -tot : tot;
#else
    const float tot = warp_reduce15_smem(a, red, lane);
    __syncwarp();
    if (lane < 16) scratch[lane] = lane == 15 ? 0.f : (((kNegatedSums >> lane) & 1u) ? -tot : tot);   // [15]: read by the finite test
#endif
    const unsigned ex = __reduce_max_sync(kFull, __float_as_uint(f.ex)), ey = __reduce_max_sync(kFull, __float_as_uint(f.ey)),
                   ez = __reduce_max_sync(kFull, __float_as_uint(f.ez));
    if (lane == 0) { scratch[16] = __uint_as_float(ex); scratch[17] = __uint_as_float(ey); scratch[18] = __uint_as_float(ez); }
    __syncwarp();
    flagged = (first || u.check) && __any_sync(kFull, flags_raised(f, win));
    if (n > n_main) {
        stash_args(arg_stash, u, lane);
        flagged = pass_remainder<WMODE>(slot, P, n_main, n, lane, first ? kPassFirst : kPassDelta, arg_stash, u.consts, scratch) || flagged;
    }
    return __all_sync(kFull, fabsf(scratch[lane & 15]) < 3.0e38f);
}

// ------------------------------------------------------------------ delta passes of LONG objects, split four ways
// EXPERIMENT (MRPNP_EXP_FAST_TEAM, off): built, parity- and sanitizer-clean, and SLOWER -- 184.7 us against 159.7 us on the
// full workload, 148.4 against 142.1 us on the diagonal one (profiles/r02_ab_variants.txt, call 31).  The fixed
// association needs four reductions per pass also when nobody helps, which lengthens exactly the long objects that form
// the tail, and help only comes when ALL other warps of the CTA are idle.
// From its kTeamFromEval-th evaluation on, an object's delta pass is computed as kTeamParts separate partial passes
// (part v = the 64-point groups v, v + 4, ...; each with its own reduction) whose totals are added in part order.  That
// fixed association makes the result independent of WHO computes the parts: the owner alone, one after the other, or --
// when every other warp of the CTA has run out of work -- four warps at once (FastTeam).  Long objects are what the tail
// of a launch consists of, and in the tail their CTA has nothing else to do.
constexpr int kTeamParts = 4;
constexpr int kTeamFromEval = 8;
constexpr int kTeamMinPoints = 256;
enum { kFtPass = 0, kFtExit = 1 };
struct FastTeam {
    uint64_t bar;            // mbarrier, one arrival per warp
    int idle;                // warps that have run out of fresh objects (they sit in the waiting room)
    int cmd, owner, n_main, P;
    float* owner_hdr;
    float* owner_slot;
};
__device__ __forceinline__ void ft_barrier(FastTeam* ft, uint32_t& phase, int lane) { cta_round(2, &ft->bar, phase, lane); }

// One part of a split delta pass over the owner's slot; args / consts: the owner's stash; red: the CALLER's reduction tile;
// out[0..14]: totals (not yet sign-corrected), out[16]: clip flag.
template <int WMODE>
__device__ __noinline__ void team_part(const float* args, const float* consts, float* slot, int P, int n_main, int v, float* red,
                                       float* out, int lane) {
    const PassArgs u = fetch_args(args, consts);
    const CamN cam = args_cam(u);
    const ClipWindow win = args_win(u);
    float2 a[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) a[i] = make_float2(0.f, 0.f);
    PassFlags f = {1e30f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int g = v; 64 * g < n_main; g += kTeamParts) pass_point<WMODE, kPassDelta, float2>(slot, P, 64 * g + 2 * lane, true, u, cam, win, a, f);
    const float tot = warp_reduce15_smem(a, red, lane);
    __syncwarp();
    if (lane < 15) out[lane] = tot;
    const bool fl = __any_sync(kFull, flags_raised(f, win));
    if (lane == 16) out[16] = fl ? 1.f : 0.f;
    __syncwarp();
}

// Roll the speculative residual update of a rejected candidate back (out of line: rare).  args: stash_args().
template <int WMODE>
__device__ __noinline__ void undo_pass(float* slot, int P, int n_main, int n, int lane, const float* args, const float* consts) {
    const PassArgs u = fetch_args(args, consts);
    const CamN cam = args_cam(u);
    const ClipWindow win = args_win(u);
    float2 a[15];
    PassFlags f;
    const int end = n_main + 2 * lane;
#pragma unroll 1
    for (int idx = 2 * lane; idx < end; idx += 64) pass_point<WMODE, kPassUndo, float2>(slot, P, idx, true, u, cam, win, a, f);
    if (n > n_main) pass_remainder<WMODE>(slot, P, n_main, n, lane, kPassUndo, args, consts, nullptr);
    __syncwarp();
}

}  // namespace mrpnp
