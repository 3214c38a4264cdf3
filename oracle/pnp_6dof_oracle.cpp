// ============================================================================
// oracle/pnp_6dof_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU fp64 statement of the 6-DoF extension of MonoRUn's uncertainty PnP named by the north star
// (BASELINE.json): unknowns [rx, ry, rz, tx, ty, tz], rotation by ceres::AngleAxisRotatePoint, everything else --
// projection, clips, per-axis or full 2x2 whitening, Ceres 1.14 trust-region LM with DENSE_QR, covariance
// (J^T J)^-1 -- exactly as the reference's 4-DoF op (monorun/ops/least_squares/src/pnp_uncert_cpu.cpp:9-74, :245-292).
//
// NO REFERENCE EXISTS for this variant: the reference's rotation vector is hard-wired to (0, yaw, 0)
// (pnp_uncert_cpu.cpp:28) and its `use_6dof` flag is never read (pnp_uncert.py:11,98,122,142).
// PARITY UNPINNED by construction.  To stay independent of the product's closed-form rotation derivative, the
// Jacobian here comes from forward-mode dual numbers pushed through a restatement of AngleAxisRotatePoint
// (ceres/rotation.h: Rodrigues formula for theta^2 > epsilon, first-order w x p otherwise) -- the computation
// Ceres' AutoDiffCostFunction would perform.  tests/test_6dof.py checks it against the 4-DoF oracle (rotation held
// about y), finite differences, scipy and noise-free recovery of general rotations.
// ============================================================================
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "ceres_lm.h"

namespace {

using namespace ceres_lm;

constexpr int N = 6;


// ceres::Jet<double, 6>, the operations the functor needs.
struct Jet {
    double a;
    double v[N];
    Jet() : a(0.0) { for (int k = 0; k < N; ++k) v[k] = 0.0; }
    explicit Jet(double s) : a(s) { for (int k = 0; k < N; ++k) v[k] = 0.0; }
    static Jet variable(double s, int k) { Jet j(s); j.v[k] = 1.0; return j; }
};
inline Jet operator+(const Jet& x, const Jet& y) { Jet r(x.a + y.a); for (int k = 0; k < N; ++k) r.v[k] = x.v[k] + y.v[k]; return r; }
inline Jet operator-(const Jet& x, const Jet& y) { Jet r(x.a - y.a); for (int k = 0; k < N; ++k) r.v[k] = x.v[k] - y.v[k]; return r; }
inline Jet operator-(const Jet& x) { Jet r(-x.a); for (int k = 0; k < N; ++k) r.v[k] = -x.v[k]; return r; }
inline Jet operator*(const Jet& x, const Jet& y) { Jet r(x.a * y.a); for (int k = 0; k < N; ++k) r.v[k] = x.a * y.v[k] + x.v[k] * y.a; return r; }
inline Jet operator/(const Jet& x, const Jet& y) {
    const double inv = 1.0 / y.a, q = x.a * inv;
    Jet r(q);
    for (int k = 0; k < N; ++k) r.v[k] = (x.v[k] - q * y.v[k]) * inv;
    return r;
}
inline Jet sqrt(const Jet& x) { const double s = std::sqrt(x.a); Jet r(s); for (int k = 0; k < N; ++k) r.v[k] = x.v[k] / (2.0 * s); return r; }
inline Jet sin(const Jet& x) { Jet r(std::sin(x.a)); const double c = std::cos(x.a); for (int k = 0; k < N; ++k) r.v[k] = c * x.v[k]; return r; }
inline Jet cos(const Jet& x) { Jet r(std::cos(x.a)); const double s = -std::sin(x.a); for (int k = 0; k < N; ++k) r.v[k] = s * x.v[k]; return r; }
inline bool operator<(const Jet& x, const Jet& y) { return x.a < y.a; }
inline bool operator>(const Jet& x, const Jet& y) { return x.a > y.a; }

// ceres::AngleAxisRotatePoint (rotation.h), templated like the original.
template <typename T>
void angle_axis_rotate_point(const T aa[3], const T pt[3], T out[3]) {
    using std::sqrt; using std::sin; using std::cos;  // doubles; the Jet overloads are found by ADL
    const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
    if (theta2 > T(std::numeric_limits<double>::epsilon())) {
        const T theta = sqrt(theta2), costheta = cos(theta), sintheta = sin(theta), inv = T(1.0) / theta;
        const T w[3] = {aa[0] * inv, aa[1] * inv, aa[2] * inv};
        const T wxp[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
        const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
        for (int i = 0; i < 3; ++i) out[i] = pt[i] * costheta + wxp[i] * sintheta + w[i] * tmp;
    } else {
        const T wxp[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2], aa[0] * pt[1] - aa[1] * pt[0]};
        for (int i = 0; i < 3; ++i) out[i] = pt[i] + wxp[i];
    }
}

struct Problem {
    const double *pts2d, *pts3d, *wgt2d;
    int pn;
    bool full_w;
    double fx, fy, cx, cy, z_min, u_min, u_max, v_min, v_max;
};

// The reference functor (pnp_uncert_cpu.cpp:24-51 diag, :189-217 whitening) with r_vec = x[0..3) instead of
// (0, yaw, 0).  max() and the clamping ternaries select a branch, value and derivative alike, as ceres::Jet does.
template <typename T>
void functor(const Problem& P, int i, const T* x, T* r) {
    const T pt[3] = {T(P.pts3d[i * 3]), T(P.pts3d[i * 3 + 1]), T(P.pts3d[i * 3 + 2])};
    T q[3];
    angle_axis_rotate_point(x, pt, q);
    q[0] = q[0] + x[3]; q[1] = q[1] + x[4]; q[2] = q[2] + x[5];
    if (q[2] < T(P.z_min)) q[2] = T(P.z_min);                                  // :36  max(z, z_min)
    T pu = T(P.fx) * q[0] / q[2] + T(P.cx), pv = T(P.fy) * q[1] / q[2] + T(P.cy);  // :38-39
    pu = (pu < T(P.u_min)) ? T(P.u_min) : (pu > T(P.u_max)) ? T(P.u_max) : pu;  // :41
    pv = (pv < T(P.v_min)) ? T(P.v_min) : (pv > T(P.v_max)) ? T(P.v_max) : pv;  // :42
    const T du = pu - T(P.pts2d[i * 2]), dv = pv - T(P.pts2d[i * 2 + 1]);
    if (P.full_w) {                                                            // :214-215
        const T wxx(P.wgt2d[i * 3]), wxy(P.wgt2d[i * 3 + 1]), wyy(P.wgt2d[i * 3 + 2]);
        r[0] = wxx * du + wxy * dv;
        r[1] = wxy * du + wyy * dv;
    } else {                                                                   // :47-48
        r[0] = T(P.wgt2d[i * 2]) * du;
        r[1] = T(P.wgt2d[i * 2 + 1]) * dv;
    }
}

// Evaluator::Evaluate: cost = 1/2 |r|^2; res 2pn; jac (2pn x 6 row-major) and grad optional.
bool evaluate(const Problem& P, const double* x, double* cost, double* res, double* jac, double* grad) {
    double acc = 0.0, g[N] = {0};
    bool ok = true;
    Jet xj[N];
    for (int k = 0; k < N; ++k) xj[k] = Jet::variable(x[k], k);
    for (int i = 0; i < P.pn; ++i) {
        double rr[2];
        if (jac) {
            Jet rj[2];
            functor<Jet>(P, i, xj, rj);
            for (int a = 0; a < 2; ++a) {
                rr[a] = rj[a].a;
                for (int k = 0; k < N; ++k) {
                    jac[(2 * i + a) * N + k] = rj[a].v[k];
                    g[k] += rj[a].v[k] * rj[a].a;
                    ok = ok && std::isfinite(rj[a].v[k]);
                }
            }
        } else {
            functor<double>(P, i, x, rr);
        }
        acc += rr[0] * rr[0] + rr[1] * rr[1];
        if (res) { res[2 * i] = rr[0]; res[2 * i + 1] = rr[1]; }
    }
    *cost = 0.5 * acc;
    if (grad) std::memcpy(grad, g, sizeof(g));
    return ok && std::isfinite(acc);
}

// (J^T J)^-1 by Cholesky (ceres::Covariance on a full-rank problem); false if not positive definite.
bool spd_inverse(const double* H, double* inv) {
    double L[N * N] = {0};
    for (int i = 0; i < N; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = H[i * N + j];
            for (int k = 0; k < j; ++k) s -= L[i * N + k] * L[j * N + k];
            if (i == j) {
                if (!(s > 0.0) || !std::isfinite(s)) return false;
                L[i * N + i] = std::sqrt(s);
            } else {
                L[i * N + j] = s / L[j * N + j];
            }
        }
    for (int c = 0; c < N; ++c) {  // solve L L^T y = e_c
        double z[N], y[N];
        for (int i = 0; i < N; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = 0; k < i; ++k) s -= L[i * N + k] * z[k];
            z[i] = s / L[i * N + i];
        }
        for (int i = N - 1; i >= 0; --i) {
            double s = z[i];
            for (int k = i + 1; k < N; ++k) s -= L[k * N + i] * y[k];
            y[i] = s / L[i * N + i];
        }
        for (int i = 0; i < N; ++i) inv[i * N + c] = y[i];
    }
    return true;
}

Problem make_problem(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K, int pn,
                     const double* clips, bool full_w) {
    Problem P;
    P.pts2d = pts2d; P.pts3d = pts3d; P.wgt2d = wgt2d; P.pn = pn; P.full_w = full_w;
    P.fx = K[0]; P.fy = K[4]; P.cx = K[2]; P.cy = K[5];
    P.z_min = clips[0]; P.u_min = clips[1]; P.u_max = clips[2]; P.v_min = clips[3]; P.v_max = clips[4];
    return P;
}

void solve_one(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K, const double* init,
               int* result_val, double* result_pose, double* result_cov, int pn, const double* clips, bool full_w,
               int* stats, double* final_cost) {
    const Problem P = make_problem(pts2d, pts3d, wgt2d, K, pn, clips, full_w);
    std::memcpy(result_pose, init, N * sizeof(double));
    LMOptions opt;
    auto eval = [&P](const double* x, double* cost, double* res, double* jac, double* grad) {
        return evaluate(P, x, cost, res, jac, grad);
    };
    const LMResult r = trust_region_lm_n<N>(eval, 2 * pn, result_pose, opt);  // oracle/ceres_lm.h
    *result_val = (r.term == CONVERGENCE || r.term == NO_CONVERGENCE);
    if (stats) { stats[0] = r.iterations; stats[1] = r.num_cost_evals; stats[2] = r.num_jac_evals; stats[3] = r.term; }
    if (final_cost) *final_cost = r.final_cost;
    if (*result_val && result_cov) {
        std::vector<double> jac(static_cast<size_t>(2 * pn) * N);
        double cost, H[N * N] = {0};
        evaluate(P, result_pose, &cost, nullptr, jac.data(), nullptr);
        for (int i = 0; i < 2 * pn; ++i)
            for (int a = 0; a < N; ++a) for (int b = 0; b < N; ++b) H[a * N + b] += jac[i * N + a] * jac[i * N + b];
        *result_val = spd_inverse(H, result_cov) ? 1 : 0;
    }
}

}  // namespace

extern "C" {

// Batched driver: object b owns pn[b] points from point offset off[b] of the packed arrays; K [nb,9], clips [nb,5],
// init / result_pose [nb,6], result_cov [nb,36] or NULL, stats [nb,4], cost [nb].
void pnp_6dof_batch(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K, const double* init,
                    int* result_val, double* result_pose, double* result_cov, const int* pn, const long long* off,
                    const double* clips, int nb, int full_w, int* stats, double* cost, int threads) {
    const int wc = full_w ? 3 : 2;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 8) num_threads(threads)
#endif
    for (int b = 0; b < nb; ++b) {
        const long long o = off[b];
        solve_one(pts2d + o * 2, pts3d + o * 3, wgt2d + o * wc, K + b * 9, init + b * N, result_val + b,
                  result_pose + b * N, result_cov ? result_cov + b * N * N : nullptr, pn[b], clips + b * 5,
                  full_w != 0, stats ? stats + b * 4 : nullptr, cost ? cost + b : nullptr);
    }
}

// cost, gradient (6) and J^T J (6x6) at `pose` (what the LM loop sees).  For tests.
void pnp_6dof_eval(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K, const double* pose,
                   int pn, const double* clips, int full_w, double* cost, double* grad, double* JtJ) {
    const Problem P = make_problem(pts2d, pts3d, wgt2d, K, pn, clips, full_w != 0);
    std::vector<double> res(2 * pn), jac(static_cast<size_t>(2 * pn) * N);
    evaluate(P, pose, cost, res.data(), jac.data(), grad);
    if (JtJ) {
        std::memset(JtJ, 0, N * N * sizeof(double));
        for (int i = 0; i < 2 * pn; ++i)
            for (int a = 0; a < N; ++a) for (int b = 0; b < N; ++b) JtJ[a * N + b] += jac[i * N + a] * jac[i * N + b];
    }
}

}  // extern "C"
