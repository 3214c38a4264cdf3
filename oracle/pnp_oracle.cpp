// ============================================================================
// oracle/pnp_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU fp64 restatement of MonoRUn's native uncertainty-PnP op, used only as the
// parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs.  Nothing under monorun_b200/ may import, link or call it.
//
// PARITY UNPINNED AGAINST THE REFERENCE BINARY (pinned to Ceres' published known answers instead, see
// below): the reference links ceres-solver
// 1.14.0 (INSTALL.md:13,29-30; monorun/ops/least_squares/setup.py:20-22), which is
// neither vendored in /root/reference nor installable here (no Ceres/Eigen/glog, no
// network), and the reference ships no tests or golden vectors for this path.  The
// trust-region Levenberg-Marquardt loop below restates the published algorithm of
// Ceres 1.14 (TrustRegionMinimizer + LevenbergMarquardtStrategy + DenseQRSolver with
// default Solver::Options) from knowledge of the upstream source.  So no output of the reference BINARY
// on a PnP input exists to compare with.  What IS pinned:
//   * the minimiser itself against Ceres' OWN published known answers: the same template that solves the
//     PnP problems (trust_region_lm_n, oracle/ceres_lm.h) reproduces, to every printed digit, the iteration tables of the Ceres
//     tutorial problems (helloworld.cc: 3 rows; powell.cc: all 15 rows of cost / cost_change / |gradient| /
//     |step| / tr_ratio / tr_radius, the final x and "Gradient max norm 3.642190e-11") --
//     tests/test_oracle.py::test_minimiser_reproduces_the_ceres_tutorial_tables.  That fixes Jacobi scaling,
//     the LM diagonal, step acceptance, the radius update and the termination tests;
//   * the residual / clip semantics follow the reference functor line by line
//     (monorun/ops/least_squares/src/pnp_uncert_cpu.cpp:24-51 diag weights,
//      :189-217 full 2x2 weights restricted to the 4 pose parameters);
//   * the pipeline covariance H = J^T J is checked against the reference's own
//     pure-torch approx_hessian (monorun/ops/least_squares/hessian.py:67-87,
//     jacobian.py:4-98), imported from /root/reference by tests/golden/make_golden.py;
//   * the minimiser itself is cross-checked against scipy.optimize.least_squares.
//
// C ABI: `pnp_uncert` has exactly the signature of
// monorun/ops/least_squares/src/ext.h:1-13 so the restated Python driver
// (oracle/pnp_driver.py) binds it the way pnp_uncert_cpu.py:102-106 does.
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

struct Problem {
    const double* pts2d;  // n,2
    const double* pts3d;  // n,3
    const double* wgt2d;  // n,2 (diag) or n,3 (full: wxx,wxy,wyy)
    int pn;
    bool full_w;
    double fx, fy, cx, cy;
    double z_min, u_min, u_max, v_min, v_max;
};

// One reprojection block, Ceres-Jet semantics (pnp_uncert_cpu.cpp:24-51 / :189-217):
// value and derivative follow the *selected branch* of max() / the ternary clamps, so a
// clipped depth keeps d(u)/d(x') but drops d(u)/d(z'), and a clamped u/v has zero derivative.
// jac: 2x4 row-major, columns [yaw, tx, ty, tz]; may be NULL.
inline void eval_point(const Problem& P, int i, double yaw_s, double yaw_c, const double* t,
                       double* r, double* jac) {
    const double X = P.pts3d[i * 3], Y = P.pts3d[i * 3 + 1], Z = P.pts3d[i * 3 + 2];
    // ceres::AngleAxisRotatePoint((0,yaw,0), X) == R_y(yaw) X   (pnp_uncert_cpu.cpp:28-31)
    const double qx = yaw_c * X + yaw_s * Z;
    const double qz = -yaw_s * X + yaw_c * Z;
    const double xc = qx + t[0], yc = Y + t[1], zc = qz + t[2];  // :32-34
    const bool z_free = !(zc < P.z_min);                          // :36  max(z, z_min)
    const double z = z_free ? zc : P.z_min;
    const double iz = 1.0 / z;
    double pu = P.fx * xc * iz + P.cx;                            // :38
    double pv = P.fy * yc * iz + P.cy;                            // :39
    bool u_free = true, v_free = true;
    if (pu < P.u_min) { pu = P.u_min; u_free = false; }           // :41
    else if (pu > P.u_max) { pu = P.u_max; u_free = false; }
    if (pv < P.v_min) { pv = P.v_min; v_free = false; }           // :42
    else if (pv > P.v_max) { pv = P.v_max; v_free = false; }
    const double du = pu - P.pts2d[i * 2], dv = pv - P.pts2d[i * 2 + 1];  // :44-45
    double w00, w01, w11;
    if (P.full_w) {                                               // :214-215
        w00 = P.wgt2d[i * 3]; w01 = P.wgt2d[i * 3 + 1]; w11 = P.wgt2d[i * 3 + 2];
    } else {                                                      // :47-48
        w00 = P.wgt2d[i * 2]; w01 = 0.0; w11 = P.wgt2d[i * 2 + 1];
    }
    r[0] = w00 * du + w01 * dv;
    r[1] = w01 * du + w11 * dv;
    if (jac) {
        // d(x',y',z')/d(yaw) = (qz, 0, -qx); d/dt = I
        const double mz = z_free ? 1.0 : 0.0;
        double ju[4] = {0, 0, 0, 0}, jv[4] = {0, 0, 0, 0};
        if (u_free) {
            const double a = P.fx * iz, b = -P.fx * xc * iz * iz * mz;  // d/dx', d/dz'
            ju[0] = a * qz + b * (-qx); ju[1] = a; ju[2] = 0.0; ju[3] = b;
        }
        if (v_free) {
            const double a = P.fy * iz, b = -P.fy * yc * iz * iz * mz;
            jv[0] = b * (-qx); jv[1] = 0.0; jv[2] = a; jv[3] = b;
        }
        for (int k = 0; k < 4; ++k) {
            jac[k] = w00 * ju[k] + w01 * jv[k];
            jac[4 + k] = w01 * ju[k] + w11 * jv[k];
        }
    }
}

// Evaluator::Evaluate: cost = 1/2 |r|^2, residuals (2n), jacobian (2n x 4 row-major),
// gradient = J^T r.  Returns false when anything is non-finite (Ceres rejects such
// evaluations: ResidualBlock::Evaluate -> IsEvaluationValid).
bool evaluate(const Problem& P, const double* x, double* cost, double* res, double* jac,
              double* grad) {
    const double s = std::sin(x[0]), c = std::cos(x[0]);
    double acc = 0.0;
    double g[4] = {0, 0, 0, 0};
    double rr[2], jj[8];
    bool ok = true;
    for (int i = 0; i < P.pn; ++i) {
        eval_point(P, i, s, c, x + 1, rr, jac ? jj : nullptr);
        acc += rr[0] * rr[0] + rr[1] * rr[1];
        if (res) { res[2 * i] = rr[0]; res[2 * i + 1] = rr[1]; }
        if (jac) {
            std::memcpy(jac + 8 * i, jj, sizeof(jj));
            for (int k = 0; k < 4; ++k) {
                g[k] += jj[k] * rr[0] + jj[4 + k] * rr[1];
                ok = ok && std::isfinite(jj[k]) && std::isfinite(jj[4 + k]);
            }
        }
    }
    *cost = 0.5 * acc;
    if (grad) std::memcpy(grad, g, sizeof(g));
    return ok && std::isfinite(acc);
}

// The PnP problem through the minimiser above (one 4-vector parameter block: yaw, t).
LMResult trust_region_lm(const Problem& P, double* x_io, const LMOptions& opt) {
    auto eval = [&P](const double* x, double* cost, double* res, double* jac, double* grad) {
        return evaluate(P, x, cost, res, jac, grad);
    };
    return trust_region_lm_n<4>(eval, 2 * P.pn, x_io, opt);
}

// (J^T J)^-1 for a symmetric positive definite 4x4 via Cholesky; false if not SPD.
bool spd_inverse4(const double* H, double* inv) {
    double L[16] = {0};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = H[i * 4 + j];
            for (int k = 0; k < j; ++k) s -= L[i * 4 + k] * L[j * 4 + k];
            if (i == j) {
                if (!(s > 0.0) || !std::isfinite(s)) return false;
                L[i * 4 + i] = std::sqrt(s);
            } else {
                L[i * 4 + j] = s / L[j * 4 + j];
            }
        }
    double Li[16] = {0};  // inverse of L (lower)
    for (int i = 0; i < 4; ++i) {
        Li[i * 4 + i] = 1.0 / L[i * 4 + i];
        for (int j = 0; j < i; ++j) {
            double s = 0.0;
            for (int k = j; k < i; ++k) s -= L[i * 4 + k] * Li[k * 4 + j];
            Li[i * 4 + j] = s / L[i * 4 + i];
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = std::max(i, j); k < 4; ++k) s += Li[k * 4 + i] * Li[k * 4 + j];
            inv[i * 4 + j] = s;
        }
    return true;
}

Problem make_problem(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K,
                     int pn, const double* clips, bool full_w) {
    Problem P;
    P.pts2d = pts2d; P.pts3d = pts3d; P.wgt2d = wgt2d; P.pn = pn; P.full_w = full_w;
    P.fx = K[0]; P.fy = K[4]; P.cx = K[2]; P.cy = K[5];  // pnp_uncert_cpu.cpp:265
    P.z_min = clips[0]; P.u_min = clips[1]; P.u_max = clips[2]; P.v_min = clips[3]; P.v_max = clips[4];
    return P;
}

LMOptions g_options;  // process-wide; only the ftol switch is ever changed (tests)

void solve_one(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K,
               const double* init_pose, int* result_val, double* result_pose, double* result_cov,
               double* result_tr, int pn, const double* clips, bool full_w, int* stats /*4 or NULL*/,
               double* final_cost /*or NULL*/) {
    Problem P = make_problem(pts2d, pts3d, wgt2d, K, pn, clips, full_w);
    std::memcpy(result_pose, init_pose, 4 * sizeof(double));           // cpp:259
    LMResult r = trust_region_lm(P, result_pose, g_options);           // cpp:270-274
    *result_val = (r.term == CONVERGENCE || r.term == NO_CONVERGENCE); // IsSolutionUsable, cpp:276
    if (result_tr) *result_tr = r.tr_radius;                           // cpp:277
    if (stats) { stats[0] = r.iterations; stats[1] = r.num_cost_evals; stats[2] = r.num_jac_evals; stats[3] = r.term; }
    if (final_cost) *final_cost = r.final_cost;
    if (*result_val && result_cov) {                                    // cpp:279-291, Ceres Covariance = (J^T J)^-1
        std::vector<double> jac(static_cast<size_t>(2 * pn) * 4);
        double cost, H[16] = {0};
        evaluate(P, result_pose, &cost, nullptr, jac.data(), nullptr);
        for (int i = 0; i < 2 * pn; ++i)
            for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) H[a * 4 + b] += jac[i * 4 + a] * jac[i * 4 + b];
        *result_val = spd_inverse4(H, result_cov) ? 1 : 0;              // rank-deficient -> Compute() fails
    }
}

}  // namespace

extern "C" {

// Exact ABI of monorun/ops/least_squares/src/ext.h:1-13.
void pnp_uncert(double* pts2d, double* pts3d, double* wgt2d, double* K, double* init_pose,
                int* result_val, double* result_pose, double* result_cov, double* result_tr,
                int pn, double* clips) {
    solve_one(pts2d, pts3d, wgt2d, K, init_pose, result_val, result_pose, result_cov, result_tr, pn, clips,
              false, nullptr, nullptr);
}

// 4-DoF solve with full symmetric 2x2 whitening W = [wxx wxy; wxy wyy]: the residual form of
// NocCovReprojectionErrorArray (pnp_uncert_cpu.cpp:214-215, ext.h:33) with the log-dimension
// unknowns held fixed at 0 -- BASELINE.json config 3.  wgt2d is [pn,3].
void pnp_uncert_fullw(double* pts2d, double* pts3d, double* wgt2d, double* K, double* init_pose,
                      int* result_val, double* result_pose, double* result_cov, double* result_tr,
                      int pn, double* clips) {
    solve_one(pts2d, pts3d, wgt2d, K, init_pose, result_val, result_pose, result_cov, result_tr, pn, clips,
              true, nullptr, nullptr);
}

// Batched driver used by the CPU-baseline timing and the parity tests.  Object b reads
// pn[b] points starting at point offset off[b] in the packed (sum pn, C) arrays (so inlier-
// compacted objects of different sizes can be stacked).  K is [nb,9], clips [nb,5].
// stats: [nb,4] = iterations, cost evals, jacobian evals, termination;  cost: [nb].
// threads <= 0 -> all OpenMP threads (the reference is single-threaded: pass 1).
void pnp_uncert_batch(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K,
                      const double* init_pose, int* result_val, double* result_pose, double* result_cov,
                      double* result_tr, const int* pn, const long long* off, const double* clips,
                      int nb, int full_w, int* stats, double* cost, int threads) {
    const int wc = full_w ? 3 : 2;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 8) num_threads(threads)
#endif
    for (int b = 0; b < nb; ++b) {
        const long long o = off[b];
        solve_one(pts2d + o * 2, pts3d + o * 3, wgt2d + o * wc, K + b * 9, init_pose + b * 4,
                  result_val + b, result_pose + b * 4, result_cov ? result_cov + b * 16 : nullptr,
                  result_tr ? result_tr + b : nullptr, pn[b], clips + b * 5, full_w != 0,
                  stats ? stats + b * 4 : nullptr, cost ? cost + b : nullptr);
    }
}

// cost = 1/2 |r|^2, gradient J^T r (4) and Gauss-Newton matrix J^T J (4x4) at `pose`
// with Ceres-Jet clip semantics (what the LM loop sees).  For tests.
void pnp_eval(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K,
              const double* pose, int pn, const double* clips, int full_w, double* cost, double* grad,
              double* JtJ) {
    Problem P = make_problem(pts2d, pts3d, wgt2d, K, pn, clips, full_w != 0);
    std::vector<double> jac(static_cast<size_t>(2 * pn) * 4);
    evaluate(P, pose, cost, nullptr, jac.data(), grad);
    for (int k = 0; k < 16; ++k) JtJ[k] = 0.0;
    for (int i = 0; i < 2 * pn; ++i)
        for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) JtJ[a * 4 + b] += jac[i * 4 + a] * jac[i * 4 + b];
}

// Pipeline covariance Hessian H = J^T J restating hessian.py:67-87 (approx_hessian) on top of
// jacobian.py:4-98: projection through the full K (jacobian.py:20-33), z clip kills BOTH rows of a
// point (:52-59), uv clip kills its own row, outliers are zeroed, weights are per-axis istd.
// inlier: [pn] bytes or NULL.  K row-major 3x3.  H: 4x4 row-major, order [yaw,tx,ty,tz].
void pnp_approx_hessian(const double* pts2d, const double* pts3d, const double* istd, const double* K,
                        const double* pose, const unsigned char* inlier, int pn, const double* clips,
                        double* H) {
    (void)pts2d;
    const double s = std::sin(pose[0]), c = std::cos(pose[0]);
    const double R[9] = {c, 0, s, 0, 1, 0, -s, 0, c};
    double KR[9], Kt[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            KR[i * 3 + j] = 0;
            for (int k = 0; k < 3; ++k) KR[i * 3 + j] += K[i * 3 + k] * R[k * 3 + j];
        }
        Kt[i] = K[i * 3] * pose[1] + K[i * 3 + 1] * pose[2] + K[i * 3 + 2] * pose[3];
    }
    // jac_yaw_m1 = K[0:2,[0,2]] @ [[-s, c], [-c, -s]]     (jacobian.py:74-80)
    const double m1[4] = {K[0] * -s + K[2] * -c, K[0] * c + K[2] * -s,
                          K[3] * -s + K[5] * -c, K[3] * c + K[5] * -s};
    for (int k = 0; k < 16; ++k) H[k] = 0.0;
    for (int i = 0; i < pn; ++i) {
        if (inlier && !inlier[i]) continue;
        const double X = pts3d[i * 3], Y = pts3d[i * 3 + 1], Z = pts3d[i * 3 + 2];
        double uvz[3];
        for (int a = 0; a < 3; ++a) uvz[a] = KR[a * 3] * X + KR[a * 3 + 1] * Y + KR[a * 3 + 2] * Z + Kt[a];
        double z = uvz[2];
        const bool zclip = z < clips[0];
        if (zclip) z = clips[0];
        double uv[2] = {uvz[0] / z, uvz[1] / z};
        const double lb[2] = {clips[1], clips[3]}, ub[2] = {clips[2], clips[4]};
        bool zero[2];
        for (int a = 0; a < 2; ++a) {
            const bool clip = uv[a] < lb[a] || uv[a] > ub[a];
            uv[a] = std::max(lb[a], std::min(ub[a], uv[a]));
            zero[a] = zclip || clip;
        }
        for (int a = 0; a < 2; ++a) {
            if (zero[a]) continue;
            const double w = istd[i * 2 + a];
            double j[4];
            j[0] = ((m1[a * 2] + uv[a] * c) * X + (m1[a * 2 + 1] + uv[a] * s) * Z) / z * w;
            j[1] = K[a * 3] / z * w;
            j[2] = K[a * 3 + 1] / z * w;
            j[3] = (K[a * 3 + 2] - uv[a]) / z * w;
            for (int p = 0; p < 4; ++p) for (int q = 0; q < 4; ++q) H[p * 4 + q] += j[p] * j[q];
        }
    }
}

// inverse of an SPD 4x4 (torch.inverse(h) in pnp_uncert.py:77-78); returns 0 if not SPD.
int pnp_spd_inverse4(const double* H, double* inv) { return spd_inverse4(H, inv) ? 1 : 0; }

// ---- Known-answer tests of the minimiser: the two problems of the Ceres Solver tutorial
// (examples/helloworld.cc: f = 10 - x from x = 0.5;  examples/powell.cc: Powell's singular function from
// (3, -1, 0, 1), both with DENSE_QR and default options), whose minimizer_progress_to_stdout tables are printed in
// docs/source/nnls_tutorial.rst.  trace: up to max_rows x 7 doubles (iteration, cost, cost_change, |gradient|,
// |step|, tr_ratio, tr_radius); returns the number of rows.  x_final: 1 / 4 doubles; summary: termination,
// iterations, final gradient max norm.
int ceres_kat_hello_world(double* trace, int max_rows, double* x_final, double* summary) {
    auto eval = [](const double* x, double* cost, double* res, double* jac, double* grad) {
        const double r = 10.0 - x[0];
        *cost = 0.5 * r * r;
        if (res) res[0] = r;
        if (jac) jac[0] = -1.0;
        if (grad) grad[0] = -r;
        return true;
    };
    std::vector<TraceRow> rows;
    double x[1] = {0.5};
    const LMResult r = trust_region_lm_n<1>(eval, 1, x, LMOptions(), &rows);
    const int n = std::min<int>(max_rows, rows.size());
    for (int i = 0; i < n; ++i) std::memcpy(trace + i * 7, rows[i].v, sizeof(rows[i].v));
    x_final[0] = x[0];
    summary[0] = r.term; summary[1] = r.iterations; summary[2] = rows.back().v[3];
    return n;
}

int ceres_kat_powell(double* trace, int max_rows, double* x_final, double* summary) {
    auto eval = [](const double* x, double* cost, double* res, double* jac, double* grad) {
        const double s5 = std::sqrt(5.0), s10 = std::sqrt(10.0);
        const double f[4] = {x[0] + 10.0 * x[1], s5 * (x[2] - x[3]), (x[1] - 2.0 * x[2]) * (x[1] - 2.0 * x[2]),
                             s10 * (x[0] - x[3]) * (x[0] - x[3])};
        *cost = 0.5 * (f[0] * f[0] + f[1] * f[1] + f[2] * f[2] + f[3] * f[3]);
        if (res) std::memcpy(res, f, sizeof(f));
        if (jac) {
            const double a = x[1] - 2.0 * x[2], b = x[0] - x[3];
            const double J[16] = {1.0, 10.0, 0.0, 0.0,
                                  0.0, 0.0, s5, -s5,
                                  0.0, 2.0 * a, -4.0 * a, 0.0,
                                  2.0 * s10 * b, 0.0, 0.0, -2.0 * s10 * b};
            std::memcpy(jac, J, sizeof(J));
            if (grad) for (int k = 0; k < 4; ++k) grad[k] = J[k] * f[0] + J[4 + k] * f[1] + J[8 + k] * f[2] + J[12 + k] * f[3];
        }
        return true;
    };
    std::vector<TraceRow> rows;
    double x[4] = {3.0, -1.0, 0.0, 1.0};
    const LMResult r = trust_region_lm_n<4>(eval, 4, x, LMOptions(), &rows);
    const int n = std::min<int>(max_rows, rows.size());
    for (int i = 0; i < n; ++i) std::memcpy(trace + i * 7, rows[i].v, sizeof(rows[i].v));
    std::memcpy(x_final, x, sizeof(x));
    summary[0] = r.term; summary[1] = r.iterations; summary[2] = rows.back().v[3];
    return n;
}

// A third, randomisable test problem for cross-checks between minimisers (tests/test_oracle.py): the 4-parameter
// curve fit  r_i = x0 exp(x1 t_i) + x2 + x3 t_i - y_i.  x: start in, result out; summary: termination, iterations,
// cost evaluations, final cost.
void ceres_lm_expfit(const double* t, const double* y, int m, double* x, double* summary) {
    auto eval = [t, y, m](const double* x, double* cost, double* res, double* jac, double* grad) {
        double acc = 0.0, g[4] = {0, 0, 0, 0};
        bool ok = true;
        for (int i = 0; i < m; ++i) {
            const double e = std::exp(x[1] * t[i]), r = x[0] * e + x[2] + x[3] * t[i] - y[i];
            acc += r * r;
            if (res) res[i] = r;
            if (jac) {
                const double J[4] = {e, x[0] * t[i] * e, 1.0, t[i]};
                for (int k = 0; k < 4; ++k) { jac[i * 4 + k] = J[k]; g[k] += J[k] * r; ok = ok && std::isfinite(J[k]); }
            }
        }
        *cost = 0.5 * acc;
        if (grad) std::memcpy(grad, g, sizeof(g));
        return ok && std::isfinite(acc);
    };
    const LMResult r = trust_region_lm_n<4>(eval, m, x, LMOptions());
    summary[0] = r.term; summary[1] = r.iterations; summary[2] = r.num_cost_evals; summary[3] = r.final_cost;
}

void pnp_oracle_set_adopt_candidate_on_ftol(int v) { g_options.adopt_candidate_on_ftol = v; }
int pnp_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
