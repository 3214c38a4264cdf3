// ============================================================================
// oracle/pnp_noc_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU fp64 restatement of the reference's two 7-parameter solvers
//   pnp_noc_uncert      (monorun/ops/least_squares/src/pnp_uncert_cpu.cpp:294-334, ext.h:15-28)
//   pnp_noc_cov_uncert  (monorun/ops/least_squares/src/pnp_uncert_cpu.cpp:336-377, ext.h:30-43)
// unknowns [log l, log h, log w, yaw, tx, ty, tz]; one 2-residual block per point
// (NocReprojectionErrorArray :122-148 / NocCovReprojectionErrorArray :189-217), one 3-residual
// block for the dimension prior (DimErrorArray :77-104), every block under the same
// ceres::HuberLoss(delta).  Used only as the parity checker by tests/.
//
// PARITY UNPINNED: the reference exports these two functions but nothing in it calls them
// (no Python binding, config, test or golden vector), and ceres-solver 1.14 cannot be built here
// (see the header of pnp_oracle.cpp).  The residual functors follow the reference line by line; the
// minimiser is the one of oracle/ceres_lm.h -- the template that also solves the 4-parameter problems and
// reproduces Ceres' published tutorial tables; the loss enters through a restatement of
// ResidualBlock::Evaluate + Corrector.  tests/test_noc.py checks it (a) against pnp_oracle.cpp with the dimensions
// pinned by a stiff prior and the loss switched off, and (b) against scipy.optimize.minimize on the
// same robustified objective.
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

constexpr int N = 7;


struct Problem {
    const double* pts2d;   // n,2
    const double* pts3d;   // n,3  (normalised object coordinates)
    const double* wgt2d;   // n,2 (diag) or n,3 (wxx,wxy,wyy)
    const double* logdim;      // 3
    const double* logdim_wgt;  // 3
    int pn;
    bool full_w;
    double fx, fy, cx, cy;
    double z_min, u_min, u_max, v_min, v_max;
    double delta;          // HuberLoss a_; b_ = delta^2
};

// ceres::HuberLoss::Evaluate (loss_function.cc): rho(s), rho'(s) for s = |r_block|^2.
inline void huber(double a, double s, double* rho0, double* rho1) {
    const double b = a * a;
    if (s > b) {
        const double r = std::sqrt(s);
        *rho0 = 2.0 * a * r - b;
        *rho1 = std::max(std::numeric_limits<double>::min(), a / r);
    } else {
        *rho0 = s;
        *rho1 = 1.0;
    }
}

// One reprojection block with Ceres-Jet derivative semantics (selected branch of max() and of
// the clamping ternaries).  r: 2, jac: 2x7 row-major or NULL.  cpp:122-148 / :189-217.
inline void eval_point(const Problem& P, int i, const double* e /*exp(logdims)*/, double sn, double cs,
                       const double* t, double* r, double* jac) {
    const double Sx = P.pts3d[i * 3] * e[0], Sy = P.pts3d[i * 3 + 1] * e[1], Sz = P.pts3d[i * 3 + 2] * e[2];  // :125-127
    const double qx = cs * Sx + sn * Sz, qz = -sn * Sx + cs * Sz;    // AngleAxisRotatePoint((0,yaw,0)) :131
    const double xc = qx + t[0], yc = Sy + t[1], zc = qz + t[2];      // :132-134
    const bool z_free = !(zc < P.z_min);                              // :136
    const double z = z_free ? zc : P.z_min, iz = 1.0 / z;
    double pu = P.fx * xc * iz + P.cx, pv = P.fy * yc * iz + P.cy;    // :138-139
    bool u_free = true, v_free = true;
    if (pu < P.u_min) { pu = P.u_min; u_free = false; } else if (pu > P.u_max) { pu = P.u_max; u_free = false; }  // :141
    if (pv < P.v_min) { pv = P.v_min; v_free = false; } else if (pv > P.v_max) { pv = P.v_max; v_free = false; }  // :142
    const double du = pu - P.pts2d[i * 2], dv = pv - P.pts2d[i * 2 + 1];  // :144-145
    double w00, w01, w11;
    if (P.full_w) { w00 = P.wgt2d[i * 3]; w01 = P.wgt2d[i * 3 + 1]; w11 = P.wgt2d[i * 3 + 2]; }  // :214-215
    else { w00 = P.wgt2d[i * 2]; w01 = 0.0; w11 = P.wgt2d[i * 2 + 1]; }                          // :147-148
    r[0] = w00 * du + w01 * dv;
    r[1] = w01 * du + w11 * dv;
    if (!jac) return;
    const double mz = z_free ? 1.0 : 0.0;
    // d(x',y',z')/d(param): columns logl, logh, logw, yaw, tx, ty, tz
    const double dx[N] = {cs * Sx, 0.0, sn * Sz, qz, 1.0, 0.0, 0.0};
    const double dy[N] = {0.0, Sy, 0.0, 0.0, 0.0, 1.0, 0.0};
    const double dz[N] = {-sn * Sx, 0.0, cs * Sz, -qx, 0.0, 0.0, 1.0};
    const double au = u_free ? P.fx * iz : 0.0, bu = u_free ? -P.fx * xc * iz * iz * mz : 0.0;
    const double av = v_free ? P.fy * iz : 0.0, bv = v_free ? -P.fy * yc * iz * iz * mz : 0.0;
    for (int k = 0; k < N; ++k) {
        const double ju = au * dx[k] + bu * dz[k], jv = av * dy[k] + bv * dz[k];
        jac[k] = w00 * ju + w01 * jv;
        jac[N + k] = w01 * ju + w11 * jv;
    }
}

// ResidualBlock::Evaluate + Corrector (corrector.cc) for a block of `nr` residuals: cost
// contribution 1/2 rho(s); residuals and Jacobian rows scaled by sqrt(rho').  HuberLoss has
// rho'' <= 0 everywhere, so Corrector always takes its `rho[2] <= 0` branch (alpha = 0).
inline double robustify(double a, int nr, double* r, double* jac /*nr x N or NULL*/) {
    double s = 0.0;
    for (int i = 0; i < nr; ++i) s += r[i] * r[i];
    double rho0, rho1;
    huber(a, s, &rho0, &rho1);
    const double sc = std::sqrt(rho1);
    for (int i = 0; i < nr; ++i) r[i] *= sc;
    if (jac) for (int i = 0; i < nr * N; ++i) jac[i] *= sc;
    return 0.5 * rho0;
}

// Evaluator::Evaluate over the pn point blocks then the prior block (cpp:308-323 order).
// res: 2pn+3, jac: (2pn+3) x 7 row-major, grad: 7.  false when something is not finite.
bool evaluate(const Problem& P, const double* x, double* cost, double* res, double* jac, double* grad) {
    const double e[3] = {std::exp(x[0]), std::exp(x[1]), std::exp(x[2])};
    const double sn = std::sin(x[3]), cs = std::cos(x[3]);
    double acc = 0.0, g[N] = {0};
    bool ok = true;
    double rr[3], jj[3 * N];
    for (int i = 0; i <= P.pn; ++i) {
        const int nr = i < P.pn ? 2 : 3;
        if (i < P.pn) {
            eval_point(P, i, e, sn, cs, x + 4, rr, jac ? jj : nullptr);
        } else {  // DimErrorArray :87-93
            std::memset(jj, 0, sizeof(jj));
            for (int k = 0; k < 3; ++k) {
                rr[k] = P.logdim_wgt[k] * (x[k] - P.logdim[k]);
                jj[k * N + k] = P.logdim_wgt[k];
            }
        }
        acc += robustify(P.delta, nr, rr, jac ? jj : nullptr);
        for (int a = 0; a < nr; ++a) {
            ok = ok && std::isfinite(rr[a]);
            if (res) res[2 * i + a] = rr[a];
            if (jac) {
                for (int k = 0; k < N; ++k) {
                    jac[(2 * i + a) * N + k] = jj[a * N + k];
                    g[k] += jj[a * N + k] * rr[a];
                    ok = ok && std::isfinite(jj[a * N + k]);
                }
            }
        }
    }
    *cost = acc;
    if (grad) std::memcpy(grad, g, sizeof(g));
    return ok && std::isfinite(acc);
}

void solve_one(const double* pts2d, const double* pts3d, const double* wgt2d, const double* logdim,
               const double* logdim_wgt, const double* K, const double* init_dimpose, int* result_val,
               double* result_dimpose, int pn, const double* clips, double delta, bool full_w,
               int* stats, double* final_cost) {
    Problem P;
    P.pts2d = pts2d; P.pts3d = pts3d; P.wgt2d = wgt2d; P.logdim = logdim; P.logdim_wgt = logdim_wgt;
    P.pn = pn; P.full_w = full_w; P.delta = delta;
    P.fx = K[0]; P.fy = K[4]; P.cx = K[2]; P.cy = K[5];               // cpp:312
    P.z_min = clips[0]; P.u_min = clips[1]; P.u_max = clips[2]; P.v_min = clips[3]; P.v_max = clips[4];
    std::memcpy(result_dimpose, init_dimpose, N * sizeof(double));   // cpp:302
    LMOptions opt;
    auto eval = [&P](const double* x, double* cost, double* res, double* jac, double* grad) {
        return evaluate(P, x, cost, res, jac, grad);
    };
    const LMResult r = trust_region_lm_n<N>(eval, 2 * pn + 3, result_dimpose, opt);  // oracle/ceres_lm.h
    *result_val = (r.term == CONVERGENCE || r.term == NO_CONVERGENCE);  // IsSolutionUsable, cpp:333
    if (stats) { stats[0] = r.iterations; stats[1] = r.num_cost_evals; stats[2] = r.num_jac_evals; stats[3] = r.term; }
    if (final_cost) *final_cost = r.final_cost;
}

}  // namespace

extern "C" {

// Exact ABI of monorun/ops/least_squares/src/ext.h:15-28.
void pnp_noc_uncert(double* pts2d, double* pts3d, double* wgt2d, double* logdim, double* logdim_wgt, double* K,
                    double* init_dimpose, int* result_val, double* result_dimpose, int pn, double* clips,
                    double delta) {
    solve_one(pts2d, pts3d, wgt2d, logdim, logdim_wgt, K, init_dimpose, result_val, result_dimpose, pn, clips,
              delta, false, nullptr, nullptr);
}

// Exact ABI of monorun/ops/least_squares/src/ext.h:30-43.
void pnp_noc_cov_uncert(double* pts2d, double* pts3d, double* wgt2d, double* logdim, double* logdim_wgt,
                        double* K, double* init_dimpose, int* result_val, double* result_dimpose, int pn,
                        double* clips, double delta) {
    solve_one(pts2d, pts3d, wgt2d, logdim, logdim_wgt, K, init_dimpose, result_val, result_dimpose, pn, clips,
              delta, true, nullptr, nullptr);
}

// Batched driver for the parity tests: object b owns pn[b] points from point offset off[b] of the
// packed arrays; K [nb,9], clips [nb,5], logdim / logdim_wgt [nb,3], init / result [nb,7],
// stats [nb,4] = iterations, cost evals, jacobian evals, termination; cost [nb].
void pnp_noc_batch(const double* pts2d, const double* pts3d, const double* wgt2d, const double* logdim,
                   const double* logdim_wgt, const double* K, const double* init_dimpose, int* result_val,
                   double* result_dimpose, const int* pn, const long long* off, const double* clips, double delta,
                   int nb, int full_w, int* stats, double* cost, int threads) {
    const int wc = full_w ? 3 : 2;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 8) num_threads(threads)
#endif
    for (int b = 0; b < nb; ++b) {
        const long long o = off[b];
        solve_one(pts2d + o * 2, pts3d + o * 3, wgt2d + o * wc, logdim + b * 3, logdim_wgt + b * 3, K + b * 9,
                  init_dimpose + b * N, result_val + b, result_dimpose + b * N, pn[b], clips + b * 5, delta,
                  full_w != 0, stats ? stats + b * 4 : nullptr, cost ? cost + b : nullptr);
    }
}

// Robustified cost 1/2 sum rho(|r_block|^2), gradient (7) and Gauss-Newton matrix J^T J (7x7) of the
// corrected problem at `dimpose` -- what the LM loop sees.  For tests.
void pnp_noc_eval(const double* pts2d, const double* pts3d, const double* wgt2d, const double* logdim,
                  const double* logdim_wgt, const double* K, const double* dimpose, int pn, const double* clips,
                  double delta, int full_w, double* cost, double* grad, double* JtJ) {
    Problem P;
    P.pts2d = pts2d; P.pts3d = pts3d; P.wgt2d = wgt2d; P.logdim = logdim; P.logdim_wgt = logdim_wgt;
    P.pn = pn; P.full_w = full_w != 0; P.delta = delta;
    P.fx = K[0]; P.fy = K[4]; P.cx = K[2]; P.cy = K[5];
    P.z_min = clips[0]; P.u_min = clips[1]; P.u_max = clips[2]; P.v_min = clips[3]; P.v_max = clips[4];
    const int m = 2 * pn + 3;
    std::vector<double> res(m), jac(static_cast<size_t>(m) * N);
    evaluate(P, dimpose, cost, res.data(), jac.data(), grad);
    if (JtJ) {
        std::memset(JtJ, 0, N * N * sizeof(double));
        for (int i = 0; i < m; ++i)
            for (int a = 0; a < N; ++a) for (int b = 0; b < N; ++b) JtJ[a * N + b] += jac[i * N + a] * jac[i * N + b];
    }
}

}  // extern "C"
