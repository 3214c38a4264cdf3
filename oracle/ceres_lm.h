// ============================================================================
// oracle/ceres_lm.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The minimiser shared by the three CPU oracles (pnp_oracle.cpp, pnp_noc_oracle.cpp, pnp_6dof_oracle.cpp):
// ceres::internal::TrustRegionMinimizer + LevenbergMarquardtStrategy + DenseQRSolver of ceres-solver 1.14 with
// default Solver::Options, restated for one dense parameter block of N unknowns (Ceres is not vendored by the
// reference and cannot be built here).  Pinned to Ceres' own published known answers: run on the two problems of
// the Ceres tutorial it reproduces the printed iteration tables to every digit
// (tests/test_oracle.py::test_minimiser_reproduces_the_ceres_tutorial_tables).
// ============================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace ceres_lm {

// Default ceres::Solver::Options of 1.14 that the reference leaves untouched
// (the reference only sets linear_solver_type = DENSE_QR: pnp_uncert_cpu.cpp:270-271, :318-319, :361-362).
struct LMOptions {
    int max_num_iterations = 50;
    double function_tolerance = 1e-6;
    double gradient_tolerance = 1e-10;
    double parameter_tolerance = 1e-8;
    double initial_trust_region_radius = 1e4;
    double max_trust_region_radius = 1e16;
    double min_trust_region_radius = 1e-32;
    double min_relative_decrease = 1e-3;
    double min_lm_diagonal = 1e-6;
    double max_lm_diagonal = 1e32;
    int max_num_consecutive_invalid_steps = 5;
    // 0: Ceres 1.14 behaviour -- on the function-tolerance exit the candidate point is
    //    NOT adopted (TrustRegionMinimizer::Minimize returns before HandleSuccessfulStep).
    // 1: adopt the candidate on that exit when it lowers the cost (documented switch,
    //    SURVEY.md section 7 "hard parts").
    int adopt_candidate_on_ftol = 0;
};

enum Termination { CONVERGENCE = 0, NO_CONVERGENCE = 1, FAILURE = 2 };

// DenseQRSolver::SolveImpl (Ceres 1.14): least squares  min |[A; diag(D)] y - [b; 0]|
// by unpivoted Householder QR (Eigen householderQr().solve()).  A is m x n row-major,
// already column-scaled.  Returns false if y is not finite.
template <int n>
bool dense_qr_solve(const double* A, const double* b, const double* D, int m, double* y,
                    std::vector<double>& work) {
    const int M = m + n;
    work.resize(static_cast<size_t>(M) * (n + 1));
    double* W = work.data();  // M x (n+1): augmented [A | b ; D | 0]
    for (int i = 0; i < m; ++i) {
        for (int k = 0; k < n; ++k) W[i * (n + 1) + k] = A[i * n + k];
        W[i * (n + 1) + n] = b[i];
    }
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k <= n; ++k) W[(m + i) * (n + 1) + k] = 0.0;
        W[(m + i) * (n + 1) + i] = D[i];
    }
    for (int k = 0; k < n; ++k) {
        double tail = 0.0;
        for (int i = k + 1; i < M; ++i) tail += W[i * (n + 1) + k] * W[i * (n + 1) + k];
        const double c0 = W[k * (n + 1) + k];
        if (tail <= std::numeric_limits<double>::min()) continue;  // column already triangular
        double beta = std::sqrt(c0 * c0 + tail);
        if (c0 >= 0) beta = -beta;
        // v = [1; essential], essential = x_tail / (c0 - beta), tau = (beta - c0) / beta
        const double inv = 1.0 / (c0 - beta), tau = (beta - c0) / beta;
        for (int i = k + 1; i < M; ++i) W[i * (n + 1) + k] *= inv;
        W[k * (n + 1) + k] = beta;
        for (int col = k + 1; col <= n; ++col) {
            double dot = W[k * (n + 1) + col];
            for (int i = k + 1; i < M; ++i) dot += W[i * (n + 1) + k] * W[i * (n + 1) + col];
            dot *= tau;
            W[k * (n + 1) + col] -= dot;
            for (int i = k + 1; i < M; ++i) W[i * (n + 1) + col] -= dot * W[i * (n + 1) + k];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        double v = W[k * (n + 1) + n];
        for (int j = k + 1; j < n; ++j) v -= W[k * (n + 1) + j] * y[j];
        y[k] = v / W[k * (n + 1) + k];
    }
    for (int k = 0; k < n; ++k)
        if (!std::isfinite(y[k])) return false;
    return true;
}

struct LMResult {
    Termination term;
    int iterations;        // index of the last iteration summary pushed (Ceres numbering)
    int num_cost_evals;    // 1 (initial point) + candidate points evaluated
    int num_jac_evals;
    double final_cost;
    double tr_radius;      // summary.iterations.back().trust_region_radius (cpp:277)
};

// One row of Solver::Summary::iterations as minimizer_progress_to_stdout prints it:
// iteration, cost, cost_change, |gradient|_max, |step|, tr_ratio, tr_radius (successful steps and iteration 0).
struct TraceRow { double v[7]; };

// TrustRegionMinimizer::Minimize of Ceres 1.14 for one N-vector parameter block: no bounds, no inner
// iterations, monotonic steps, Jacobi scaling on, LM strategy, DENSE_QR.  x holds init on entry and the
// returned parameters on exit.  eval(x, &cost, res | NULL, jac | NULL, grad | NULL) -> evaluation is valid;
// m = number of residuals.  The PnP solver below instantiates it with N = 4; the known-answer tests at the
// end of this file (Ceres' own tutorial problems) with N = 1 and N = 4.
template <int N, class Eval>
LMResult trust_region_lm_n(const Eval& eval, int m, double* x_io, const LMOptions& opt,
                           std::vector<TraceRow>* trace = nullptr) {
    std::vector<double> res(m), jac(static_cast<size_t>(m) * N), model_res(m), work;
    double x[N], grad[N], scale[N], diag[N], lm_diag[N], step[N], delta[N], cand[N];
    std::memcpy(x, x_io, sizeof(x));
    LMResult out{FAILURE, 0, 0, 0, 0.0, opt.initial_trust_region_radius};

    double x_cost, cand_cost;
    double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
    bool reuse_diagonal = false;
    int num_invalid = 0;
    double minimum_cost = std::numeric_limits<double>::max();

    // ---- IterationZero -> EvaluateGradientAndJacobian(new_point) ----
    bool ok = eval(x, &x_cost, res.data(), jac.data(), grad);
    out.num_cost_evals++; out.num_jac_evals++;
    if (!ok) { out.final_cost = x_cost; return out; }  // FAILURE, parameters untouched
    {   // jacobi_scaling: 1 / (1 + sqrt(squared column norm)), from the initial Jacobian only
        double cn[N];
        for (int k = 0; k < N; ++k) cn[k] = 0.0;
        for (int i = 0; i < m; ++i) for (int k = 0; k < N; ++k) cn[k] += jac[i * N + k] * jac[i * N + k];
        for (int k = 0; k < N; ++k) scale[k] = 1.0 / (1.0 + std::sqrt(cn[k]));
    }
    auto scale_columns = [&]() {
        for (int i = 0; i < m; ++i) for (int k = 0; k < N; ++k) jac[i * N + k] *= scale[k];
    };
    scale_columns();
    auto max_norm = [](const double* g) {
        double v = 0; for (int k = 0; k < N; ++k) v = std::max(v, std::fabs(g[k])); return v; };
    auto norm_n = [](const double* v) {
        double s = 0; for (int k = 0; k < N; ++k) s += v[k] * v[k]; return std::sqrt(s); };
    double x_norm = norm_n(x);
    double gradient_max_norm = max_norm(grad);
    if (trace) trace->push_back(TraceRow{{0.0, x_cost, 0.0, gradient_max_norm, 0.0, 0.0, radius}});

    int iteration = 0;
    bool step_is_successful = true;  // iteration 0 counts as successful
    out.term = NO_CONVERGENCE;
    while (true) {
        // ---- FinalizeIterationAndCheckIfMinimizerCanContinue ----
        if (step_is_successful && x_cost < minimum_cost) {
            minimum_cost = x_cost;
            std::memcpy(x_io, x, sizeof(x));
        }
        out.tr_radius = radius;
        out.iterations = iteration;
        if (iteration >= opt.max_num_iterations) { out.term = NO_CONVERGENCE; break; }
        if (step_is_successful && gradient_max_norm <= opt.gradient_tolerance) { out.term = CONVERGENCE; break; }
        if (radius <= opt.min_trust_region_radius) { out.term = CONVERGENCE; break; }
        ++iteration;
        step_is_successful = false;

        // ---- ComputeTrustRegionStep -> LevenbergMarquardtStrategy::ComputeStep ----
        if (!reuse_diagonal) {
            for (int k = 0; k < N; ++k) diag[k] = 0.0;
            for (int i = 0; i < m; ++i) for (int k = 0; k < N; ++k) diag[k] += jac[i * N + k] * jac[i * N + k];
            for (int k = 0; k < N; ++k)
                diag[k] = std::min(std::max(diag[k], opt.min_lm_diagonal), opt.max_lm_diagonal);
        }
        for (int k = 0; k < N; ++k) lm_diag[k] = std::sqrt(diag[k] / radius);
        bool solved = dense_qr_solve<N>(jac.data(), res.data(), lm_diag, m, step, work);
        reuse_diagonal = true;
        bool step_is_valid = false;
        double model_cost_change = 0.0;
        if (solved) {
            for (int k = 0; k < N; ++k) step[k] = -step[k];
            double dot = 0.0;  // -(J step)^T (f + J step / 2)
            for (int i = 0; i < m; ++i) {
                double mr = 0.0;
                for (int k = 0; k < N; ++k) mr += jac[i * N + k] * step[k];
                dot += mr * (res[i] + mr / 2.0);
            }
            model_cost_change = -dot;
            step_is_valid = model_cost_change > 0.0;
        }
        if (!step_is_valid) {
            // ---- HandleInvalidStep ----
            if (++num_invalid >= opt.max_num_consecutive_invalid_steps) { out.term = FAILURE; break; }
            radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;  // StepIsInvalid
            continue;
        }
        num_invalid = 0;
        for (int k = 0; k < N; ++k) delta[k] = step[k] * scale[k];  // undo column scaling

        // ---- ComputeCandidatePointAndEvaluateCost ----
        for (int k = 0; k < N; ++k) cand[k] = x[k] + delta[k];
        if (!eval(cand, &cand_cost, nullptr, nullptr, nullptr))
            cand_cost = std::numeric_limits<double>::max();
        out.num_cost_evals++;

        // ---- ParameterToleranceReached ----
        const double step_norm = norm_n(delta);
        if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
            out.term = CONVERGENCE; break;
        }
        // ---- FunctionToleranceReached ----
        const double cost_change = x_cost - cand_cost;
        if (std::fabs(cost_change) <= opt.function_tolerance * x_cost) {
            if (opt.adopt_candidate_on_ftol && cand_cost < minimum_cost) {
                minimum_cost = cand_cost; x_cost = cand_cost;
                std::memcpy(x_io, cand, sizeof(cand));
            }
            out.term = CONVERGENCE; break;
        }
        // ---- IsStepSuccessful (monotonic TrustRegionStepEvaluator) ----
        const double relative_decrease = cost_change / model_cost_change;
        if (relative_decrease > opt.min_relative_decrease) {
            // ---- HandleSuccessfulStep ----
            std::memcpy(x, cand, sizeof(x));
            x_norm = norm_n(x);
            ok = eval(x, &x_cost, res.data(), jac.data(), grad);
            out.num_jac_evals++;
            if (!ok) { out.term = FAILURE; break; }
            scale_columns();
            gradient_max_norm = max_norm(grad);
            step_is_successful = true;
            // LevenbergMarquardtStrategy::StepAccepted
            radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
            radius = std::min(opt.max_trust_region_radius, radius);
            decrease_factor = 2.0;
            reuse_diagonal = false;
            if (trace) trace->push_back(TraceRow{{double(iteration), x_cost, cost_change, gradient_max_norm, step_norm,
                                                  relative_decrease, radius}});
        } else {
            // ---- HandleUnsuccessfulStep -> StepRejected ----
            radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
        }
    }
    out.final_cost = minimum_cost;
    return out;
}

}  // namespace ceres_lm
