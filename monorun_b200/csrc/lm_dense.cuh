// lm_dense.cuh -- ceres::internal::TrustRegionMinimizer (1.14) for ONE small dense parameter block, on the 1 + NP +
// NP(NP+1)/2 numbers a pass over the residual blocks reduces to: cost, gradient J^T r and the upper triangle of J^T J.
// Shared by the 7-parameter solvers (pnp_noc.cuh) and the 6-DoF solver (pnp_6dof.cuh).  Ceres solves each
// Levenberg-Marquardt step by Householder QR of [J S; sqrt(D/radius)]; here the same step comes from the normal
// equations (S J^T J S + D/radius) y = S J^T r by Cholesky in fp64 (agreement ~1e-10, tests/test_noc.py).
// Everything is __host__ __device__ scalar code: on the GPU every lane of a warp runs it redundantly (and
// warp-uniformly) on bit-identical reduced sums; tests/harness/ compiles it with g++ for the CPU suite.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define MRLM_HD __host__ __device__ __forceinline__
#define MRLM_HD_NOINLINE __host__ __device__ __noinline__
#else
#define MRLM_HD inline
#define MRLM_HD_NOINLINE inline
#endif

namespace mrlm {

// accumulator layout: [cost | gradient (NP) | upper triangle of J^T J, row-major]
template <int NP>
struct Layout {
    static constexpr int kNH = NP * (NP + 1) / 2;
    static constexpr int kNAcc = 1 + NP + kNH;
    static constexpr int kAccG = 1, kAccH = 1 + NP;
};

template <int NP>
MRLM_HD int tri(int a, int b) { return a * NP - a * (a - 1) / 2 + (b - a); }  // a <= b

// Cholesky solve of the symmetric NP x NP system A y = b (A full, row-major).  false when A is not
// numerically positive definite or y is not finite -- the step is then "invalid", like a failed
// DenseQRSolver::Solve.
template <int NP>
MRLM_HD_NOINLINE bool cholesky_solve(const double* A, const double* b, double* y) {
    double L[NP * NP];
    for (int i = 0; i < NP; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = A[i * NP + j];
            for (int k = 0; k < j; ++k) s -= L[i * NP + k] * L[j * NP + k];
            if (i == j) {
                if (!(s > 0.0) || !isfinite(s)) return false;
                L[i * NP + i] = sqrt(s);
            } else {
                L[i * NP + j] = s / L[j * NP + j];
            }
        }
    double z[NP];
    for (int i = 0; i < NP; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i * NP + k] * z[k];
        z[i] = s / L[i * NP + i];
    }
    for (int i = NP - 1; i >= 0; --i) {
        double s = z[i];
        for (int k = i + 1; k < NP; ++k) s -= L[k * NP + i] * y[k];
        y[i] = s / L[i * NP + i];
    }
    bool ok = true;
    for (int i = 0; i < NP; ++i) ok = ok && isfinite(y[i]);
    return ok;
}

// The same solve in two halves on caller-supplied storage (L: NP x NP, z: NP), fully unrolled (no index arithmetic or loop
// branches; the operations and their order are those of cholesky_solve): factor once, solve for several right-hand sides.
template <int NP>
MRLM_HD bool cholesky_factor(const double* A, double* L) {
#pragma unroll
    for (int i = 0; i < NP; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double s = A[i * NP + j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= L[i * NP + k] * L[j * NP + k];
            if (i == j) {
                if (!(s > 0.0) || !isfinite(s)) return false;
                L[i * NP + i] = sqrt(s);
            } else {
                L[i * NP + j] = s / L[j * NP + j];
            }
        }
    return true;
}

template <int NP>
MRLM_HD bool cholesky_backsolve(const double* L, const double* b, double* y, double* z) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= L[i * NP + k] * z[k];
        z[i] = s / L[i * NP + i];
    }
#pragma unroll
    for (int i = NP - 1; i >= 0; --i) {
        double s = z[i];
#pragma unroll
        for (int k = i + 1; k < NP; ++k) s -= L[k * NP + i] * y[k];
        y[i] = s / L[i * NP + i];
    }
    bool ok = true;
#pragma unroll
    for (int i = 0; i < NP; ++i) ok = ok && isfinite(y[i]);
    return ok;
}

enum Termination { kConvergence = 0, kNoConvergence = 1, kFailure = 2 };

struct LMResult {
    int term, iterations, cost_evals, jac_evals;
    double final_cost;
};

// ceres::Solver::Options defaults of 1.14 (the reference only chooses DENSE_QR: pnp_uncert_cpu.cpp:270-271,
// :318-319, :361-362).
struct LMOptions {
    int max_num_iterations;
    double function_tolerance, gradient_tolerance, parameter_tolerance;
    double initial_radius, max_radius, min_radius, min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
    int max_consecutive_invalid;
};

MRLM_HD LMOptions default_options() {
    LMOptions o;
    o.max_num_iterations = 50;
    o.function_tolerance = 1e-6; o.gradient_tolerance = 1e-10; o.parameter_tolerance = 1e-8;
    o.initial_radius = 1e4; o.max_radius = 1e16; o.min_radius = 1e-32;
    o.min_relative_decrease = 1e-3; o.min_lm_diagonal = 1e-6; o.max_lm_diagonal = 1e32;
    o.max_consecutive_invalid = 5;
    return o;
}

MRLM_HD bool all_finite(const double* acc, int n) {
    bool ok = true;
    for (int i = 0; i < n; ++i) ok = ok && isfinite(acc[i]);
    return ok;
}

// ceres::internal::TrustRegionMinimizer::Minimize (1.14) for one NP-vector parameter block: LM
// strategy, Jacobi scaling, monotonic steps, no bounds, no inner iterations.  `pass(x, jac, acc)`
// fills acc[0] (jac == false) or acc[0..kNAcc) (jac == true) for the parameter vector x.
// x_io: initial parameters in, best accepted parameters out (pnp_uncert_cpu.cpp:259 / :302 memcpy + in-place solve).
template <int NP, class Pass>
MRLM_HD_NOINLINE LMResult minimize(Pass& pass, double* x_io, const LMOptions& opt) {
    double x[NP], grad[NP], scale[NP], diag[NP], Hs[NP * NP], bs[NP], A[NP * NP], step[NP],
        delta[NP], cand[NP], acc[Layout<NP>::kNAcc];
    for (int k = 0; k < NP; ++k) x[k] = x_io[k];
    for (int i = 0; i < Layout<NP>::kNAcc; ++i) acc[i] = 0.0;
    LMResult out;
    out.term = kFailure; out.iterations = 0; out.cost_evals = 0; out.jac_evals = 0; out.final_cost = 0.0;
    double radius = opt.initial_radius, decrease_factor = 2.0, x_cost, x_norm, gradient_max_norm;
    bool reuse_diagonal = false;
    int num_invalid = 0;
    double minimum_cost = 1.7976931348623157e308;

    // the scaled Gauss-Newton system of the current point: Hs = S J^T J S, bs = S J^T r
    auto load_point = [&](bool first) {
        x_cost = acc[0];
        gradient_max_norm = 0.0;
        for (int k = 0; k < NP; ++k) {
            grad[k] = acc[Layout<NP>::kAccG + k];
            gradient_max_norm = fmax(gradient_max_norm, fabs(grad[k]));
        }
        if (first)  // Jacobi scaling from the initial Jacobian only: 1 / (1 + |column|)
            for (int k = 0; k < NP; ++k) scale[k] = 1.0 / (1.0 + sqrt(acc[Layout<NP>::kAccH + tri<NP>(k, k)]));
        for (int a = 0; a < NP; ++a) {
            bs[a] = scale[a] * grad[a];
            for (int b = a; b < NP; ++b) {
                const double h = acc[Layout<NP>::kAccH + tri<NP>(a, b)] * scale[a] * scale[b];
                Hs[a * NP + b] = h;
                Hs[b * NP + a] = h;
            }
        }
        x_norm = 0.0;
        for (int k = 0; k < NP; ++k) x_norm += x[k] * x[k];
        x_norm = sqrt(x_norm);
    };

    pass(x, true, acc);
    out.cost_evals++; out.jac_evals++;
    if (!all_finite(acc, Layout<NP>::kNAcc)) { out.final_cost = acc[0]; return out; }
    load_point(true);

    int iteration = 0;
    bool step_is_successful = true;
    out.term = kNoConvergence;
    while (true) {
        if (step_is_successful && x_cost < minimum_cost) {
            minimum_cost = x_cost;
            for (int k = 0; k < NP; ++k) x_io[k] = x[k];
        }
        out.iterations = iteration;
        if (iteration >= opt.max_num_iterations) { out.term = kNoConvergence; break; }
        if (step_is_successful && gradient_max_norm <= opt.gradient_tolerance) { out.term = kConvergence; break; }
        if (radius <= opt.min_radius) { out.term = kConvergence; break; }
        ++iteration;
        step_is_successful = false;

        // LevenbergMarquardtStrategy::ComputeStep
        if (!reuse_diagonal)
            for (int k = 0; k < NP; ++k)
                diag[k] = fmin(fmax(Hs[k * NP + k], opt.min_lm_diagonal), opt.max_lm_diagonal);
        for (int i = 0; i < NP * NP; ++i) A[i] = Hs[i];
        for (int k = 0; k < NP; ++k) A[k * NP + k] += diag[k] / radius;
        const bool solved = cholesky_solve<NP>(A, bs, step);
        reuse_diagonal = true;
        bool step_is_valid = false;
        double model_cost_change = 0.0;
        if (solved) {
            double lin = 0.0, quad = 0.0;  // -(J s)^T (r + J s / 2) with s = -y
            for (int a = 0; a < NP; ++a) {
                step[a] = -step[a];
            }
            for (int a = 0; a < NP; ++a) {
                double hs = 0.0;
                for (int b = 0; b < NP; ++b) hs += Hs[a * NP + b] * step[b];
                lin += step[a] * bs[a];
                quad += step[a] * hs;
            }
            model_cost_change = -(lin + 0.5 * quad);
            step_is_valid = model_cost_change > 0.0;
        }
        if (!step_is_valid) {  // HandleInvalidStep
            if (++num_invalid >= opt.max_consecutive_invalid) { out.term = kFailure; break; }
            radius /= decrease_factor; decrease_factor *= 2.0;
            continue;
        }
        num_invalid = 0;
        double step_norm = 0.0;
        for (int k = 0; k < NP; ++k) {
            delta[k] = step[k] * scale[k];
            cand[k] = x[k] + delta[k];
            step_norm += delta[k] * delta[k];
        }
        step_norm = sqrt(step_norm);
        acc[0] = 0.0;
        pass(cand, false, acc);
        out.cost_evals++;
        const double cand_cost = isfinite(acc[0]) ? acc[0] : 1.7976931348623157e308;
        if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { out.term = kConvergence; break; }
        const double cost_change = x_cost - cand_cost;
        if (fabs(cost_change) <= opt.function_tolerance * x_cost) { out.term = kConvergence; break; }
        const double relative_decrease = cost_change / model_cost_change;
        if (relative_decrease > opt.min_relative_decrease) {  // HandleSuccessfulStep
            for (int k = 0; k < NP; ++k) x[k] = cand[k];
            for (int i = 0; i < Layout<NP>::kNAcc; ++i) acc[i] = 0.0;
            pass(x, true, acc);
            out.jac_evals++;
            if (!all_finite(acc, Layout<NP>::kNAcc)) { out.term = kFailure; break; }
            load_point(false);
            step_is_successful = true;
            const double q = 2.0 * relative_decrease - 1.0;  // LevenbergMarquardtStrategy::StepAccepted
            radius = radius / fmax(1.0 / 3.0, 1.0 - q * q * q);
            radius = fmin(opt.max_radius, radius);
            decrease_factor = 2.0;
            reuse_diagonal = false;
        } else {  // StepRejected
            radius /= decrease_factor; decrease_factor *= 2.0;
        }
    }
    out.final_cost = minimum_cost;
    return out;
}

}  // namespace mrlm
