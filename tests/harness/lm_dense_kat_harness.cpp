// tests/harness/lm_dense_kat_harness.cpp -- TEST CODE.  Runs the trust-region controller the fp64 CUDA kernels
// execute (monorun_b200/csrc/lm_dense.cuh: normal equations + Cholesky) on the two Ceres tutorial problems
// (examples/helloworld.cc, examples/powell.cc) under g++ and logs every Jacobian evaluation -- i.e. the initial
// point and each accepted step -- so tests/test_oracle.py can compare cost, |gradient| and |step| per iteration with
// the tables printed in Ceres' documentation.  Nothing in monorun_b200/ loads it.
#include <cmath>
#include <cstring>
#include <vector>

#include "lm_dense.cuh"

namespace {

template <int NP>
struct Logged {
    std::vector<double> rows;  // per Jacobian evaluation: cost, |gradient|_max, x[NP]
    void log(const double* x, const double* acc) {
        double g = 0.0;
        for (int k = 0; k < NP; ++k) g = std::fmax(g, std::fabs(acc[mrlm::Layout<NP>::kAccG + k]));
        rows.push_back(acc[0]);
        rows.push_back(g);
        for (int k = 0; k < NP; ++k) rows.push_back(x[k]);
    }
};

struct HelloWorld : Logged<1> {
    void operator()(const double* x, bool jac, double* acc) {
        const double r = 10.0 - x[0];
        acc[0] += 0.5 * r * r;
        if (jac) { acc[1] += -r; acc[2] += 1.0; log(x, acc); }
    }
};

struct Powell : Logged<4> {
    void operator()(const double* x, bool jac, double* acc) {
        const double s5 = std::sqrt(5.0), s10 = std::sqrt(10.0);
        const double a = x[1] - 2.0 * x[2], b = x[0] - x[3];
        const double f[4] = {x[0] + 10.0 * x[1], s5 * (x[2] - x[3]), a * a, s10 * b * b};
        for (int i = 0; i < 4; ++i) acc[0] += 0.5 * f[i] * f[i];
        if (!jac) return;
        const double J[4][4] = {{1.0, 10.0, 0.0, 0.0}, {0.0, 0.0, s5, -s5}, {0.0, 2.0 * a, -4.0 * a, 0.0},
                                {2.0 * s10 * b, 0.0, 0.0, -2.0 * s10 * b}};
        for (int i = 0; i < 4; ++i)
            for (int p = 0; p < 4; ++p) {
                acc[mrlm::Layout<4>::kAccG + p] += J[i][p] * f[i];
                for (int q = p; q < 4; ++q) acc[mrlm::Layout<4>::kAccH + mrlm::tri<4>(p, q)] += J[i][p] * J[i][q];
            }
        log(x, acc);
    }
};

template <int NP, class Pass>
int run(Pass& pass, const double* x0, double* rows, int max_rows, double* x_final, double* summary) {
    double x[NP];
    std::memcpy(x, x0, sizeof(x));
    const mrlm::LMResult r = mrlm::minimize<NP>(pass, x, mrlm::default_options());
    const int stride = 2 + NP, n = std::min<int>(max_rows, pass.rows.size() / stride);
    std::memcpy(rows, pass.rows.data(), sizeof(double) * n * stride);
    std::memcpy(x_final, x, sizeof(x));
    summary[0] = r.term; summary[1] = r.iterations; summary[2] = r.final_cost;
    return n;
}

}  // namespace

// rows: [k, 2 + NP] = cost, |gradient|_max, x.  Returns k.
extern "C" int lm_dense_kat_hello_world(double* rows, int max_rows, double* x_final, double* summary) {
    HelloWorld p;
    const double x0[1] = {0.5};
    return run<1>(p, x0, rows, max_rows, x_final, summary);
}

extern "C" int lm_dense_kat_powell(double* rows, int max_rows, double* x_final, double* summary) {
    Powell p;
    const double x0[4] = {3.0, -1.0, 0.0, 1.0};
    return run<4>(p, x0, rows, max_rows, x_final, summary);
}

namespace {

struct ExpFit {
    const double *t, *y;
    int m, cost_evals;
    void operator()(const double* x, bool jac, double* acc) {
        ++cost_evals;
        for (int i = 0; i < m; ++i) {
            const double e = std::exp(x[1] * t[i]), r = x[0] * e + x[2] + x[3] * t[i] - y[i];
            acc[0] += 0.5 * r * r;
            if (!jac) continue;
            const double J[4] = {e, x[0] * t[i] * e, 1.0, t[i]};
            for (int p = 0; p < 4; ++p) {
                acc[mrlm::Layout<4>::kAccG + p] += J[p] * r;
                for (int q = p; q < 4; ++q) acc[mrlm::Layout<4>::kAccH + mrlm::tri<4>(p, q)] += J[p] * J[q];
            }
        }
    }
};

}  // namespace

// The kernels' controller on r_i = x0 exp(x1 t_i) + x2 + x3 t_i - y_i.  summary: termination, iterations,
// cost evaluations (as Ceres counts them), final cost.
extern "C" void lm_dense_expfit(const double* t, const double* y, int m, double* x, double* summary) {
    ExpFit p{t, y, m, 0};
    const mrlm::LMResult r = mrlm::minimize<4>(p, x, mrlm::default_options());
    summary[0] = r.term; summary[1] = r.iterations; summary[2] = r.cost_evals; summary[3] = r.final_cost;
}
