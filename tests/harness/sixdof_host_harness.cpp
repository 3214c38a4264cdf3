// tests/harness/sixdof_host_harness.cpp -- TEST CODE.  Compiles the __host__ __device__ part of
// monorun_b200/csrc/pnp_6dof.cuh + lm_dense.cuh with g++ and drives it with an emulated warp (32 lane accumulators
// over the strided point loop, then the kernel's xor-butterfly), so the CPU suite can check the solver logic the GPU
// kernel executes against the oracle.  Nothing in monorun_b200/ loads it.
// Second entry: the mixed kernel's controller (mr6::lm_advance, pnp_6dof_fast.cuh -- mrlm::minimize cut at its cost
// evaluation) against mrlm::minimize itself on identical numbers.
#include <cstring>

#include "pnp_6dof.cuh"
#include "pnp_6dof_fast.cuh"

namespace {

template <bool FULLW>
struct EmulatedWarpPass {
    mr6::Camera cam;
    const float *c3, *c2, *cw;
    const unsigned char* mask;
    int n_pts;

    template <bool JAC>
    void run(const double* x, double* acc) const {
        mr6::Pose6 ps;
        mr6::make_pose(x, &ps);
        constexpr int wc = FULLW ? 3 : 2;
        constexpr int n = JAC ? mr6::kNAcc : 1;
        double lanes[32][mr6::kNAcc];
        std::memset(lanes, 0, sizeof(lanes));
        for (int base = 0; base < n_pts; base += 32)
            for (int lane = 0; lane < 32; ++lane) {
                const int p = base + lane;
                if (p >= n_pts || (mask && !mask[p])) continue;
                mr6::add_point<FULLW, JAC>(cam, ps, c3[p * 3], c3[p * 3 + 1], c3[p * 3 + 2], c2[p * 2], c2[p * 2 + 1],
                                           cw[p * wc], cw[p * wc + 1], FULLW ? cw[p * wc + 2] : 0.0, lanes[lane]);
            }
        for (int i = 0; i < n; ++i) {
            double v[32], w[32];
            for (int l = 0; l < 32; ++l) v[l] = lanes[l][i];
            for (int m = 16; m > 0; m >>= 1) {
                for (int l = 0; l < 32; ++l) w[l] = v[l] + v[l ^ m];
                std::memcpy(v, w, sizeof(v));
            }
            acc[i] += v[0];
        }
    }

    void operator()(const double* x, bool jac, double* acc) const {
        if (jac) run<true>(x, acc); else run<false>(x, acc);
    }
};

template <bool FULLW>
void solve(const float* c3, const float* c2, const float* cw, const unsigned char* mask, const float* K,
           const float* uv_range, const float* init, int n_pts, double z_min, double* result) {
    EmulatedWarpPass<FULLW> pass;
    pass.cam.fx = K[0]; pass.cam.fy = K[4]; pass.cam.cx = K[2]; pass.cam.cy = K[5]; pass.cam.z_min = z_min;
    pass.cam.u_min = uv_range[0]; pass.cam.u_max = uv_range[1]; pass.cam.v_min = uv_range[2]; pass.cam.v_max = uv_range[3];
    pass.c3 = c3; pass.c2 = c2; pass.cw = cw; pass.mask = mask; pass.n_pts = n_pts;
    double x[mr6::kNP];
    for (int k = 0; k < mr6::kNP; ++k) x[k] = init[k];
    const mrlm::LMOptions opt = mrlm::default_options();
    const mrlm::LMResult r = mrlm::minimize<mr6::kNP>(pass, x, opt);
    const bool valid = (r.term == mrlm::kConvergence || r.term == mrlm::kNoConvergence);
    double acc[mr6::kNAcc] = {0}, cov[36];
    pass(x, true, acc);
    const bool spd = mr6::covariance(acc, cov);
    for (int k = 0; k < 6; ++k) result[k] = x[k];
    for (int i = 0; i < 36; ++i) result[6 + i] = (valid && spd) ? cov[i] : ((i % 7 == 0) ? 1.0 : 0.0);
    result[42] = (valid && spd) ? 1.0 : 0.0;
    result[43] = r.iterations; result[44] = r.final_cost; result[45] = r.cost_evals; result[46] = r.term; result[47] = 0.0;
}

}  // namespace

// Interleaved tensors as mrpnp_solve_6dof (weights: istd [N,P,2] or full [N,P,3]); mask [N,P] bytes or NULL;
// result [N,48].
extern "C" void sixdof_host_harness(const float* coords_3d, const float* coords_2d, const float* weights,
                                    const unsigned char* mask, const float* cam_mats, int cam_stride,
                                    const float* uv_range, int range_stride, const float* init, int n_obj, int n_pts,
                                    int full_w, double z_min, double* result) {
    const int wc = full_w ? 3 : 2;
    for (int b = 0; b < n_obj; ++b) {
        const float* c3 = coords_3d + (size_t)b * n_pts * 3;
        const float* c2 = coords_2d + (size_t)b * n_pts * 2;
        const float* cw = weights + (size_t)b * n_pts * wc;
        const unsigned char* m = mask ? mask + (size_t)b * n_pts : nullptr;
        if (full_w)
            solve<true>(c3, c2, cw, m, cam_mats + (size_t)b * cam_stride, uv_range + (size_t)b * range_stride,
                        init + b * 6, n_pts, z_min, result + b * 48);
        else
            solve<false>(c3, c2, cw, m, cam_mats + (size_t)b * cam_stride, uv_range + (size_t)b * range_stride,
                         init + b * 6, n_pts, z_min, result + b * 48);
    }
}

namespace {

// What the mixed kernel's evaluation hands its controller: the cost in fp64, the 27 sums rounded to fp32.
template <bool FULLW>
struct MixedNumbersPass {
    EmulatedWarpPass<FULLW> inner;
    void full(const double* x, double* acc) const {
        for (int i = 0; i < mr6::kNAcc; ++i) acc[i] = 0.0;
        inner(x, true, acc);
        for (int i = 1; i < mr6::kNAcc; ++i) acc[i] = (double)(float)acc[i];
    }
    void operator()(const double* x, bool jac, double* acc) const {
        double all[mr6::kNAcc];
        full(x, all);
        if (jac) for (int i = 0; i < mr6::kNAcc; ++i) acc[i] = all[i];
        else acc[0] = all[0];
    }
};

template <bool FULLW>
void solve_both(const MixedNumbersPass<FULLW>& pass, const float* init, double* out_min, double* out_adv) {
    const mrlm::LMOptions opt = mrlm::default_options();
    {   // mrlm::minimize
        MixedNumbersPass<FULLW> p = pass;
        double x[mr6::kNP];
        for (int k = 0; k < mr6::kNP; ++k) x[k] = init[k];
        const mrlm::LMResult r = mrlm::minimize<mr6::kNP>(p, x, opt);
        for (int k = 0; k < 6; ++k) out_min[k] = x[k];
        out_min[6] = r.iterations; out_min[7] = r.final_cost; out_min[8] = r.cost_evals; out_min[9] = r.term;
    }
    {   // the kernel's loop around lm_advance: evaluate into the stash, advance, evaluate the candidate ...
        mr6::LMState S;
        mr6::StashEntry stash[2];
        std::memset(&S, 0, sizeof(S));
        std::memset(stash, 0, sizeof(stash));
        S.opt = opt;
        for (int k = 0; k < mr6::kNP; ++k) S.x[k] = init[k];
        S.cur = 0;
        auto evaluate = [&](const double* x, int e) {
            double acc[mr6::kNAcc];
            pass.full(x, acc);
            stash[e].cost = acc[0];
            for (int i = 0; i < mr6::kNAcc - 1; ++i) stash[e].tot[i] = (float)acc[1 + i];
            for (int k = 0; k < mr6::kNP; ++k) stash[e].x[k] = x[k];
        };
        evaluate(S.x, 0);
        bool first = true;
        while (mr6::lm_advance(S, stash, first) == mr6::kCmdEvaluate) {
            evaluate(S.cand, S.cur ^ 1);
            first = false;
        }
        for (int k = 0; k < 6; ++k) out_adv[k] = S.best[k];
        out_adv[6] = S.iterations; out_adv[7] = S.final_cost; out_adv[8] = S.cost_evals; out_adv[9] = S.term;
        bool same = true;   // the covariance evaluation finds the returned pose in the current stash entry
        for (int k = 0; k < 6; ++k) same = same && (stash[S.cur].x[k] == S.best[k]);
        out_adv[10] = same ? 1.0 : 0.0;
    }
}

}  // namespace

// Interleaved tensors as sixdof_host_harness; out_minimize [N,10], out_advance [N,11]: pose(6), iterations, final_cost,
// cost_evals, termination (, returned pose found in the stash).
extern "C" void sixdof_controller_harness(const float* coords_3d, const float* coords_2d, const float* weights,
                                          const unsigned char* mask, const float* cam_mats, const float* uv_range,
                                          const float* init, int n_obj, int n_pts, int full_w, double z_min,
                                          double* out_minimize, double* out_advance) {
    const int wc = full_w ? 3 : 2;
    for (int b = 0; b < n_obj; ++b) {
        auto fill = [&](auto& pass) {
            const float* K = cam_mats;
            pass.inner.cam.fx = K[0]; pass.inner.cam.fy = K[4]; pass.inner.cam.cx = K[2]; pass.inner.cam.cy = K[5];
            pass.inner.cam.z_min = z_min;
            pass.inner.cam.u_min = uv_range[0]; pass.inner.cam.u_max = uv_range[1];
            pass.inner.cam.v_min = uv_range[2]; pass.inner.cam.v_max = uv_range[3];
            pass.inner.c3 = coords_3d + (size_t)b * n_pts * 3;
            pass.inner.c2 = coords_2d + (size_t)b * n_pts * 2;
            pass.inner.cw = weights + (size_t)b * n_pts * wc;
            pass.inner.mask = mask ? mask + (size_t)b * n_pts : nullptr;
            pass.inner.n_pts = n_pts;
        };
        if (full_w) {
            MixedNumbersPass<true> pass;
            fill(pass);
            solve_both<true>(pass, init + b * 6, out_minimize + b * 10, out_advance + b * 11);
        } else {
            MixedNumbersPass<false> pass;
            fill(pass);
            solve_both<false>(pass, init + b * 6, out_minimize + b * 10, out_advance + b * 11);
        }
    }
}
