// tests/harness/noc_host_harness.cpp -- TEST CODE.  Compiles the __host__ __device__ part of
// monorun_b200/csrc/pnp_noc.cuh (point functor, Huber corrector, trust-region controller) with g++ and
// drives it with an emulated warp: 32 lane accumulators over the strided point loop, then the same
// xor-butterfly the kernel runs with shuffles.  It exists so that `pytest -m "not gpu"` can check the
// solver logic the GPU kernel executes against the oracle.  Nothing in monorun_b200/ loads it.
#include <cstring>
#include <vector>

#include "pnp_noc.cuh"

namespace {

template <bool FULLW>
struct EmulatedWarpPass {
    mrnoc::Camera cam;
    const float *c3, *c2, *cw;  // interleaved [P,3], [P,2], [P,2|3]
    const unsigned char* mask;  // [P] or NULL
    double logdim[3], logdim_wgt[3], delta;
    int n_pts;

    template <bool JAC>
    void run(const double* x, double* acc) const {
        const mrnoc::DimPose d = mrnoc::make_dimpose(x);
        constexpr int wc = FULLW ? 3 : 2;
        constexpr int n = JAC ? mrnoc::kNAcc : 1;
        double lanes[32][mrnoc::kNAcc];
        std::memset(lanes, 0, sizeof(lanes));
        for (int base = 0; base < n_pts; base += 32)
            for (int lane = 0; lane < 32; ++lane) {
                const int p = base + lane;
                if (p >= n_pts || (mask && !mask[p])) continue;
                mrnoc::add_point<FULLW, JAC>(cam, d, delta, c3[p * 3], c3[p * 3 + 1], c3[p * 3 + 2], c2[p * 2],
                                             c2[p * 2 + 1], cw[p * wc], cw[p * wc + 1], FULLW ? cw[p * wc + 2] : 0.0,
                                             lanes[lane]);
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
        mrnoc::add_dim_prior<JAC>(x, logdim, logdim_wgt, delta, acc);
    }

    void operator()(const double* x, bool jac, double* acc) const {
        if (jac) run<true>(x, acc); else run<false>(x, acc);
    }
};

template <bool FULLW>
void solve(const float* c3, const float* c2, const float* cw, const unsigned char* mask, const float* logdim,
           const float* logdim_wgt, const float* K, const float* uv_range, const float* init, int n_pts,
           double z_min, double delta, double* result) {
    EmulatedWarpPass<FULLW> pass;
    pass.cam.fx = K[0]; pass.cam.fy = K[4]; pass.cam.cx = K[2]; pass.cam.cy = K[5];
    pass.cam.z_min = z_min;
    pass.cam.u_min = uv_range[0]; pass.cam.u_max = uv_range[1]; pass.cam.v_min = uv_range[2]; pass.cam.v_max = uv_range[3];
    pass.c3 = c3; pass.c2 = c2; pass.cw = cw; pass.mask = mask; pass.n_pts = n_pts; pass.delta = delta;
    double x[mrnoc::kNP];
    for (int k = 0; k < 3; ++k) { pass.logdim[k] = logdim[k]; pass.logdim_wgt[k] = logdim_wgt[k]; }
    for (int k = 0; k < mrnoc::kNP; ++k) x[k] = init[k];
    const mrnoc::LMOptions opt = mrnoc::default_options();
    const mrnoc::LMResult r = mrlm::minimize<mrnoc::kNP>(pass, x, opt);
    for (int k = 0; k < mrnoc::kNP; ++k) result[k] = x[k];
    result[7] = (r.term == mrnoc::kConvergence || r.term == mrnoc::kNoConvergence) ? 1.0 : 0.0;
    result[8] = r.iterations; result[9] = r.final_cost; result[10] = r.cost_evals; result[11] = r.term;
}

}  // namespace

// Same tensors as mrpnp_solve_noc (interleaved layout), host pointers, one object after another.
// mask: [N,P] bytes or NULL.  result [N,12].
extern "C" void noc_host_harness(const float* coords_3d, const float* coords_2d, const float* weights,
                                 const unsigned char* mask, const float* logdim, const float* logdim_wgt,
                                 const float* cam_mats, int cam_stride, const float* uv_range, int range_stride,
                                 const float* init, int n_obj, int n_pts, int full_w, double z_min, double delta,
                                 double* result) {
    const int wc = full_w ? 3 : 2;
    for (int b = 0; b < n_obj; ++b) {
        const float* c3 = coords_3d + (size_t)b * n_pts * 3;
        const float* c2 = coords_2d + (size_t)b * n_pts * 2;
        const float* cw = weights + (size_t)b * n_pts * wc;
        const unsigned char* m = mask ? mask + (size_t)b * n_pts : nullptr;
        if (full_w)
            solve<true>(c3, c2, cw, m, logdim + b * 3, logdim_wgt + b * 3, cam_mats + (size_t)b * cam_stride,
                        uv_range + (size_t)b * range_stride, init + b * 7, n_pts, z_min, delta, result + b * 12);
        else
            solve<false>(c3, c2, cw, m, logdim + b * 3, logdim_wgt + b * 3, cam_mats + (size_t)b * cam_stride,
                         uv_range + (size_t)b * range_stride, init + b * 7, n_pts, z_min, delta, result + b * 12);
    }
}
