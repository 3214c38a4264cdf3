// tests/harness/exact_hessian_host_harness.cpp -- TEST CODE.  Compiles the __host__ __device__ part of
// monorun_b200/csrc/pnp_exact_hessian.cuh (point functor, 4x4 inverse) with g++ so that the CPU suite can check it
// against the reference's own autograd result (tests/golden/exact_hessian_ref.npz).  Nothing in monorun_b200/
// loads it.
#include <cstddef>

#include "pnp_exact_hessian.cuh"

// Interleaved tensors [N,P,3], [N,P,2], [N,P,2] (istd), cam [N,9], uv_range [N,4], pose [N,4], mask [N,P] bytes or
// NULL.  hessian [N,16], inverse [N,16], ok [N].
extern "C" void exact_hessian_host_harness(const float* coords_3d, const float* coords_2d, const float* istd,
                                           const float* cam_mats, const float* uv_range, const float* pose,
                                           const unsigned char* mask, int n_obj, int n_pts, double z_min,
                                           double* hessian, double* inverse, int* ok) {
    for (int b = 0; b < n_obj; ++b) {
        const float* K = cam_mats + (size_t)b * 9;
        const float* rg = uv_range + (size_t)b * 4;
        mrxh::Camera cam;
        cam.fx = K[0]; cam.fy = K[4]; cam.cx = K[2]; cam.cy = K[5]; cam.z_min = z_min;
        cam.u_min = rg[0]; cam.u_max = rg[1]; cam.v_min = rg[2]; cam.v_max = rg[3];
        const double yaw = pose[b * 4], t[3] = {pose[b * 4 + 1], pose[b * 4 + 2], pose[b * 4 + 3]};
        const double sn = sin(yaw), cs = cos(yaw);
        double acc[10] = {0};
        for (int p = 0; p < n_pts; ++p) {
            if (mask && !mask[(size_t)b * n_pts + p]) continue;
            const size_t i = (size_t)b * n_pts + p;
            mrxh::add_point(cam, sn, cs, t, coords_3d[i * 3], coords_3d[i * 3 + 1], coords_3d[i * 3 + 2],
                            coords_2d[i * 2], coords_2d[i * 2 + 1], istd[i * 2], istd[i * 2 + 1], acc);
        }
        int k = 0;
        for (int i = 0; i < 4; ++i)
            for (int j = i; j < 4; ++j) { hessian[b * 16 + i * 4 + j] = acc[k]; hessian[b * 16 + j * 4 + i] = acc[k]; ++k; }
        ok[b] = mrxh::invert4(acc, inverse + b * 16) ? 1 : 0;
    }
}
