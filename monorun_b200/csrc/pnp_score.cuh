// pnp_score.cuh -- the immediate consumer of (pose, covariance): what the reference does with ~15 small torch launches
// between the PnP op and the score head's first Linear layer, and after its last one.
//
//   mrpnp_pose_features   covariance calibration          uncert_prop_pnp_optimizer.py:96-97
//                         test-time covariance correction  distance_invar_proj_error_coder.py:62-63,
//                                                          monorun_roi_head.py:530-534
//                         lower triangle of the 4x4 cov, concatenation [yaw, t, tril(cov), dims] and the
//                         eval-mode pose_norm              mlp_score_head.py:99-106, :177-178
//   mrpnp_finish_scores   sigmoid, invalid objects -> 0, product with the 2-D score and the [l,h,w,x,y,z,ry,score]
//                         rows of get_bbox_3d_result       monorun_roi_head.py:544-556, :612-613
// One thread per object; the arrays are tiny (24 floats in, 17 + 16 floats out per object).
#pragma once
#include "pnp_device.cuh"

namespace mrpnp {

struct ScoreParams {
    const float* rows;      // [N,24] result rows of the solver
    const float* dims;      // [N,3] decoded dimensions
    const float* calib_logscale;  // cov_calib_logscale [4] (device) or NULL
    float corr_sd;          // scaling_denominator of the covariance correction, 0 = off
    int corr_z_depth;       // distance = t_z instead of |t| (UncertProjectionHead.distance_mode)
    int use_calib;          // features use the calibrated (and corrected) covariance (test_cfg.calib_scoring)
    const float* norm_mean; // pose_norm running_mean [17] or NULL
    const float* norm_var;
    const float* norm_weight;
    const float* norm_bias;
    float norm_eps;
    float* feat;            // [N,17]
    float* cov_calib;       // [N,16] or NULL
    int n;
};

__global__ void __launch_bounds__(128) pose_features_kernel(const ScoreParams sp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sp.n) return;
    const float* r = sp.rows + (size_t)i * MRPNP_RESULT_STRIDE;
    float pose[4], cov[16];
#pragma unroll
    for (int k = 0; k < 4; ++k) pose[k] = r[k];
#pragma unroll
    for (int k = 0; k < 16; ++k) cov[k] = r[4 + k];
    // pose_cov_calib = (s s^T) * pose_cov, then * (sd / distance)^2
    float corr = 1.f;
    if (sp.corr_sd > 0.f) {
        const float d = sp.corr_z_depth ? pose[3] : sqrtf(pose[1] * pose[1] + pose[2] * pose[2] + pose[3] * pose[3]);
        const float q = sp.corr_sd / d;
        corr = q * q;
    }
    float cs[4] = {1.f, 1.f, 1.f, 1.f};
    if (sp.calib_logscale) {
#pragma unroll
        for (int a = 0; a < 4; ++a) cs[a] = expf(__ldg(sp.calib_logscale + a));
    }
    float cal[16];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) cal[a * 4 + b] = (cs[b] * cs[a]) * cov[a * 4 + b] * corr;
    if (sp.cov_calib) {
#pragma unroll
        for (int k = 0; k < 16; ++k) sp.cov_calib[(size_t)i * 16 + k] = cal[k];
    }
    // x = [yaw, t(3), tril(cov)(10, row-major lower triangle as torch.tril_indices(4, 4)), dims(3)]
    float x[17];
    x[0] = pose[0]; x[1] = pose[1]; x[2] = pose[2]; x[3] = pose[3];
    int q = 4;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) x[q++] = sp.use_calib ? cal[a * 4 + b] : cov[a * 4 + b];
#pragma unroll
    for (int k = 0; k < 3; ++k) x[14 + k] = sp.dims[(size_t)i * 3 + k];
    if (sp.norm_mean) {  // BatchNormSmooth1D in eval mode: (x - mean) / sqrt(var + eps) * weight + bias
#pragma unroll
        for (int k = 0; k < 17; ++k)
            x[k] = (x[k] - sp.norm_mean[k]) / sqrtf(sp.norm_var[k] + sp.norm_eps) * sp.norm_weight[k] + sp.norm_bias[k];
    }
#pragma unroll
    for (int k = 0; k < 17; ++k) sp.feat[(size_t)i * 17 + k] = x[k];
}

__global__ void __launch_bounds__(128) finish_scores_kernel(const float* __restrict__ logits, const float* __restrict__ rows,
                                                            const float* __restrict__ dims, const float* __restrict__ det_scores,
                                                            int pre_sigmoid, float* __restrict__ scores,
                                                            float* __restrict__ bbox3d, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = rows + (size_t)i * MRPNP_RESULT_STRIDE;
    float s = logits[i];
    if (pre_sigmoid) s = 1.f / (1.f + expf(-s));
    if (!(r[20] > 0.5f)) s = 0.f;                 // scores[~ret_val] = 0
    if (det_scores) s = det_scores[i] * s;        // mult_2d_score
    if (scores) scores[i] = s;
    if (bbox3d) {                                 // [l, h, w, x, y, z, ry, score]
        float* o = bbox3d + (size_t)i * 8;
        o[0] = dims[(size_t)i * 3 + 0]; o[1] = dims[(size_t)i * 3 + 1]; o[2] = dims[(size_t)i * 3 + 2];
        o[3] = r[1]; o[4] = r[2]; o[5] = r[3]; o[6] = r[0]; o[7] = s;
    }
}

}  // namespace mrpnp
