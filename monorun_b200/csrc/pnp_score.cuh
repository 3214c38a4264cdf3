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

// Features of object i: x[17] (normalised) and the calibrated / corrected covariance cal[16].
__device__ __forceinline__ void pose_features_of(const ScoreParams& sp, int i, float x[17], float cal[16]) {
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
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) cal[a * 4 + b] = (cs[b] * cs[a]) * cov[a * 4 + b] * corr;
    // x = [yaw, t(3), tril(cov)(10, row-major lower triangle as torch.tril_indices(4, 4)), dims(3)]
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
}

__global__ void __launch_bounds__(128) pose_features_kernel(const ScoreParams sp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sp.n) return;
    float x[17], cal[16];
    pose_features_of(sp, i, x, cal);
    if (sp.cov_calib) {
#pragma unroll
        for (int k = 0; k < 16; ++k) sp.cov_calib[(size_t)i * 16 + k] = cal[k];
    }
#pragma unroll
    for (int k = 0; k < 17; ++k) sp.feat[(size_t)i * 17 + k] = x[k];
}

// sigmoid, invalid -> 0, product with the 2-D score, [l, h, w, x, y, z, ry, score] row (monorun_roi_head.py:544-556, :612-613)
__device__ __forceinline__ void finish_score_of(float s, const float* __restrict__ rows, const float* __restrict__ dims,
                                                const float* __restrict__ det_scores, int pre_sigmoid,
                                                float* __restrict__ scores, float* __restrict__ bbox3d, int i) {
    const float* r = rows + (size_t)i * MRPNP_RESULT_STRIDE;
    if (pre_sigmoid) s = 1.f / (1.f + expf(-s));
    if (!(r[20] > 0.5f)) s = 0.f;                 // scores[~ret_val] = 0
    if (det_scores) s = det_scores[i] * s;        // mult_2d_score
    if (scores) scores[i] = s;
    if (bbox3d) {
        float* o = bbox3d + (size_t)i * 8;
        o[0] = dims[(size_t)i * 3 + 0]; o[1] = dims[(size_t)i * 3 + 1]; o[2] = dims[(size_t)i * 3 + 2];
        o[3] = r[1]; o[4] = r[2]; o[5] = r[3]; o[6] = r[0]; o[7] = s;
    }
}

// ------------------------------------------------------------------ the whole score stage in ONE launch
// MLPScoreHead.forward (mlp_score_head.py:94-115) in the shape every reference config uses -- one pose layer, fusion
// 'add', one fused layer: h1 = relu(W1 x + b1) + reg_fc_out;  h2 = relu(W2 h1 + b2);  logit = w3 . h2 + b3 -- between
// the feature build above and the score finish below.  For the <= 100 objects of an image (max_per_img) this is a few
// MFLOP: one CTA owns kScoreTile objects, keeps their h1 rows in shared memory and streams the weights once per tile
// (W2 is passed TRANSPOSED, [H1, H2], so that consecutive threads read consecutive addresses).
constexpr int kScoreTile = 8;
constexpr int kScoreThreads = 256;
struct MlpParams {
    const float* reg_fc_out;   // [N, H1] or NULL
    const float* w1;           // [H1, 17]
    const float* b1;           // [H1]
    const float* w2t;          // [H1, H2]  (fused_fcs[0].weight transposed)
    const float* b2;           // [H2]
    const float* w3;           // [H2]
    const float* b3;           // [1]
    int h1, h2;
    const float* det_scores;   // [N] or NULL
    int pre_sigmoid;
    float* logits;             // [N] or NULL (raw scores, for tests)
    float* scores;             // [N] or NULL
    float* bbox3d;             // [N, 8] or NULL
};

__global__ void __launch_bounds__(kScoreThreads) score_stage_kernel(const ScoreParams sp, const MlpParams mp) {
    extern __shared__ float sm[];
    float* xs = sm;                               // [tile][17 (+3 pad)]
    float* hs = sm + kScoreTile * 20;             // [tile][h1]
    float* red = hs + kScoreTile * mp.h1;         // [tile][warps]
    const int t = threadIdx.x, first = blockIdx.x * kScoreTile;
    const int cnt = min(kScoreTile, sp.n - first);
    if (t < cnt) {
        float x[17], cal[16];
        pose_features_of(sp, first + t, x, cal);
#pragma unroll
        for (int k = 0; k < 17; ++k) xs[t * 20 + k] = x[k];
        if (sp.cov_calib) {
#pragma unroll
            for (int k = 0; k < 16; ++k) sp.cov_calib[(size_t)(first + t) * 16 + k] = cal[k];
        }
        if (sp.feat) {
#pragma unroll
            for (int k = 0; k < 17; ++k) sp.feat[(size_t)(first + t) * 17 + k] = x[k];
        }
    }
    __syncthreads();
    // pose layer + fusion: thread t owns the units t, t + 256, ...
    for (int j = t; j < mp.h1; j += kScoreThreads) {
        float w[17];
#pragma unroll
        for (int k = 0; k < 17; ++k) w[k] = __ldg(mp.w1 + (size_t)j * 17 + k);
        const float b = __ldg(mp.b1 + j);
        for (int o = 0; o < kScoreTile; ++o) {
            float a = 0.f;
            if (o < cnt) {
                a = b;
#pragma unroll
                for (int k = 0; k < 17; ++k) a = fmaf(w[k], xs[o * 20 + k], a);
                a = fmaxf(a, 0.f);
                if (mp.reg_fc_out) a += __ldg(mp.reg_fc_out + (size_t)(first + o) * mp.h1 + j);
            }
            hs[o * mp.h1 + j] = a;   // rows past the last object of the tile: zeros
        }
    }
    __syncthreads();
    // fused layer + output layer: thread t owns the units t, t + 256, ... of h2 and its share of w3 . h2
    float part[kScoreTile];
#pragma unroll
    for (int o = 0; o < kScoreTile; ++o) part[o] = 0.f;
    for (int u = t; u < mp.h2; u += kScoreThreads) {
        float acc[kScoreTile];
#pragma unroll
        for (int o = 0; o < kScoreTile; ++o) acc[o] = 0.f;
#pragma unroll 4
        for (int k = 0; k < mp.h1; ++k) {
            const float w = __ldg(mp.w2t + (size_t)k * mp.h2 + u);
#pragma unroll
            for (int o = 0; o < kScoreTile; ++o) acc[o] = fmaf(w, hs[o * mp.h1 + k], acc[o]);
        }
        const float b = __ldg(mp.b2 + u), w3 = __ldg(mp.w3 + u);
#pragma unroll
        for (int o = 0; o < kScoreTile; ++o) part[o] = fmaf(w3, fmaxf(acc[o] + b, 0.f), part[o]);
    }
#pragma unroll
    for (int o = 0; o < kScoreTile; ++o) {
        float v = part[o];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if ((t & 31) == 0) red[o * (kScoreThreads / 32) + (t >> 5)] = v;
    }
    __syncthreads();
    if (t < cnt) {
        float logit = __ldg(mp.b3);
#pragma unroll
        for (int w = 0; w < kScoreThreads / 32; ++w) logit += red[t * (kScoreThreads / 32) + w];
        if (mp.logits) mp.logits[first + t] = logit;
        finish_score_of(logit, sp.rows, sp.dims, mp.det_scores, mp.pre_sigmoid, mp.scores, mp.bbox3d, first + t);
    }
}

__global__ void __launch_bounds__(128) finish_scores_kernel(const float* __restrict__ logits, const float* __restrict__ rows,
                                                            const float* __restrict__ dims, const float* __restrict__ det_scores,
                                                            int pre_sigmoid, float* __restrict__ scores,
                                                            float* __restrict__ bbox3d, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    finish_score_of(logits[i], rows, dims, det_scores, pre_sigmoid, scores, bbox3d, i);
}

}  // namespace mrpnp
