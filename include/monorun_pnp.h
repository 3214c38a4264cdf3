/* monorun_pnp.h -- C ABI of libmonorun_pnp.so: batched uncertainty-weighted PnP on B200 (sm_100a).
 *
 * This is the drop-in boundary for MonoRUn's native least-squares op.  The reference binds ONE
 * per-object host function through cffi,
 *
 *     void pnp_uncert(double* pts2d, double* pts3d, double* wgt2d, double* K, double* init_pose,
 *                     int* result_val, double* result_pose, double* result_cov, double* result_tr,
 *                     int pn, double* clips);
 *         -- monorun/ops/least_squares/src/ext.h:1-13, called per object from
 *            monorun/ops/least_squares/pnp_uncert_cpu.py:102-106 inside the serial loop :180-191,
 *
 * after a device->host copy of every input (monorun/ops/least_squares/pnp_uncert.py:34-43).  The entry
 * points below replace that call *and* the loop around it with one batched launch on device-resident
 * tensors; mrpnp_solve_host keeps the reference's host-buffer calling convention for callers that have
 * not moved their data to the GPU.  Plain pointers and sizes only; no torch types.  All functions are
 * thread-safe; a context owns one device's scratch memory and may be used by one thread at a time.
 *
 * Per-object semantics (identical to the reference):
 *   residual / clip rules   pnp_uncert_cpu.cpp:24-51 (diag weights), :189-217 (full 2x2 weights)
 *   solver                  Ceres 1.14 trust-region Levenberg-Marquardt, default options, DENSE_QR
 *   istd inlier test        pnp_uncert_cpu.py:164-168, "<=4 inliers -> use all" :23-32
 *   weights from log-std    uncert_prop_pnp_optimizer.py:73
 *   pose covariance         inverse of approx_hessian, hessian.py:67-87 / pnp_uncert.py:71-85
 */
#ifndef MONORUN_PNP_H_
#define MONORUN_PNP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRPNP_VERSION 1
#define MRPNP_MAX_POINTS 1024   /* points per object (H*W); 28x28 = 784 in every reference config */
#define MRPNP_RESULT_STRIDE 24  /* floats per result row */
#define MRPNP_MAX_PEERS 8       /* GPUs of one NVLink / NVSwitch domain whose gathered buffers a launch can write */

/* tensor layout of the correspondence arrays */
#define MRPNP_LAYOUT_PLANAR 0      /* [N, C, P]  == [N, C, H, W] head-level tensors (monorun_roi_head.py:513-529) */
#define MRPNP_LAYOUT_INTERLEAVED 1 /* [N, P, C]  op-level tensors of PnPUncert.forward (pnp_uncert.py:125-142)    */

/* meaning of the `weights` array */
#define MRPNP_W_LOGSTD 0 /* C=2 log-std; istd = exp(-logstd)/std_scale (uncert_prop_pnp_optimizer.py:73) */
#define MRPNP_W_ISTD 1   /* C=2 inverse std per axis (coords_2d_istd of pnp_uncert.py:7)                   */
#define MRPNP_W_FULL 2   /* C=3 symmetric whitening matrix [wxx, wxy, wyy] (ext.h:33, .cpp:214-215)         */

/* arithmetic of the solver */
#define MRPNP_PREC_FP64 0  /* residual, Jacobian and normal equations in fp64: reproduces the fp64 reference decisions */
#define MRPNP_PREC_MIXED 1 /* fp64 residual/cost chain + fp32 Jacobian sums in every pass, fp64 4x4 solve              */
#define MRPNP_PREC_FAST 2  /* residuals evaluated once in fp64, then tracked incrementally: candidate evaluations are   */
                           /* packed-fp32 "delta" passes whose cost CHANGE is accurate to ~1e-6 of itself.  Objects the */
                           /* fp32 path must not decide (a point near a clip bound, an accept / function-tolerance      */
                           /* decision within the rounding band of its threshold -- a handful per 8192) are solved by   */
                           /* the exact fp64 routine inside the same launch, so the results follow the fp64 reference   */
                           /* decision for decision.  Needs inlier_opt_only = 1 and an even n_pts (otherwise the solve  */
                           /* runs as MIXED).  Default.                                                                  */

/* pose covariance written to the result row */
#define MRPNP_COV_NONE 0
#define MRPNP_COV_PIPELINE 1 /* inverse(J^T J) with jacobian.py:48-98 masks -- what pnp_uncert.py:71-85 returns */
#define MRPNP_COV_CERES 2    /* inverse(J^T J) with Ceres-Jet masks -- what pnp_uncert_cpu.cpp:279-291 returns  */

/* where LM starts */
#define MRPNP_INIT_GIVEN 0  /* init_pose[N,4] supplied by the caller (e.g. an EPnP result)                     */
#define MRPNP_INIT_LINEAR 1 /* on-device weighted linear 4-DoF initialiser (replaces cv2.solvePnP, .py:34-58) */

/* status codes */
#define MRPNP_OK 0
#define MRPNP_ERR_ARG (-1)
#define MRPNP_ERR_CUDA (-2)
#define MRPNP_ERR_ALIGN (-3)

typedef struct mrpnp_params {
    int32_t n_obj;          /* N objects                                                              */
    int32_t n_pts;          /* P points per object, 4 <= P <= MRPNP_MAX_POINTS                         */
    int32_t layout;         /* MRPNP_LAYOUT_*                                                          */
    int32_t weight_mode;    /* MRPNP_W_*                                                               */
    int32_t cam_stride;     /* 0: one 3x3 K for all objects, 9: K per object (cam_mats (N|1,3,3))      */
    int32_t range_stride;   /* 0: one [u_min,u_max,v_min,v_max] for all, 4: per object                 */
    int32_t precision;      /* MRPNP_PREC_*                                                            */
    int32_t cov_mode;       /* MRPNP_COV_*                                                             */
    int32_t init_mode;      /* MRPNP_INIT_*                                                            */
    int32_t inlier_opt_only;/* 1: LM sees inliers only (all reference configs), 0: all points          */
    int32_t max_iterations; /* 0: Ceres default 50; <0: no LM step, evaluate cost/covariance at init_pose  */
    int32_t adopt_candidate_on_ftol; /* 0: Ceres 1.14 (candidate dropped on the function-tolerance exit) */
    float z_min;            /* PnPUncert(z_min=0.5)                                                    */
    float std_scale;        /* UncertPropPnPOptimizer(std_scale=10); only for MRPNP_W_LOGSTD           */
    float istd_thres;       /* epnp_istd_thres (0.6); <= 0 disables the istd inlier test               */
    float reserved;
    /* Fused all-gather of the result rows (multi-GPU; objects are sharded contiguously over the ranks):
     * with n_peers > 0 the kernel stores the row of local object i into EVERY peer's gathered buffer at row
     * row_offset + i -- peer-to-peer stores over NVLink / NVSwitch from the solver's epilogue -- instead of into
     * `result` (which may then be NULL).  peer_results[r] = device pointer, valid on THIS device, of rank r's
     * [n_total, 24] float buffer (this rank's own buffer included); e.g. the buffer_ptrs of a torch symmetric-memory
     * allocation.  The rows are complete on all ranks once every rank's launch has finished: either follow the call
     * with a cross-rank barrier on the same stream, or let the kernel signal completion itself (below).
     *
     * Completion flags (no barrier, no rank waits for another rank's solve): with peer_flags[0] != NULL the LAST thread
     * block of the launch, after a system-scope fence, stores flag_value into peer_flags[r][flag_slot] of every rank r
     * (release store into that rank's flag array, `flag_slot` = this rank).  A consumer waits for all ranks' slots of ITS
     * array to reach the value (mrpnp_gather_wait) when it needs the rows, not before.  Write-after-read protection for
     * the next use of the same buffers: with acks != NULL the launch first waits until acks[r] >= ack_value for every
     * r < n_peers -- `acks` is THIS rank's array, which rank r's consumer advances with mrpnp_gather_wait(...,
     * peer_acks) once it has read the previous contents.  Both waits give up after about a second and count a
     * time-out (mrpnp_gather_timeouts) instead of hanging the device. */
    int32_t n_peers;
    int32_t flag_slot;
    int64_t row_offset;
    float* peer_results[MRPNP_MAX_PEERS];
    uint32_t* peer_flags[MRPNP_MAX_PEERS];
    uint32_t flag_value;
    uint32_t ack_value;
    const uint32_t* acks;
    /* MRPNP_PREC_FAST: half-widths of the bands around Ceres' decision thresholds (|cost change| = 1e-6 cost,
     * rho = 1e-3) inside which a decision is left to the exact fp64 routine.  band_first: error of the FIRST step's cost
     * change relative to the cost (two independently rounded fp32 sums); later steps: band_rel |change| +
     * band_mix sqrt(model change * cost).  mrpnp_default_params sets 8e-6, 4e-3, 2e-6; 0 disables a term. */
    float band_first, band_rel, band_mix;
    /* band_ratio > 0 narrows the later-step band for objects that converge slowly (at least band_ratio_from evaluations):
     * the deviation of the fp32 trajectory from the oracle's is about 1e-6 of the PREVIOUS step, so its effect on a cost
     * change, relative to that change, scales with |previous step| / |this step|: the band is multiplied by
     * clamp(band_ratio * that ratio, band_rel_min, band_rel) / band_rel.  These are the objects that are expensive to
     * solve twice.  mrpnp_default_params sets 2e-5, 1e-4, 10 (over 24 seeded data sets x 8192 objects: 295 instead of 344
     * hand-backs, 8 instead of 7 objects with an evaluation count different from the fp64 kernel's, the same 5 outside the
     * tolerance, mean launch 160 instead of 187 us); band_ratio = 0 switches this off. */
    float band_ratio, band_rel_min;
    int32_t band_ratio_from;
    /* Diagnostic, DEVICE pointer to [N] int32 or NULL: for every object the fp32 path handed to the exact routine,
     * reason | (evaluations so far << 8); reason 1 = a point near a clip bound, 2 = first-step band, 3 = function-
     * tolerance band, 4 = accept band.  Entries of other objects are left untouched. */
    int32_t* hand_back_log;
    /* Reprojection-threshold consensus after the start pose -- the deterministic counterpart of the inlier refinement
     * the reference gets from cv2.solvePnPRansac (pnp_uncert_cpu.py:34-51): with the start pose (on-device linear
     * initialiser or init_pose) as the model, an istd inlier whose reprojection error exceeds the threshold is dropped,
     * provided more than four points survive; LM, the covariance and inlier_out then see the survivors only, and the
     * linear initialiser is run again on them.  ransac_thres: DEVICE pointer to [N] thresholds in pixels
     * (epnp_ransac_thres of pnp_uncert.py:9; NULL = off).  mrpnp_solve_dense instead takes ransac_ratio
     * (epnp_ransac_thres_ratio, uncert_prop_pnp_optimizer.py:86-88: threshold = ratio * (v[last row] - v[first row]);
     * 0 = off).  Needs inlier_opt_only = 1. */
    float ransac_ratio;
    const float* ransac_thres;
} mrpnp_params;

typedef struct mrpnp_ctx mrpnp_ctx;

/* Fills `p` with the reference defaults (configs/kitti_multiclass.py:122-132) for n_obj x n_pts; planar layout,
 * log-std weights, caller-supplied init_pose, pipeline covariance, precision MRPNP_PREC_FAST. */
void mrpnp_default_params(mrpnp_params* p, int32_t n_obj, int32_t n_pts);

/* Creates a context on CUDA device `device` (scratch buffers, work counter, streams for the host path). */
int mrpnp_create(mrpnp_ctx** ctx, int device);
void mrpnp_destroy(mrpnp_ctx* ctx);

/* Batched solve on DEVICE pointers, asynchronous on `stream` (a cudaStream_t, may be NULL).
 *   coords_3d   [N,3,P] or [N,P,3] float   object-frame points
 *   coords_2d   [N,2,P] or [N,P,2] float   pixel observations
 *   weights     [N,2|3,P] or [N,P,2|3] float, meaning per weight_mode
 *   cam_mats    [N|1,3,3] float row-major (only fx,fy,cx,cy are used, as in pnp_uncert_cpu.cpp:265)
 *   uv_range    [N|1,4] float  u_min,u_max,v_min,v_max  (clips[1..4] of ext.h:12)
 *   init_pose   [N,4] float  yaw,tx,ty,tz (ignored for MRPNP_INIT_LINEAR, may be NULL then)
 *   inlier_in   [N,ceil(P/32)] uint32 or NULL: externally supplied inlier mask, packed (bit j of word k of an
 *               object = point 32k+j); skips the istd test
 *   result      [N,24] float: yaw,tx,ty,tz | cov 4x4 row-major | valid, lm_iterations, final_cost, tr_radius
 *   inlier_out  [N,ceil(P/32)] uint32 or NULL: inlier mask actually used, packed the same way
 *   result64    [N,8] double or NULL: yaw,tx,ty,tz,final_cost,tr_radius,cost_evals,termination (for parity tests)
 * Planar inputs must be 16-byte aligned with P % 4 == 0 to take the TMA path; otherwise a plain
 * coalesced-load path is used (same results). Returns MRPNP_OK or a negative status. */
int mrpnp_solve(mrpnp_ctx* ctx, const mrpnp_params* p,
                const float* coords_3d, const float* coords_2d, const float* weights,
                const float* cam_mats, const float* uv_range, const float* init_pose,
                const uint32_t* inlier_in,
                float* result, uint32_t* inlier_out, double* result64, void* stream);

/* Fused head -> PnP entry: takes the dense head's RAW outputs and does, in the kernel prologue, what the reference
 * does with ~15 torch launches between FCNNOCDecoder and the op:
 *   NOCCoder.decode (core/bbox_3d/coord_coder/noc_coder.py:50-73), DistanceInvarProjErrorCoder.decode_logstd with
 *   distance=None (core/bbox_3d/proj_error_coder/distance_invar_proj_error_coder.py:39-60), roi_align of the pixel
 *   grid (models/roi_heads/monorun_roi_head.py:521-523, bin centres), exp(-logstd)/std_scale and the NCHW permutes
 *   (uncert_prop_pnp_optimizer.py:73-84).
 *   noc_pred    [N,3,H,W] float  class-sliced NOC map (FCNNOCDecoder.slice_pred, fcn_noc_decoder.py:242-267)
 *   proj_logstd [N,2,H,W] float  class-sliced raw log-std
 *   rois        [N,4] float      x1,y1,x2,y2 of each detection box at the test scale
 *   labels      [N] int64 or NULL class ids, only read when dp->num_classes > 0
 *   dims        [N,3] float      decoded dimensions (l,h,w);  dims_var [N,3] float or NULL (epistemic variance)
 *   distance    [N] float or NULL  predicted distance (global_head.pred_distance); NULL = scaling_denominator
 * p->n_pts = H*W, p->layout / p->weight_mode are ignored (planar, log-std).  Other arguments as mrpnp_solve. */
typedef struct mrpnp_dense_params {
    float noc_mean[3];   /* NOCCoder.target_means (configs/kitti_multiclass.py:107)                               */
    float noc_std[3];    /* NOCCoder.target_stds                                                                  */
    float focal_gain;    /* ref_focal_y * epistemic_std_gain        (distance_invar_proj_error_coder.py:52)       */
    float scaling_denominator; /* ref_length * ref_focal_y * target_std   (:23)                                   */
    float distance_min;  /* clamp of the optional per-object distance (:41)                                       */
    int32_t roi_w;       /* W of the H x W map (28)                                                               */
    int32_t num_classes; /* 0: noc_pred / proj_logstd are class-sliced maps.  C > 0: fused slice_pred -- noc_pred  */
                         /* points at the head's full output all_pred [N, >= 5*C, H, W] (the flip half already      */
                         /* selected by the pointer), proj_logstd is ignored and `labels` picks channels            */
                         /* [3c,3c+3) and [3C+2c, 3C+2c+2) of each object (fcn_noc_decoder.py:242-267)              */
    int64_t pred_stride; /* floats between consecutive objects of all_pred (2*5*C*H*W for the flip-paired view)    */
    /* MultiClassNormDimCoder.decode in the prologue (dim_coder/multiclass_norm_dim_coder.py:28-36).  dim_means != NULL: */
    /* `dims` / `dims_var` of mrpnp_solve_dense are the ENCODED regression outputs, decoded per object with its label   */
    /* (required then): dims = dim * std_c + mean_c, dims_var = dim_var * std_c^2.  The decoded values are also written  */
    /* to dims_out / dims_var_out [N,3] when those are not NULL (the score stage and the result rows need them).         */
    const float* dim_means;   /* [n_dim_classes, 3] device, or NULL = dims are already decoded                          */
    const float* dim_stds;    /* [n_dim_classes, 3] device                                                             */
    int32_t n_dim_classes;    /* labels outside [0, n_dim_classes) fail the object (result_val 0)                       */
    float* dims_out;          /* [N,3] or NULL                                                                         */
    float* dims_var_out;      /* [N,3] or NULL                                                                         */
} mrpnp_dense_params;

int mrpnp_solve_dense(mrpnp_ctx* ctx, const mrpnp_params* p, const mrpnp_dense_params* dp,
                      const float* noc_pred, const float* proj_logstd, const float* rois,
                      const int64_t* labels, const float* dims, const float* dims_var, const float* distance,
                      const float* cam_mats, const float* uv_range, const float* init_pose,
                      float* result, uint32_t* inlier_out, void* stream);

/* Same contract on HOST pointers: copies inputs to the device in chunks overlapped with the solve,
 * copies `result` (and inlier_out) back, and returns when they are valid -- the calling convention of
 * the reference's CPU op (numpy buffers in, numpy buffers out; pnp_uncert_cpu.py:128-209). */
int mrpnp_solve_host(mrpnp_ctx* ctx, const mrpnp_params* p,
                     const float* coords_3d, const float* coords_2d, const float* weights,
                     const float* cam_mats, const float* uv_range, const float* init_pose,
                     const uint32_t* inlier_in,
                     float* result, uint32_t* inlier_out);

/* The immediate consumer of (pose, covariance), fused: covariance calibration (uncert_prop_pnp_optimizer.py:96-97),
 * test-time covariance correction cov * (sd / distance)^2 (distance_invar_proj_error_coder.py:62-63,
 * monorun_roi_head.py:530-534), the lower triangle + concatenation [yaw, t, tril(cov), dims] and the eval-mode
 * pose_norm that open MLPScoreHead.forward (mlp_score_head.py:99-106, :177-178).  One launch instead of ~12.
 *   rows [N,24] result rows; dims [N,3]; cov_calib_logscale [4] or NULL (DEVICE pointer, the nn.Parameter);
 *   cov_correction_sd: scaling_denominator, 0 = no correction; distance_z_depth: distance = t_z instead of |t|;
 *   use_calib: features take the calibrated covariance (test_cfg.calib_scoring);
 *   norm_mean/var/weight/bias [17] or all NULL; feat [N,17] out; cov_calib [N,16] out or NULL. */
int mrpnp_pose_features(mrpnp_ctx* ctx, const float* rows, const float* dims, const float* cov_calib_logscale,
                        float cov_correction_sd, int32_t distance_z_depth, int32_t use_calib,
                        const float* norm_mean, const float* norm_var, const float* norm_weight, const float* norm_bias,
                        float norm_eps, float* feat, float* cov_calib, int32_t n, void* stream);

/* After the score head's last Linear layer (monorun_roi_head.py:544-556, :612-613): sigmoid (pre_sigmoid != 0),
 * invalid objects -> 0, product with the 2-D detection score (det_scores, NULL = mult_2d_score off), and the
 * [l,h,w,x,y,z,ry,score] rows of get_bbox_3d_result.  scores [N] and/or bbox_3d [N,8] may be NULL. */
int mrpnp_finish_scores(mrpnp_ctx* ctx, const float* score_logits, const float* rows, const float* dims,
                        const float* det_scores, int32_t pre_sigmoid, float* scores, float* bbox_3d, int32_t n,
                        void* stream);

/* The whole score stage in ONE launch: mrpnp_pose_features -> MLPScoreHead's layers -> mrpnp_finish_scores, for the
 * network shape of every reference config (mlp_score_head.py:11-115 with num_pose_fcs = num_fused_fcs = 1, fusion 'add'):
 *   h1 = relu(W1 x + b1) + reg_fc_out;   h2 = relu(W2 h1 + b2);   logit = w3 . h2 + b3
 * w1 [H1,17] = pose_fcs[0].weight, w2t [H1,H2] = fused_fcs[0].weight TRANSPOSED, w3 [H2] = fc_out.weight, b3 [1];
 * reg_fc_out [N,H1] or NULL.  The first twelve arguments are those of mrpnp_pose_features, det_scores / pre_sigmoid /
 * scores / bbox_3d those of mrpnp_finish_scores; cov_calib [N,16] and logits [N] (raw scores) may be NULL. */
int mrpnp_score_stage(mrpnp_ctx* ctx, const float* rows, const float* dims, const float* cov_calib_logscale,
                      float cov_correction_sd, int32_t distance_z_depth, int32_t use_calib,
                      const float* norm_mean, const float* norm_var, const float* norm_weight, const float* norm_bias,
                      float norm_eps, const float* reg_fc_out, const float* w1, const float* b1, const float* w2t,
                      const float* b2, const float* w3, const float* b3, int32_t h1, int32_t h2,
                      const float* det_scores, int32_t pre_sigmoid, float* scores, float* bbox_3d, float* cov_calib,
                      float* logits, int32_t n, void* stream);

/* Class-wise 3-D NMS on bird's-eye-view rotated boxes (MonoRUnRoIHead.multiclass_3d_result_nms,
 * monorun_roi_head.py:619-680; mmdet3d nms_gpu on centre (x, z), extent (l, w), angle ry): per image and class, in
 * descending score order, a box is dropped if its rotated IoU with an earlier kept box exceeds iou_thr.
 *   bbox_3d [N,8] l,h,w,x,y,z,ry,score (mrpnp_finish_scores); labels [N] int64 or NULL (one class);
 *   group_offsets [G+1] int32 DEVICE: image g owns objects [off[g], off[g+1]); max_group: an upper bound of the
 *   largest group (<= 4096); keep [N] uint8 out (1 = kept). */
int mrpnp_nms_bev(mrpnp_ctx* ctx, const float* bbox_3d, const int64_t* labels, const int32_t* group_offsets,
                  int32_t n_groups, int32_t max_group, float iou_thr, uint8_t* keep, void* stream);

/* The reference's two 7-parameter solvers, batched on the device -- one call replaces N calls of
 *   pnp_noc_uncert      (monorun/ops/least_squares/src/ext.h:15-28, pnp_uncert_cpu.cpp:294-334)  weight_mode MRPNP_W_ISTD
 *   pnp_noc_cov_uncert  (monorun/ops/least_squares/src/ext.h:30-43, pnp_uncert_cpu.cpp:336-377)  weight_mode MRPNP_W_FULL
 * Unknowns per object: [log l, log h, log w, yaw, tx, ty, tz].  The points are NORMALISED object coordinates; the
 * solver scales them by exp(log dims), so the dimensions are refined together with the pose under a prior
 * r_k = logdim_wgt_k (x_k - logdim_k) (DimErrorArray, .cpp:77-104).  Every residual block (one per point, one for
 * the prior) goes through ceres::HuberLoss(huber_delta).  Ceres 1.14 default trust-region options, fp64 arithmetic.
 *   coords_3d [N,3,P]|[N,P,3], coords_2d [N,2,P]|[N,P,2], weights [N,2|3,P]|[N,P,2|3] float (layout as mrpnp_solve)
 *   logdim, logdim_wgt [N,3] float;  cam_mats [N|1,3,3];  uv_range [N|1,4];  init_dimpose [N,7] float
 *   inlier_in [N,ceil(P/32)] packed uint32 or NULL (all points; the reference passes the points it wants solved)
 *   result [N,12] DOUBLE: dimpose[7], valid (summary.IsSolutionUsable), lm_iterations, final_cost, cost_evals,
 *   termination (0 convergence, 1 no convergence, 2 failure). */
typedef struct mrpnp_noc_params {
    int32_t n_obj;
    int32_t n_pts;          /* 1 <= P <= MRPNP_MAX_POINTS */
    int32_t layout;         /* MRPNP_LAYOUT_* */
    int32_t weight_mode;    /* MRPNP_W_ISTD or MRPNP_W_FULL */
    int32_t cam_stride;     /* 0 or 9 */
    int32_t range_stride;   /* 0 or 4 */
    int32_t max_iterations; /* 0: Ceres default 50 */
    int32_t reserved;
    float z_min;            /* clips[0] */
    float huber_delta;      /* `delta` of ext.h:27 / :42 */
} mrpnp_noc_params;

int mrpnp_solve_noc(mrpnp_ctx* ctx, const mrpnp_noc_params* p,
                    const float* coords_3d, const float* coords_2d, const float* weights,
                    const float* logdim, const float* logdim_wgt,
                    const float* cam_mats, const float* uv_range, const float* init_dimpose,
                    const uint32_t* inlier_in, double* result, void* stream);

/* PnPUncert(forward_exact_hessian=True): the second-order pose Hessian of monorun/ops/least_squares/hessian.py:5-64
 * (autograd of J^T e through jacobian.py) at `pose`, and its inverse as the pose covariance (pnp_uncert.py:63-85).
 * Uses p->n_obj, n_pts, layout, weight_mode (MRPNP_W_LOGSTD or MRPNP_W_ISTD), cam_stride, range_stride, z_min,
 * std_scale.  Tensors as mrpnp_solve;
 *   pose [N, pose_stride] float: yaw,tx,ty,tz lead each row (pose_stride = 24 reads the result rows of mrpnp_solve)
 *   inlier_in packed mask or NULL (all points; pass mrpnp_solve's inlier_out)
 *   hessian [N,16] float out or NULL;  rows [N,24] result rows or NULL: the inverse is written into the covariance
 *   slots, and `valid` is cleared (covariance = identity) when the Hessian is not invertible. */
int mrpnp_exact_hessian(mrpnp_ctx* ctx, const mrpnp_params* p,
                        const float* coords_3d, const float* coords_2d, const float* weights,
                        const float* cam_mats, const float* uv_range,
                        const float* pose, int32_t pose_stride, const uint32_t* inlier_in,
                        float* hessian, float* rows, void* stream);

/* 6-DoF extension of the uncertainty PnP (the north star's "6-DoF LM normal-equation solve"): unknowns
 * [rx, ry, rz, tx, ty, tz], rotation as ceres::AngleAxisRotatePoint; residuals, clips, whitening, Ceres 1.14 LM and
 * covariance (J^T J)^-1 as in pnp_uncert (ext.h:1-13, pnp_uncert_cpu.cpp:245-292).  The reference has no 6-DoF solver
 * (its rotation vector is (0, yaw, 0), pnp_uncert_cpu.cpp:28; `use_6dof` is never read, pnp_uncert.py:11,98,122,142),
 * so this entry has no reference interface to replace: PnPUncert(use_6dof=True).forward_6dof calls it.
 * Uses p->n_obj, n_pts, layout, weight_mode (any MRPNP_W_*), cam_stride, range_stride, z_min, std_scale,
 * max_iterations, precision: MRPNP_PREC_FP64 -- everything in fp64 (pnp_6dof.cuh); MRPNP_PREC_MIXED or _FAST -- the
 * correspondences staged once in shared memory, fp64 cost chain, fp32 Jacobian and normal equations, the same Ceres
 * decisions (pnp_6dof_fast.cuh; 3-4 x faster, pose within ~5e-7 of the fp64 kernel, covariance within ~3e-4).  Tensors as mrpnp_solve; init_pose6 [N,6] float (e.g. (0, yaw, 0, t) from mrpnp_solve);
 * inlier_in packed mask or NULL (all points).
 *   result [N,48] DOUBLE: rvec(3), t(3) | cov 6x6 row-major | valid, lm_iterations, final_cost, cost_evals,
 *   termination, pad. */
int mrpnp_solve_6dof(mrpnp_ctx* ctx, const mrpnp_params* p,
                     const float* coords_3d, const float* coords_2d, const float* weights,
                     const float* cam_mats, const float* uv_range, const float* init_pose6,
                     const uint32_t* inlier_in, double* result, void* stream);

/* The reference's native entry point itself, GPU-backed: the signature of monorun/ops/least_squares/src/ext.h:1-13, so
 * that the reference's own binding (pnp_uncert_cpu.py:102-106: ``lib.pnp_uncert(...)`` on numpy fp64 buffers, one object
 * per call) can load this library unchanged.  Host pointers; clips = z_min, u_min, u_max, v_min, v_max; every point
 * handed in is used; result_cov (4x4, may be NULL) is Ceres' covariance (pnp_uncert_cpu.cpp:279-291); no return code:
 * success is *result_val != 0.  pn is limited to MRPNP_MAX_POINTS.  Runs on the current CUDA device, fp64 precision. */
void pnp_uncert(double* pts2d, double* pts3d, double* wgt2d, double* K, double* init_pose, int* result_val,
                double* result_pose, double* result_cov, double* result_tr, int pn, double* clips);

/* Number of kernel launches issued by this context since creation (bench.py's gpu_launches). */
int64_t mrpnp_launch_count(const mrpnp_ctx* ctx);

/* MRPNP_PREC_FAST: number of objects the fp32 path handed to the exact fp64 routine (a point near a clip bound, a
 * decision inside its rounding band) since the context was created.  Synchronises the device. */
int64_t mrpnp_handed_back_count(mrpnp_ctx* ctx);

/* Consumer side of the fused all-gather's completion flags, asynchronous on `stream` (one warp):
 *   flags != NULL: wait until flags[r] >= value for every r < n (this rank's flag array; slot r is written by rank r's
 *                  solver launch, see mrpnp_params.peer_flags) -- afterwards all n ranks' rows are in the local buffer;
 *   peer_acks != NULL: then store ack_value into peer_acks[r][ack_slot] for every r < n (release, system scope): "this
 *                  rank has finished reading", the signal mrpnp_params.acks waits for before the buffers are rewritten.
 * Call it with flags only before the consumer, with peer_acks only after it, or with both when nothing reads the rows
 * in between.  The wait gives up after about a second (mrpnp_gather_timeouts). */
int mrpnp_gather_wait(mrpnp_ctx* ctx, const uint32_t* flags, int32_t n, uint32_t value, uint32_t* const* peer_acks,
                      int32_t ack_slot, uint32_t ack_value, void* stream);

/* Number of flag / ack waits that timed out since the context was created (0 in a healthy run).  Synchronises. */
int64_t mrpnp_gather_timeouts(mrpnp_ctx* ctx);

/* Static facts about the solver kernel for the given problem: writes warps per CTA, CTAs, dynamic
 * shared memory bytes and whether the TMA path is taken into info[0..3]. */
int mrpnp_kernel_info(mrpnp_ctx* ctx, const mrpnp_params* p, int32_t info[4]);

int mrpnp_version(void);
/* Message of the last error on the calling thread ("" if none). */
const char* mrpnp_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MONORUN_PNP_H_ */
