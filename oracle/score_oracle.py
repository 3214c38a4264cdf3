"""TEST INFRASTRUCTURE (never imported by the product path): numpy restatement of what the reference does between the
PnP op and the score head's first Linear layer, and after its last one.  fp64 throughout.

    pose_features_ref   covariance calibration          monorun/models/roi_heads/bbox_3d_heads/optimizers/
                                                        uncert_prop_pnp_optimizer.py:96-97
                        covariance correction           monorun/core/bbox_3d/proj_error_coder/
                                                        distance_invar_proj_error_coder.py:62-63, called at
                                                        monorun/models/roi_heads/monorun_roi_head.py:530-534
                        tril + concat + pose_norm       monorun/models/roi_heads/bbox_3d_heads/score_heads/
                                                        mlp_score_head.py:99-106, :177-178
    finish_scores_ref   sigmoid, invalid -> 0, x 2-D    monorun_roi_head.py:544-551
                        [l,h,w,x,y,z,ry,score] rows     monorun_roi_head.py:612-613

Parity unpinned by the reference (it has no tests); pinned here against a torch transcription of the same lines
(tests/test_score.py), which is how the reference itself computes them.
"""
import numpy as np

# torch.tril_indices(4, 4): row-major lower triangle
TRIL = [(a, b) for a in range(4) for b in range(a + 1)]


def pose_features_ref(rows, dims, cov_calib_logscale=None, cov_correction_sd=0.0, distance_z_depth=False,
                      use_calib=False, norm=None):
    """rows [N,24] (yaw,t | cov 4x4 | valid,...), dims [N,3]; norm = (mean, var, weight, bias, eps) or None.
    Returns (feat [N,17], pose_cov_calib [N,4,4])."""
    rows = np.asarray(rows, np.float64)
    n = rows.shape[0]
    pose, cov = rows[:, :4], rows[:, 4:20].reshape(n, 4, 4)
    s = np.exp(np.asarray(cov_calib_logscale, np.float64)) if cov_calib_logscale is not None else np.ones(4)
    cal = (s * s[:, None]) * cov                                        # uncert_prop_pnp_optimizer.py:96-97
    if cov_correction_sd > 0:
        t = pose[:, 1:4]
        dist = t[:, 2] if distance_z_depth else np.linalg.norm(t, axis=1)
        cal = cal * np.square(cov_correction_sd / dist)[:, None, None]  # distance_invar_proj_error_coder.py:62-63
    src = cal if use_calib else cov
    tril = np.stack([src[:, a, b] for a, b in TRIL], 1)                 # mlp_score_head.py:99-100
    x = np.concatenate([pose, tril, np.asarray(dims, np.float64)], 1)   # :101
    if norm is not None:                                                # :102-103 -> :177-178
        mean, var, weight, bias, eps = norm
        x = (x - mean) / np.sqrt(var + eps) * weight + bias
    return x, cal


def finish_scores_ref(logits, rows, dims, det_scores=None, pre_sigmoid=True):
    rows = np.asarray(rows, np.float64)
    s = np.asarray(logits, np.float64).copy()
    if pre_sigmoid:
        s = 1.0 / (1.0 + np.exp(-s))                                    # monorun_roi_head.py:544-545
    s[~(rows[:, 20] > 0.5)] = 0.0                                       # :546
    if det_scores is not None:
        s = np.asarray(det_scores, np.float64) * s                      # :548-550
    bbox = np.concatenate([np.asarray(dims, np.float64), rows[:, 1:4], rows[:, 0:1], s[:, None]], 1)  # :612-613
    return s, bbox
