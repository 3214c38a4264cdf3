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


# ------------------------------------------------------------------ class-wise rotated BEV NMS
# MonoRUnRoIHead.multiclass_3d_result_nms (monorun_roi_head.py:619-655) with boxes_for_nms =
# xywhr2xyxyr(bbox_3d[:, [3, 5, 0, 2, 6]]) (:660-680) handed to mmdet3d.ops.iou3d.nms_gpu.  mmdet3d is NOT in the
# reference tree or this container; its nms_gpu (mmdet3d 0.6-0.8, from OpenPCDet's iou3d_nms) sorts by descending
# score and drops a box whose rotated-rectangle IoU in the (x, z) plane with an earlier kept box is > thresh.  The
# IoU is restated here as exact polygon clipping in fp64 (parity unpinned against mmdet3d's own fp32 routine).
def _corners(cx, cz, l, w, ry):
    c, s = np.cos(ry), np.sin(ry)
    u = np.array([l / 2, -l / 2, -l / 2, l / 2]); v = np.array([w / 2, w / 2, -w / 2, -w / 2])
    return np.stack([cx + c * u + s * v, cz - s * u + c * v], 1)


def _clip(poly, a, b):
    """Sutherland-Hodgman: keep the part of `poly` on the left of the directed line a -> b."""
    out = []
    n = len(poly)
    for i in range(n):
        p, q = poly[i], poly[(i + 1) % n]
        sp = (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
        sq = (b[0] - a[0]) * (q[1] - a[1]) - (b[1] - a[1]) * (q[0] - a[0])
        if sp >= 0:
            out.append(p)
        if (sp >= 0) != (sq >= 0):
            t = sp / (sp - sq)
            out.append(p + t * (q - p))
    return out


def bev_iou_ref(b1, b2):
    """b = (l, h, w, x, y, z, ry, ...) rows of get_bbox_3d_result."""
    p1 = _corners(b1[3], b1[5], b1[0], b1[2], b1[6])
    p2 = _corners(b2[3], b2[5], b2[0], b2[2], b2[6])
    def area(p):
        p = np.asarray(p)
        return 0.5 * (p[:, 0] * np.roll(p[:, 1], -1) - np.roll(p[:, 0], -1) * p[:, 1]).sum()
    if area(p2) < 0:
        p2 = p2[::-1]
    poly = [p for p in p1]
    for i in range(4):
        poly = _clip(poly, p2[i], p2[(i + 1) % 4])
        if not poly:
            return 0.0
    inter = abs(area(poly))
    return inter / max(b1[0] * b1[2] + b2[0] * b2[2] - inter, 1e-12)


def nms_bev_ref(bbox_3d, labels=None, group_offsets=None, iou_thr=0.25):
    """Returns (keep [N] bool, min |IoU - thr| over the pairs that were compared)."""
    b = np.asarray(bbox_3d, np.float64)
    n = b.shape[0]
    labels = np.zeros(n, np.int64) if labels is None else np.asarray(labels)
    group_offsets = [0, n] if group_offsets is None else list(group_offsets)
    keep = np.zeros(n, bool)
    margin = np.inf
    for lo, hi in zip(group_offsets[:-1], group_offsets[1:]):
        for c in np.unique(labels[lo:hi]):
            ids = lo + np.flatnonzero(labels[lo:hi] == c)
            order = ids[np.argsort(-b[ids, 7], kind='stable')]
            kept = []
            for i in order:
                ok = True
                for j in kept:
                    iou = bev_iou_ref(b[j], b[i])
                    margin = min(margin, abs(iou - iou_thr))
                    if iou > iou_thr:
                        ok = False
                        break
                if ok:
                    kept.append(i)
            keep[kept] = True
    return keep, margin
