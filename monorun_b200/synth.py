"""Seeded synthetic correspondence generator.

Produces KITTI-like objects and the tensors that sit at the head->PnP boundary of the reference
(``MonoRUnRoIHead.simple_test``, monorun/models/roi_heads/monorun_roi_head.py:513-529):
``coords_3d [N,3,28,28]``, ``coords_2d [N,2,28,28]``, ``coords_2d_logstd [N,2,28,28]`` (or the full
2x2 whitening ``W=[wxx,wxy,wyy]`` as ``[N,3,28,28]``).  Constants come from the reference's shipped
files: camera matrix demo/calib.csv:1-3, class dimension statistics configs/kitti_multiclass.py:73-80,
``std_scale=10`` uncert_prop_pnp_optimizer.py:28, ``allowed_border=200`` configs/kitti_multiclass.py:131.

numpy only: generation is host-side so that the CPU oracle and the CUDA path see identical bytes.
"""
import numpy as np

KITTI_K = np.array([[707.0912, 0.0, 601.8873],
                    [0.0, 707.0912, 183.1104],
                    [0.0, 0.0, 1.0]], np.float64)
IMG_SHAPE = (375, 1242)  # (h, w)
DIM_MEANS = np.array([(3.89, 1.53, 1.62), (0.82, 1.78, 0.63), (1.77, 1.72, 0.57)])  # (l, h, w)
DIM_STDS = np.array([(0.44, 0.14, 0.11), (0.25, 0.13, 0.12), (0.15, 0.10, 0.14)])
STD_SCALE = 10.0
ALLOWED_BORDER = 200.0
ROI = 28
BASE_SEED = 20261017


def rng_for(config, rank=0):
    return np.random.default_rng(BASE_SEED + 1000 * int(config) + int(rank))


def rot_y(yaw):
    c, s = np.cos(yaw), np.sin(yaw)
    r = np.zeros(yaw.shape + (3, 3))
    r[..., 0, 0] = c
    r[..., 0, 2] = s
    r[..., 1, 1] = 1.0
    r[..., 2, 0] = -s
    r[..., 2, 2] = c
    return r


def project(K, yaw, t, pts):
    """pts (N,P,3) object frame -> pixel (N,P,2), depth (N,P).  x' = R_y(yaw) X + t."""
    cam = np.einsum('nij,npj->npi', rot_y(yaw), pts) + t[:, None, :]
    z = cam[..., 2]
    uv = np.stack([K[0, 0] * cam[..., 0] / z + K[0, 2], K[1, 1] * cam[..., 1] / z + K[1, 2]], -1)
    return uv, z


def _box_corners(dims):
    l, h, w = dims[:, 0], dims[:, 1], dims[:, 2]
    xs = np.stack([l, l, l, l, -l, -l, -l, -l], 1) * 0.5
    ys = np.stack([0 * h, 0 * h, -h, -h, 0 * h, 0 * h, -h, -h], 1)
    zs = np.stack([w, -w, w, -w, w, -w, w, -w], 1) * 0.5
    return np.stack([xs, ys, zs], -1)  # (N,8,3)


def sample_objects(rng, n, classes=(0,), K=KITTI_K, img_shape=IMG_SHAPE):
    """Class, dims (l,h,w), yaw, t per object; redraw while a projected box corner leaves the image by
    more than ALLOWED_BORDER px (or falls behind the camera)."""
    labels = np.empty(n, np.int64)
    dims = np.empty((n, 3))
    yaw = np.empty(n)
    t = np.empty((n, 3))
    todo = np.arange(n)
    while todo.size:
        m = todo.size
        lab = rng.choice(np.asarray(classes), size=m)
        d = np.maximum(rng.normal(DIM_MEANS[lab], DIM_STDS[lab]), 0.3)
        y = rng.uniform(-np.pi, np.pi, m)
        z = rng.uniform(5.0, 60.0, m)
        x = rng.uniform(-0.4, 0.4, m) * z
        yy = rng.normal(1.65, 0.1, m)
        tt = np.stack([x, yy, z], 1)
        uv, zc = project(K, y, tt, _box_corners(d))
        ok = ((zc > 1.0).all(1)
              & (uv[..., 0] > -ALLOWED_BORDER).all(1) & (uv[..., 0] < img_shape[1] + ALLOWED_BORDER).all(1)
              & (uv[..., 1] > -ALLOWED_BORDER).all(1) & (uv[..., 1] < img_shape[0] + ALLOWED_BORDER).all(1))
        idx = todo[ok]
        labels[idx], dims[idx], yaw[idx], t[idx] = lab[ok], d[ok], y[ok], tt[ok]
        todo = todo[~ok]
    return labels, dims, yaw, t


def _points_in_box(rng, dims, p):
    n = dims.shape[0]
    u = rng.uniform(0.0, 1.0, (n, p, 3))
    l, h, w = dims[:, None, 0], dims[:, None, 1], dims[:, None, 2]
    return np.stack([(u[..., 0] - 0.5) * l, -u[..., 1] * h, (u[..., 2] - 0.5) * w], -1)


def _ray_box(K, yaw, t, dims, uv):
    """First intersection of the pixel rays uv (N,P,2) with each object's box, in the object frame.
    Returns (points (N,P,3), hit (N,P) bool)."""
    r = rot_y(yaw)
    o = -np.einsum('nji,nj->ni', r, t)                      # camera centre in object frame: R^T(-t)
    dcam = np.stack([(uv[..., 0] - K[0, 2]) / K[0, 0], (uv[..., 1] - K[1, 2]) / K[1, 1],
                     np.ones(uv.shape[:-1])], -1)
    d = np.einsum('nji,npj->npi', r, dcam)                  # R^T d
    lo = np.stack([-dims[:, 0] / 2, -dims[:, 1], -dims[:, 2] / 2], -1)[:, None, :]
    hi = np.stack([dims[:, 0] / 2, 0 * dims[:, 1], dims[:, 2] / 2], -1)[:, None, :]
    with np.errstate(divide='ignore', invalid='ignore'):
        inv = 1.0 / d
        t0 = (lo - o[:, None, :]) * inv
        t1 = (hi - o[:, None, :]) * inv
    tn = np.minimum(t0, t1).max(-1)
    tf = np.maximum(t0, t1).min(-1)
    hit = (tn <= tf) & (tn > 0)
    pts = o[:, None, :] + np.where(hit, tn, 0.0)[..., None] * d
    return pts, hit


def make_batch(n, config=2, rank=0, mode='S0', weights='diag', classes=None, rng=None, roi=ROI,
               K=KITTI_K, img_shape=IMG_SHAPE):
    """Synthetic head->PnP boundary tensors.

    mode    'S0' points uniform in the box volume + pixel noise drawn from the stated covariance;
            'S1' grid-faithful: 28x28 bin centres of the projected 2-D box, object coordinates from the
                 ray/box-face intersection + N(0,(0.05 m)^2); rays missing the box become low-weight outliers.
    weights 'identity' (istd == 1), 'diag' (per-axis sigma ~ logU(e^-1, e^1) px) or
            'full' (Sigma = Rot(th) diag(s1^2,s2^2) Rot(th)^T, s ~ logU(0.5, 8) px, W = Sigma^-1/2).

    Returns a dict of numpy arrays: coords_3d [n,3,r,r] f32, coords_2d [n,2,r,r] f32,
    logstd [n,2,r,r] f32 (head-level log-std, istd = exp(-logstd)/STD_SCALE; absent for 'full'),
    w_full [n,3,r,r] f32 ('full' only), cam_mat (3,3) f32, img_shape (2,) f32, labels, dims,
    gt_pose (n,4) f64 [yaw,tx,ty,tz], init_pose (n,4) f32.
    """
    if rng is None:
        rng = rng_for(config, rank)
    if classes is None:
        classes = (0,) if config in (1, 2) else (0, 1, 2)
    p = roi * roi
    labels, dims, yaw, t = sample_objects(rng, n, classes, K, img_shape)
    out = dict(labels=labels, dims=dims.astype(np.float32),
               gt_pose=np.concatenate([yaw[:, None], t], 1),
               cam_mat=K.astype(np.float32), img_shape=np.asarray(img_shape, np.float32))

    # ---- per-point pixel covariance ----
    if weights == 'identity':
        sig = np.ones((n, p, 2))
    elif weights == 'diag':
        sig = np.exp(rng.uniform(-1.0, 1.0, (n, p, 2)))
    elif weights == 'full':
        th = rng.uniform(0.0, np.pi, (n, p))
        s12 = np.exp(rng.uniform(np.log(0.5), np.log(8.0), (n, p, 2)))
    else:
        raise ValueError(weights)

    if mode == 'S0':
        pts = _points_in_box(rng, dims, p)
        uv, _ = project(K, yaw, t, pts)
        e = rng.standard_normal((n, p, 2))
        if weights == 'full':
            c, s = np.cos(th), np.sin(th)
            a, b = e[..., 0] * s12[..., 0], e[..., 1] * s12[..., 1]
            noise = np.stack([c * a - s * b, s * a + c * b], -1)
        else:
            noise = e * sig
        uv = uv + noise
    elif mode == 'S1':
        cu, _ = project(K, yaw, t, _box_corners(dims))
        x1 = np.clip(cu[..., 0].min(1), 0, img_shape[1] - 1.0)
        x2 = np.clip(cu[..., 0].max(1), x1 + 2.0, img_shape[1] + 1.0)
        y1 = np.clip(cu[..., 1].min(1), 0, img_shape[0] - 1.0)
        y2 = np.clip(cu[..., 1].max(1), y1 + 2.0, img_shape[0] + 1.0)
        out['boxes'] = np.stack([x1, y1, x2, y2], 1).astype(np.float32)
        k = (np.arange(roi) + 0.5) / roi
        # RoIAlign(aligned=True) of the pixel-index grid = bin centres
        uu = x1[:, None] - 0.5 + k[None, :] * (x2 - x1)[:, None]
        vv = y1[:, None] - 0.5 + k[None, :] * (y2 - y1)[:, None]
        uv = np.stack([np.broadcast_to(uu[:, None, :], (n, roi, roi)),
                       np.broadcast_to(vv[:, :, None], (n, roi, roi))], -1).reshape(n, p, 2)
        pts, hit = _ray_box(K, yaw, t, dims, uv)
        pts = pts + rng.normal(0.0, 0.05, pts.shape)
        miss = _points_in_box(rng, dims, p)
        pts = np.where(hit[..., None], pts, miss)
        if weights == 'full':
            s12 = np.where(hit[..., None], s12, s12 * np.exp(3.0))
        else:
            sig = np.where(hit[..., None], sig, sig * np.exp(3.0))
        out['hit'] = hit
    else:
        raise ValueError(mode)

    def chw(a):  # (n,p,c) -> [n,c,r,r], point index p = i*roi + j
        return np.ascontiguousarray(a.transpose(0, 2, 1).reshape(n, a.shape[2], roi, roi), np.float32)

    out['coords_3d'] = chw(pts)
    out['coords_2d'] = chw(uv)
    if weights == 'full':
        c, s = np.cos(th), np.sin(th)
        i1, i2 = 1.0 / s12[..., 0], 1.0 / s12[..., 1]
        w = np.stack([c * c * i1 + s * s * i2, c * s * (i1 - i2), s * s * i1 + c * c * i2], -1)
        out['w_full'] = chw(w)
    else:
        out['logstd'] = chw(np.log(sig) - np.log(STD_SCALE))

    dyaw = rng.normal(0.0, 0.05, n)
    dt = rng.normal(0.0, 1.0, (n, 3)) * (0.02 * t[:, 2:3])
    out['init_pose'] = np.concatenate([(yaw + dyaw)[:, None], t + dt], 1).astype(np.float32)
    return out


def to_op_level(batch, std_scale=STD_SCALE, allowed_border=ALLOWED_BORDER):
    """Head-level tensors -> the op-level arguments of ``PnPUncert.forward``, as
    UncertPropPnPOptimizer.forward does (uncert_prop_pnp_optimizer.py:71-88), in numpy fp32."""
    c2 = batch['coords_2d']
    n, _, h, w = c2.shape
    out = dict(coords_2d=np.ascontiguousarray(c2.transpose(0, 2, 3, 1).reshape(n, h * w, 2)),
               coords_3d=np.ascontiguousarray(batch['coords_3d'].transpose(0, 2, 3, 1).reshape(n, h * w, 3)),
               cam_mats=batch['cam_mat'][None].astype(np.float32))
    if 'logstd' in batch:
        istd = (np.exp(-batch['logstd']) / np.float32(std_scale)).astype(np.float32)
        out['coords_2d_istd'] = np.ascontiguousarray(istd.transpose(0, 2, 3, 1).reshape(n, h * w, 2))
    else:
        out['w_full'] = np.ascontiguousarray(batch['w_full'].transpose(0, 2, 3, 1).reshape(n, h * w, 3))
    ih, iw = batch['img_shape']
    out['u_range'] = np.array([[-allowed_border, iw + allowed_border]], np.float32)
    out['v_range'] = np.array([[-allowed_border, ih + allowed_border]], np.float32)
    return out


NOC_MEANS = (-0.1, -0.5, 0.0)   # configs/kitti_multiclass.py:106-108 (NOCCoder target_means / target_stds)
NOC_STDS = (0.35, 0.23, 0.34)


def to_head_raw(batch, rng=None, dims_var_scale=0.3):
    """An 'S1' batch -> what the fused head->PnP entry consumes: the dense head's RAW maps and per-object rows.

    noc_pred [n,3,r,r] = (coords_3d / dims - mean) / std   (inverse of NOCCoder.decode, noc_coder.py:50-73),
    proj_logstd [n,2,r,r] raw log-std, rois [n,5] = (0, x1, y1, x2, y2), dims [n,3], dims_var [n,3] =
    (dims_var_scale * class sigma)^2 -- the epistemic variance MC dropout would produce -- all float32."""
    if 'boxes' not in batch:
        raise ValueError("to_head_raw needs a grid-faithful batch (mode='S1')")
    n = batch['dims'].shape[0]
    dims = batch['dims'].astype(np.float64)
    mean = np.asarray(NOC_MEANS)[None, :, None, None]
    std = np.asarray(NOC_STDS)[None, :, None, None]
    noc = (batch['coords_3d'].astype(np.float64) / dims[:, :, None, None] - mean) / std
    sd = DIM_STDS[batch['labels']] * dims_var_scale
    if rng is not None:
        sd = sd * np.exp(rng.uniform(-0.5, 0.5, sd.shape))
    rois = np.concatenate([np.zeros((n, 1)), batch['boxes']], 1)
    return dict(noc_pred=noc.astype(np.float32), proj_logstd=batch['logstd'].astype(np.float32),
                rois=rois.astype(np.float32), dims=batch['dims'].astype(np.float32),
                dims_var=(sd * sd).astype(np.float32))
