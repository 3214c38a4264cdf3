"""Shared input builder for the 6-DoF solver tests (CPU and GPU)."""
import numpy as np

from monorun_b200 import synth


def rodrigues(w):
    """[N,3] angle-axis -> [N,3,3]."""
    th = np.linalg.norm(w, axis=1)[:, None, None]
    k = w / np.maximum(np.linalg.norm(w, axis=1, keepdims=True), 1e-300)
    K = np.zeros((w.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -k[:, 2], k[:, 1], k[:, 2], -k[:, 0], -k[:, 1], k[:, 0]
    return np.eye(3)[None] + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def make_case(n, full=False, seed=2, tilt=0.15, noise=True, far=False, cfg=2, mode='S0'):
    """fp32 op-level tensors for a 6-DoF problem: the generator's object points and weights, a pose whose angle-axis
    vector is (rx, yaw, rz) with rx, rz ~ N(0, tilt), observations re-projected under that pose plus noise drawn from
    the stated per-pixel covariance, and a perturbed start."""
    b = synth.make_batch(n, config=cfg, weights='full' if full else 'diag', mode=mode)
    op = synth.to_op_level(b)
    rng = np.random.default_rng(seed)
    gt4 = np.asarray(b['gt_pose'], np.float64)
    rvec = np.stack([rng.normal(0, tilt, n), gt4[:, 0], rng.normal(0, tilt, n)], 1)
    t = gt4[:, 1:].copy()
    K = np.asarray(op['cam_mats'], np.float64).reshape(-1, 3, 3)[0]
    c3 = op['coords_3d'].astype(np.float64)
    cam = np.einsum('nij,npj->npi', rodrigues(rvec), c3) + t[:, None, :]
    uv = np.stack([K[0, 0] * cam[..., 0] / cam[..., 2] + K[0, 2], K[1, 1] * cam[..., 1] / cam[..., 2] + K[1, 2]], -1)
    w = (op['w_full'] if full else op['coords_2d_istd']).astype(np.float64)
    if noise:
        e = rng.normal(size=uv.shape)
        if full:   # W = Sigma^-1/2 (symmetric): noise = W^-1 e
            det = w[..., 0] * w[..., 2] - w[..., 1] ** 2
            uv = uv + np.stack([(w[..., 2] * e[..., 0] - w[..., 1] * e[..., 1]) / det,
                                (-w[..., 1] * e[..., 0] + w[..., 0] * e[..., 1]) / det], -1)
        else:
            uv = uv + e / w
    init = np.concatenate([rvec, t], 1)
    s = (0.3, 0.15) if far else (0.03, 0.02)
    init[:, :3] += rng.normal(0, s[0], (n, 3))
    init[:, 3:] += rng.normal(0, 1.0, (n, 3)) * (s[1] * t[:, 2:3])
    rg = np.array([[op['u_range'][0, 0], op['u_range'][0, 1], op['v_range'][0, 0], op['v_range'][0, 1]]], np.float32)
    return dict(c3=op['coords_3d'].astype(np.float32), c2=uv.astype(np.float32), w=w.astype(np.float32),
                cam=np.ascontiguousarray(K[None].astype(np.float32)), uv_range=rg,
                clips=np.concatenate([[[0.5]], rg.astype(np.float64)], 1), init=init.astype(np.float32),
                gt=np.concatenate([rvec, t], 1), yaw_only_pose=gt4, coords_2d_yaw=op['coords_2d'].astype(np.float32),
                init4=np.asarray(b['init_pose'], np.float32))


def oracle_solve(sd, c, full, mask=None, threads=0, init=None):
    return sd.solve_batch(c['c2'], c['c3'], c['w'], c['cam'], c['init'] if init is None else init, c['clips'],
                          inlier_mask=mask, full_w=full, threads=threads)
