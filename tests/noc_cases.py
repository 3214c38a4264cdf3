"""Shared input builder for the 7-parameter solver tests (CPU and GPU)."""
import numpy as np

from monorun_b200 import synth


def make_case(n, full=False, mode='S0', cfg=2, seed=1, prior_sd=0.1, prior_wgt=10.0, far=False):
    """fp32 op-level tensors for pnp_noc_uncert / pnp_noc_cov_uncert: normalised object coordinates, a noisy
    log-dimension prior and an initial [log dims, yaw, t].  far=True perturbs the start strongly (rejected steps,
    depth / image-border clips)."""
    b = synth.make_batch(n, config=cfg, weights='full' if full else 'diag', mode=mode)
    op = synth.to_op_level(b)
    dims = b['dims'].astype(np.float64)
    rng = np.random.default_rng(seed)
    logdim = (np.log(dims) + rng.normal(0, prior_sd, (n, 3))).astype(np.float32)
    init_pose = np.array(b['init_pose'], np.float64)
    if far:
        init_pose[:, 0] += rng.normal(0, 0.6, n)
        init_pose[:, 1:] += rng.normal(0, 1.0, (n, 3)) * np.array([2.0, 0.5, 6.0])
    rg = np.array([[op['u_range'][0, 0], op['u_range'][0, 1], op['v_range'][0, 0], op['v_range'][0, 1]]], np.float32)
    return dict(
        noc=(op['coords_3d'] / dims[:, None, :]).astype(np.float32),
        c2=op['coords_2d'].astype(np.float32),
        w=(op['w_full'] if full else op['coords_2d_istd']).astype(np.float32),
        logdim=logdim, logdim_wgt=np.full((n, 3), prior_wgt, np.float32),
        init=np.concatenate([logdim, init_pose], 1).astype(np.float32),
        cam=np.ascontiguousarray(np.asarray(op['cam_mats'], np.float32).reshape(-1, 3, 3)[:1]),
        uv_range=rg, clips=np.concatenate([[[0.5]], rg.astype(np.float64)], 1),
        dims=dims, gt_pose=np.asarray(b['gt_pose'], np.float64), coords_3d=op['coords_3d'],
        init_pose=np.asarray(b['init_pose'], np.float64))


def oracle_solve(nd, c, delta, full, mask=None, threads=0):
    return nd.noc_batch(c['c2'], c['noc'], c['w'], c['logdim'], c['logdim_wgt'], c['cam'], c['init'], c['clips'],
                        delta, inlier_mask=mask, full_w=full, threads=threads)
