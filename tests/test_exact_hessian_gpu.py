"""GPU parity tests (-m gpu) of mrpnp_exact_hessian and PnPUncert(forward_exact_hessian=True) against the REFERENCE's
own autograd outputs (tests/golden/exact_hessian_ref.npz) and the oracle restatement of hessian.py:5-64."""
import os

import numpy as np
import pytest
import torch

from monorun_b200 import synth
from tests.test_exact_hessian import VARIANTS, load

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rel(a, b):
    return np.linalg.norm(a - b, axis=(1, 2)) / np.linalg.norm(b, axis=(1, 2))


@pytest.mark.parametrize('layout', ['interleaved', 'planar'])
@pytest.mark.parametrize('tag,pose_key,masked', VARIANTS)
def test_exact_hessian_matches_reference_autograd(cuda_lib, tag, pose_key, masked, layout):
    from monorun_b200 import pnp
    g = load()
    c3, c2, w = (dev(g[k].astype(np.float32)) for k in ('coords_3d', 'coords_2d', 'coords_2d_istd'))
    if layout == 'planar':
        c3, c2, w = (t.permute(0, 2, 1).contiguous() for t in (c3, c2, w))
    uvr = dev(np.concatenate([g['u_range'], g['v_range']], 1).astype(np.float32))
    rows = torch.zeros((8, 24), device='cuda')
    rows[:, :4] = dev(g[pose_key].astype(np.float32))
    rows[:, 20] = 1.0
    h = pnp.exact_hessian(c3, c2, w, dev(g['cam_mats'].astype(np.float32)), uvr, rows,
                          dev(g['inlier_mask']) if masked else None, layout=layout, weight_mode='istd', rows=rows)
    torch.cuda.synchronize()
    ref = g['H_exact64' + tag]
    assert _rel(h.cpu().numpy().astype(np.float64), ref).max() < 1e-6   # fp64 accumulation, fp32 output
    cov = rows[:, 4:20].reshape(8, 4, 4).cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(cov, np.linalg.inv(ref), rtol=1e-5, atol=1e-6 * np.abs(np.linalg.inv(ref)).max())
    assert (rows[:, 20] == 1).all()


def test_logstd_weights_and_singular_objects(cuda_lib, oracle):
    from monorun_b200 import pnp
    g = load()
    istd = g['coords_2d_istd'].astype(np.float64)
    logstd = (-np.log(istd * 10.0)).astype(np.float32)
    mask = g['inlier_mask'].copy()
    mask[3] = False  # no contributing row: singular
    rows = torch.zeros((8, 24), device='cuda')
    rows[:, :4] = dev(g['pose'].astype(np.float32))
    rows[:, 20] = 1.0
    h = pnp.exact_hessian(dev(g['coords_3d'].astype(np.float32)), dev(g['coords_2d'].astype(np.float32)), dev(logstd),
                          dev(g['cam_mats'].astype(np.float32)),
                          dev(np.concatenate([g['u_range'], g['v_range']], 1).astype(np.float32)), rows, dev(mask),
                          layout='interleaved', weight_mode='logstd', std_scale=10.0, rows=rows)
    torch.cuda.synchronize()
    ref = oracle.exact_hessian(g['coords_2d'], np.exp(-logstd.astype(np.float64)) / 10.0, g['coords_3d'], g['cam_mats'],
                               g['u_range'], g['v_range'], 0.5, g['pose'][:, :1], g['pose'][:, 1:], mask)
    keep = np.arange(8) != 3
    assert _rel(h.cpu().numpy()[keep].astype(np.float64), ref[keep]).max() < 1e-6
    assert rows[3, 20] == 0 and torch.equal(rows[3, 4:20].reshape(4, 4), torch.eye(4, device='cuda'))
    assert (rows[torch.from_numpy(keep).cuda(), 20] == 1).all()


def test_pnp_uncert_with_forward_exact_hessian(cuda_lib, oracle):
    """Drop-in signature: the pose is the one of the default call, the covariance is the inverse of the reference's
    second-order Hessian at that pose over the inlier mask actually used."""
    import monorun_b200
    b = synth.make_batch(64, config=2, mode='S1')
    op = synth.to_op_level(b)
    args = [dev(op[k].astype(np.float32)) for k in ('coords_2d', 'coords_2d_istd', 'coords_3d', 'cam_mats', 'u_range',
                                                    'v_range')]
    mod = monorun_b200.build_pnp(dict(type='PnPUncert', z_min=0.5, epnp_istd_thres=0.6, inlier_opt_only=True,
                                      forward_exact_hessian=True)).cuda()
    ref_mod = monorun_b200.build_pnp(dict(type='PnPUncert', z_min=0.5, epnp_istd_thres=0.6, inlier_opt_only=True,
                                          forward_exact_hessian=False)).cuda()
    init = dev(b['init_pose'].astype(np.float32))
    val, r_vec, t_vec, cov, inl = mod(*args, init_pose=init)
    val0, r0, t0, cov0, inl0 = ref_mod(*args, init_pose=init)
    torch.cuda.synchronize()
    assert torch.equal(r_vec, r0) and torch.equal(t_vec, t0) and torch.equal(inl, inl0) and val.all()
    h = oracle.exact_hessian(op['coords_2d'], op['coords_2d_istd'].astype(np.float32), op['coords_3d'], op['cam_mats'],
                             op['u_range'], op['v_range'], 0.5, r_vec.cpu().numpy(), t_vec.cpu().numpy(),
                             inl.cpu().numpy())
    ref_cov = np.linalg.inv(h)
    got = cov.cpu().numpy().astype(np.float64)
    assert _rel(got, ref_cov).max() < 1e-4
    assert _rel(cov0.cpu().numpy().astype(np.float64), ref_cov).max() > 1e-4   # it is not the Gauss-Newton covariance

    # head-level entry (planar tensors, log-std weights)
    hl = [dev(b[k]) for k in ('coords_2d', 'logstd', 'coords_3d')]
    uvr = torch.cat([args[4], args[5]], 1)
    val2, r2, t2, cov2, _ = mod.forward_dense(hl[0], hl[1], hl[2], args[3], uvr, 10.0, init_pose=init)
    torch.cuda.synchronize()
    assert torch.allclose(t2, t_vec, rtol=1e-4, atol=1e-5)
    assert _rel(cov2.cpu().numpy().astype(np.float64), ref_cov).max() < 5e-3
