"""Generates the committed golden fixtures (run in the authoring container, where /root/reference exists):

    python tests/golden/make_golden.py

1. ``hessian_ref.npz`` -- outputs of the REFERENCE's own pure-torch code (monorun/ops/least_squares/jacobian.py
   + hessian.py: forward_proj, get_pose_jacobians, approx_hessian), imported from /root/reference and evaluated in
   fp64 and fp32 on seeded inputs that include z-clipped, uv-clipped and outlier points.  This pins the
   covariance semantics of the oracle and of the CUDA kernel to the reference implementation itself.
2. ``lm_cfg{1,2,3}.npz`` -- seeded inputs of BASELINE.json configs 1-3 (N=16) together with the CPU oracle's LM
   results (pose, cost, evaluation counts, covariance).  The oracle restates Ceres 1.14 (parity unpinned for the
   control flow, see oracle/pnp_oracle.cpp); these vectors freeze its behaviour so that oracle or generator drift
   is caught, and give the GPU tests fixed targets that do not depend on /root/reference.
3. ``exact_hessian_ref.npz`` -- outputs of the REFERENCE's ``exact_hessian`` (hessian.py:5-64, double autograd
   through jacobian.py) on the inputs of ``hessian_ref.npz`` at two poses per object.  On torch >= 2 the reference
   code raises (jacobian.py:28-29 writes in place into the views ``Tensor.split`` returns, which autograd now
   forbids), so the generator runs it with ``Tensor.split`` replaced by a version that returns clones of the same
   pieces -- same values, same autograd graph semantics, no aliasing.  Nothing else is touched.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference/monorun/ops/least_squares'


def load_reference_hessian():
    """Import jacobian.py / hessian.py without triggering monorun/__init__ (which needs mmdet)."""
    pkg = types.ModuleType('ref_ls')
    pkg.__path__ = [REF]
    sys.modules['ref_ls'] = pkg
    mods = {}
    for name in ('jacobian', 'hessian'):
        spec = importlib.util.spec_from_file_location(f'ref_ls.{name}', os.path.join(REF, f'{name}.py'))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f'ref_ls.{name}'] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods['jacobian'], mods['hessian']


def make_hessian_fixture():
    import torch
    from monorun_b200 import synth
    jac, hes = load_reference_hessian()
    b = synth.make_batch(8, config=2, rng=np.random.default_rng(7))
    op = synth.to_op_level(b)
    rng = np.random.default_rng(11)
    pose = b['gt_pose'].copy()
    pose[:, 0] += rng.normal(0, 0.02, 8)
    pose[:, 1:] += rng.normal(0, 0.05, (8, 3))
    c3 = op['coords_3d'].copy()
    # objects 5-7: force clips.  5: some points behind z_min; 6: narrow u range; 7: narrow v range + tiny z
    pose[5, 3] = 1.2
    pose[6, 1] += 0.3 * pose[6, 3]
    u_range = np.tile(op['u_range'], (8, 1)).astype(np.float32)
    v_range = np.tile(op['v_range'], (8, 1)).astype(np.float32)
    u_range[6] = [500.0, 700.0]
    v_range[7] = [150.0, 200.0]
    mask = rng.uniform(size=(8, 784)) > 0.3
    mask[0] = True
    out = dict(coords_2d=op['coords_2d'], coords_2d_istd=op['coords_2d_istd'], coords_3d=c3,
               cam_mats=op['cam_mats'], u_range=u_range, v_range=v_range, z_min=np.float32(0.5),
               pose=pose.astype(np.float32), inlier_mask=mask)
    for tag, dt in (('64', torch.float64), ('32', torch.float32)):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
        yaw, tv = t(out['pose'][:, :1]), t(out['pose'][:, 1:])
        h = hes.approx_hessian(t(out['coords_2d']), t(out['coords_2d_istd']), t(out['coords_3d']),
                               t(out['cam_mats']).expand(8, 3, 3).contiguous(), t(u_range), t(v_range), 0.5, yaw, tv,
                               torch.from_numpy(mask))
        uv, z, zc, uvc, *_ = jac.forward_proj(t(out['coords_2d']), t(out['coords_3d']),
                                              t(out['cam_mats']).expand(8, 3, 3).contiguous(), 0.5, t(u_range),
                                              t(v_range), yaw, tv)
        out[f'H_ref{tag}'] = h.numpy()
        if tag == '64':
            out['uv_ref64'] = uv.numpy()
            out['z_clip_ref'] = zc.numpy()
            out['uv_clip_ref'] = uvc.numpy()
    print('hessian fixture: z-clipped points', int(out['z_clip_ref'].sum()), 'uv-clipped', int(out['uv_clip_ref'].sum()))
    np.savez_compressed(os.path.join(HERE, 'hessian_ref.npz'), **out)


def make_exact_hessian_fixture():
    import torch
    jac, hes = load_reference_hessian()
    g = np.load(os.path.join(HERE, 'hessian_ref.npz'))
    rng = np.random.default_rng(23)
    pose2 = g['pose'].astype(np.float64)
    pose2[:, 0] += rng.normal(0, 0.01, 8)
    pose2[:, 1:] += rng.normal(0, 0.02, (8, 3))
    pose2 = pose2.astype(np.float32)
    orig_split = torch.Tensor.split
    torch.Tensor.split = lambda self, *a, **k: tuple(x.clone() for x in orig_split(self, *a, **k))
    out = dict(pose2=pose2)
    try:
        for tag, dt in (('64', torch.float64), ('32', torch.float32)):
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
            for name, pose in (('', g['pose']), ('_pose2', pose2)):
                for mtag, mask in (('', torch.from_numpy(g['inlier_mask'])), ('_nomask', None)):
                    h = hes.exact_hessian(t(g['coords_2d']), t(g['coords_2d_istd']), t(g['coords_3d']),
                                          t(g['cam_mats']).expand(8, 3, 3).contiguous(), t(g['u_range']),
                                          t(g['v_range']), 0.5, t(pose[:, :1]), t(pose[:, 1:]), mask)
                    out[f'H_exact{tag}{name}{mtag}'] = h.detach().numpy()
    finally:
        torch.Tensor.split = orig_split
        torch.set_grad_enabled(True)
    np.savez_compressed(os.path.join(HERE, 'exact_hessian_ref.npz'), **out)
    print('exact hessian fixture:', sorted(out))


def make_lm_fixtures():
    from monorun_b200 import synth
    from oracle import pnp_driver as od
    n = 16
    for cfg, weights, mode in ((1, 'identity', 'S0'), (2, 'diag', 'S1'), (3, 'full', 'S0')):
        b = synth.make_batch(n, config=cfg, weights=weights, mode=mode)
        op = synth.to_op_level(b)
        full = weights == 'full'
        w = op['w_full'] if full else op['coords_2d_istd']
        mask = od.istd_inlier_masks(w[..., [0, 2]] if full else w, 0.6)
        mask[mask.sum(1) <= 4] = True
        clips = np.array([[0.5, op['u_range'][0, 0], op['u_range'][0, 1], op['v_range'][0, 0], op['v_range'][0, 1]]])
        r = od.lm_batch(op['coords_2d'], op['coords_3d'], w, op['cam_mats'], b['init_pose'], clips, mask, full_w=full,
                        with_pose_cov=True)
        out = dict(coords_3d=b['coords_3d'], coords_2d=b['coords_2d'], cam_mat=b['cam_mat'], img_shape=b['img_shape'],
                   init_pose=b['init_pose'], gt_pose=b['gt_pose'], inlier_mask=mask, oracle_pose=r['pose'],
                   oracle_cost=r['cost'], oracle_stats=r['stats'], oracle_val=r['val'], oracle_tr=r['tr'],
                   oracle_cov_ceres=r['cov'])
        if full:
            out['w_full'] = b['w_full']
        else:
            out['logstd'] = b['logstd']
            h = od.approx_hessian(op['coords_2d'], op['coords_2d_istd'], op['coords_3d'], op['cam_mats'],
                                  op['u_range'], op['v_range'], 0.5, r['pose'][:, :1], r['pose'][:, 1:], mask)
            out['oracle_cov_pipeline'] = np.linalg.inv(h)
        np.savez_compressed(os.path.join(HERE, f'lm_cfg{cfg}.npz'), **out)
        print(f'cfg{cfg}: evals', np.bincount(r['stats'][:, 1]), 'valid', r['val'].all())


if __name__ == '__main__':
    if '--exact-hessian-only' not in sys.argv:
        make_hessian_fixture()
        make_lm_fixtures()
    make_exact_hessian_fixture()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')
