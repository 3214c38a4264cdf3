"""GPU parity tests (-m gpu) of mrpnp_solve_6dof against its CPU oracle (oracle/pnp_6dof_oracle.cpp) on identical
seeded inputs, through the C ABI.  fp64 kernel: held to 1e-7 on objects whose trust-region decisions match (>= 95 %).
Mixed kernel (the default; fp64 cost chain, fp32 normal equations): the same trust-region paths, the north star's
tolerances on the pose (1e-4 relative on the translation, 1e-3 rad on the rotation) with two decades to spare, 1e-3
relative (Frobenius) on the covariance."""
import numpy as np
import pytest
import torch

from tests.sixdof_cases import make_case, oracle_solve

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def sd():
    from oracle import sixdof_driver
    sixdof_driver.build()
    return sixdof_driver


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_solve(c, full, mask=None, layout='interleaved', init=None, logstd=False, precision='fp64'):
    from monorun_b200 import pnp
    c3, c2, w = dev(c['c3']), dev(c['c2']), dev(c['w'])
    mode = 'full' if full else 'istd'
    if logstd:
        w, mode = -torch.log(w * 10.0), 'logstd'
    if layout == 'planar':
        c3, c2, w = (t.permute(0, 2, 1).contiguous() for t in (c3, c2, w))
    res = pnp.solve_6dof_batched(c3, c2, w, dev(c['cam']), dev(c['uv_range']), dev(c['init'] if init is None else init),
                                 dev(mask) if mask is not None else None, layout=layout, weight_mode=mode, precision=precision)
    torch.cuda.synchronize()
    return res.cpu().numpy()


def check(g, r, min_same=0.95, tol=1e-7):
    # identical trust-region paths: same evaluation count and the same final cost to rounding.  (FMA contraction differs
    # between nvcc and g++, so an object sitting on an accept / reject or tolerance boundary can take another path and
    # stop up to Ceres' function tolerance away; those are held to the cost check below.)
    same = (g[:, 45] == r['stats'][:, 1]) & (np.abs(g[:, 44] - r['cost']) <= 1e-9 * r['cost'])
    assert same.mean() >= min_same, same.mean()
    np.testing.assert_allclose(g[same, :6], r['pose'][same], rtol=tol, atol=tol * 0.1)
    np.testing.assert_array_equal(g[:, 42] > 0, r['val'])
    np.testing.assert_allclose(g[:, 44], r['cost'], rtol=1e-4)
    cov = g[same, 6:42].reshape(-1, 6, 6)
    rel = np.linalg.norm(cov - r['cov'][same], axis=(1, 2)) / np.linalg.norm(r['cov'][same], axis=(1, 2))
    assert rel.max() < 1e-5, rel.max()


def check_mixed(g, r, min_same=0.99):
    """The mixed kernel against the oracle: identical evaluation counts (its cost chain is the fp64 kernel's) on >= 99 %
    of the objects, and on those the pose twenty times inside the north star's tolerances (1e-4 relative on t, 1e-3 rad)
    and the covariance to 1e-3 relative.  An object on an accept / reject or function-tolerance boundary can take another
    path (as with the fp64 kernel, see check()): those stop within Ceres' function tolerance of the same cost and are
    held to 1e-3 on the pose."""
    from tests.sixdof_cases import rodrigues
    np.testing.assert_array_equal(g[:, 42] > 0, r['val'])
    ok = r['val']
    same = (g[:, 45] == r['stats'][:, 1]) & ok
    assert same.sum() >= min_same * ok.sum(), same.mean()
    t_rel = np.linalg.norm(g[:, 3:6] - r['pose'][:, 3:], axis=1) / np.linalg.norm(r['pose'][:, 3:], axis=1)
    cosang = np.clip((np.einsum('nij,nij->n', rodrigues(g[:, :3]), rodrigues(r['pose'][:, :3])) - 1) / 2, -1, 1)
    rot = np.arccos(cosang)
    assert t_rel[same].max() < 5e-6 and rot[same].max() < 5e-6, (t_rel[same].max(), rot[same].max())
    assert t_rel[ok].max() < 1e-3 and rot[ok].max() < 1e-3, (t_rel[ok].max(), rot[ok].max())
    np.testing.assert_allclose(g[ok, 44], r['cost'][ok], rtol=2e-6)
    cov = g[same, 6:42].reshape(-1, 6, 6)
    rel = np.linalg.norm(cov - r['cov'][same], axis=(1, 2)) / np.linalg.norm(r['cov'][same], axis=(1, 2))
    assert rel.max() < 1e-3, rel.max()
    assert (np.linalg.eigvalsh(0.5 * (cov + cov.transpose(0, 2, 1))) > 0).all()


@pytest.mark.parametrize('full', [False, True])
@pytest.mark.parametrize('layout', ['interleaved', 'planar'])
def test_6dof_parity_with_oracle(cuda_lib, sd, full, layout):
    for far in (False, True):
        c = make_case(256, full=full, far=far)
        check(gpu_solve(c, full, layout=layout), oracle_solve(sd, c, full))


@pytest.mark.parametrize('full', [False, True])
@pytest.mark.parametrize('layout', ['interleaved', 'planar'])
def test_6dof_mixed_parity_with_oracle(cuda_lib, sd, full, layout):
    for far in (False, True):
        c = make_case(256, full=full, far=far)
        check_mixed(gpu_solve(c, full, layout=layout, precision='mixed'), oracle_solve(sd, c, full))


def test_6dof_mixed_masks_layouts_and_fp64_kernel(cuda_lib, sd):
    """Ragged inlier masks (an object with 13 inliers, one with every point), both layouts bit-identical, log-std weights,
    and the fp64 kernel as second reference; repeated launches reuse the self-resetting work counters."""
    n = 200
    c = make_case(n, tilt=0.05, seed=7)
    rng = np.random.default_rng(11)
    mask = rng.uniform(size=c['c3'].shape[:2]) < rng.uniform(0.2, 1.0, (n, 1))
    mask[0] = False
    mask[0, 5::64] = True     # 13 inliers
    mask[1] = True
    r = oracle_solve(sd, c, False, mask=mask)
    gi = gpu_solve(c, False, mask=mask, precision='mixed')
    gp = gpu_solve(c, False, mask=mask, layout='planar', precision='mixed')
    np.testing.assert_array_equal(gi, gp)
    for _ in range(3):
        np.testing.assert_array_equal(gpu_solve(c, False, mask=mask, precision='mixed'), gi)
    check_mixed(gi, r, min_same=0.98)
    g64 = gpu_solve(c, False, mask=mask)
    both = (g64[:, 42] > 0) & (gi[:, 42] > 0) & (g64[:, 45] == gi[:, 45])
    assert both.mean() > 0.97
    np.testing.assert_allclose(gi[both, :6], g64[both, :6], rtol=2e-6, atol=2e-7)
    gl = gpu_solve(c, False, mask=mask, logstd=True, layout='planar', precision='mixed')
    close = np.abs(gl[:, :6] - gi[:, :6]).max(1) < 1e-3
    assert close.mean() >= 0.95 and ((gl[:, 42] > 0) == (gi[:, 42] > 0)).mean() > 0.98


def test_6dof_mixed_at_full_size(cuda_lib):
    """8192 x 784 (config 3's size), no oracle: against the fp64 kernel, objects with the same evaluation count inside a
    twentieth of the tolerances, every object's cost within Ceres' function tolerance of the fp64 kernel's."""
    from tests.sixdof_cases import rodrigues
    c = make_case(8192, full=True, cfg=3, mode='S1')
    gm = gpu_solve(c, True, layout='planar', precision='mixed')
    g64 = gpu_solve(c, True, layout='planar')
    assert ((gm[:, 42] > 0) == (g64[:, 42] > 0)).all()
    same = gm[:, 45] == g64[:, 45]
    assert same.mean() > 0.999, same.mean()
    t_rel = np.linalg.norm(gm[:, 3:6] - g64[:, 3:6], axis=1) / np.linalg.norm(g64[:, 3:6], axis=1)
    cosang = np.clip((np.einsum('nij,nij->n', rodrigues(gm[:, :3]), rodrigues(g64[:, :3])) - 1) / 2, -1, 1)
    assert t_rel[same].max() < 5e-6 and np.arccos(cosang)[same].max() < 5e-6
    assert t_rel.max() < 1e-3 and np.arccos(cosang).max() < 1e-3
    np.testing.assert_allclose(gm[:, 44], g64[:, 44], rtol=2e-6)
    cm, c64 = gm[same, 6:42].reshape(-1, 6, 6), g64[same, 6:42].reshape(-1, 6, 6)
    rel = np.linalg.norm(cm - c64, axis=(1, 2)) / np.linalg.norm(c64, axis=(1, 2))
    assert rel.max() < 2e-3 and np.median(rel) < 2e-4, (rel.max(), np.median(rel))


def test_6dof_masks_first_order_branch_and_logstd_weights(cuda_lib, sd):
    c = make_case(128, tilt=0.05)
    init = c['init'].copy()
    init[:, :3] = 0.0
    rng = np.random.default_rng(4)
    mask = rng.uniform(size=c['c3'].shape[:2]) < rng.uniform(0.3, 1.0, (128, 1))
    r = oracle_solve(sd, c, False, mask=mask, init=init)
    check(gpu_solve(c, False, mask=mask, init=init), r, min_same=0.9)
    # log-std weights are exponentiated in the kernel: same problem up to the rounding of log / exp
    g = gpu_solve(c, False, mask=mask, init=init, logstd=True, layout='planar')
    # (the start r_vec = 0 is far from most true yaws, so a few objects sit on chaotic trajectories where the 1e-7
    # weight perturbation selects another local minimum)
    close = np.abs(g[:, :6] - r['pose']).max(1) < 1e-3
    assert close.mean() >= 0.9 and (g[:, 42] > 0).all(), close.mean()


def test_pnp_uncert_forward_6dof(cuda_lib, sd):
    """PnPUncert(use_6dof=True).forward_6dof: the 4-DoF solve seeds (0, yaw, 0, t); all six parameters are refined."""
    import monorun_b200
    c = make_case(64, tilt=0.08)
    m = monorun_b200.build_pnp(dict(type='PnPUncert', z_min=0.5, epnp_istd_thres=0.6, inlier_opt_only=True,
                                    use_6dof=True)).cuda()
    u_range, v_range = dev(c['uv_range'][:, :2]), dev(c['uv_range'][:, 2:])
    ret, r_vec, t_vec, cov, inl = m.forward_6dof(dev(c['c2']), dev(c['w']), dev(c['c3']), dev(c['cam']), u_range, v_range)
    torch.cuda.synchronize()
    assert ret.all() and r_vec.shape == (64, 3) and t_vec.shape == (64, 3) and cov.shape == (64, 6, 6)
    pose = torch.cat([r_vec, t_vec], 1).double().cpu().numpy()
    from tests.sixdof_cases import rodrigues
    rot_err = np.linalg.norm(rodrigues(pose[:, :3]) - rodrigues(c['gt'][:, :3]), axis=(1, 2))
    assert rot_err.max() < 0.05, rot_err.max()                         # the tilt is recovered (angle-axis may wrap)
    t_err = np.linalg.norm(pose[:, 3:] - c['gt'][:, 3:], axis=1) / np.linalg.norm(c['gt'][:, 3:], axis=1)
    assert np.median(t_err) < 5e-3
    # against the oracle started from the same 4-DoF result over the same inlier mask
    init6 = np.zeros((64, 6), np.float32)
    ret4, yaw4, t4, _, _ = m(dev(c['c2']), dev(c['w']), dev(c['c3']), dev(c['cam']), u_range, v_range)
    init6[:, 1] = yaw4[:, 0].cpu().numpy()
    init6[:, 3:] = t4.cpu().numpy()
    r = oracle_solve(sd, c, False, mask=inl.cpu().numpy(), init=init6)
    same = r['val']
    assert np.linalg.norm(rodrigues(pose[same, :3]) - rodrigues(r['pose'][same, :3]), axis=(1, 2)).max() < 1e-4
    assert np.abs(pose[same, 3:] - r['pose'][same, 3:]).max() < 1e-3
    eig = np.linalg.eigvalsh(cov.double().cpu().numpy())
    assert (eig > 0).all()


def test_6dof_argument_errors_and_empty_batch(cuda_lib):
    from monorun_b200 import pnp
    c = make_case(2)
    out = pnp.solve_6dof_batched(dev(c['c3'][:0]), dev(c['c2'][:0]), dev(c['w'][:0]), dev(c['cam']), dev(c['uv_range']),
                                 dev(c['init'][:0]), layout='interleaved', weight_mode='istd')
    assert out.shape == (0, 48)
    with pytest.raises(KeyError):
        pnp.solve_6dof_batched(dev(c['c3']), dev(c['c2']), dev(c['w']), dev(c['cam']), dev(c['uv_range']),
                               dev(c['init']), layout='interleaved', weight_mode='bogus')
