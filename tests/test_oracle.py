"""CPU tests of the oracle (oracle/): pinned against the reference's own torch code, the committed golden
vectors, finite differences, ground-truth recovery and an independent minimiser (scipy)."""
import os

import numpy as np
import pytest

from monorun_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _clips(op):
    return np.array([[0.5, op['u_range'][0, 0], op['u_range'][0, 1], op['v_range'][0, 0], op['v_range'][0, 1]]])


def _mask(od, w, full=False):
    m = od.istd_inlier_masks(w[..., [0, 2]] if full else w, 0.6)
    m[m.sum(1) <= 4] = True
    return m


def test_approx_hessian_matches_reference_torch_code(oracle):
    """hessian_ref.npz holds outputs of the reference's hessian.py/jacobian.py (imported from /root/reference by
    tests/golden/make_golden.py) on inputs with z-clipped, uv-clipped and outlier points."""
    g = np.load(os.path.join(GOLD, 'hessian_ref.npz'))
    assert g['z_clip_ref'].sum() > 0 and g['uv_clip_ref'].sum() > 0
    h = oracle.approx_hessian(g['coords_2d'], g['coords_2d_istd'], g['coords_3d'], g['cam_mats'], g['u_range'],
                              g['v_range'], 0.5, g['pose'][:, :1], g['pose'][:, 1:], g['inlier_mask'])
    rel64 = np.linalg.norm(h - g['H_ref64'], axis=(1, 2)) / np.linalg.norm(g['H_ref64'], axis=(1, 2))
    assert rel64.max() < 1e-12, rel64
    rel32 = np.linalg.norm(h - g['H_ref32'], axis=(1, 2)) / np.linalg.norm(g['H_ref64'], axis=(1, 2))
    assert rel32.max() < 1e-4, rel32  # the reference itself evaluates this in fp32


@pytest.mark.parametrize('cfg', [1, 2, 3])
def test_oracle_reproduces_golden_lm_vectors(oracle, cfg):
    g = np.load(os.path.join(GOLD, f'lm_cfg{cfg}.npz'))
    full = 'w_full' in g.files
    weights = 'full' if full else ('identity' if cfg == 1 else 'diag')
    b = synth.make_batch(16, config=cfg, weights=weights, mode='S1' if cfg == 2 else 'S0')
    for k in ('coords_3d', 'coords_2d', 'init_pose'):  # the generator is frozen too
        np.testing.assert_array_equal(b[k], g[k])
    op = synth.to_op_level(b)
    w = op['w_full'] if full else op['coords_2d_istd']
    r = oracle.lm_batch(op['coords_2d'], op['coords_3d'], w, op['cam_mats'], g['init_pose'], _clips(op),
                        g['inlier_mask'], full_w=full, with_pose_cov=True)
    np.testing.assert_allclose(r['pose'], g['oracle_pose'], rtol=1e-10, atol=1e-12)
    np.testing.assert_array_equal(r['stats'], g['oracle_stats'])
    np.testing.assert_allclose(r['cov'], g['oracle_cov_ceres'], rtol=1e-7)
    assert r['val'].all()


def test_single_call_abi_equals_batch(oracle):
    b = synth.make_batch(4, config=2)
    op = synth.to_op_level(b)
    m = _mask(oracle, op['coords_2d_istd'])
    r = oracle.lm_batch(op['coords_2d'], op['coords_3d'], op['coords_2d_istd'], op['cam_mats'], b['init_pose'],
                        _clips(op), m, with_pose_cov=True)
    for i in range(4):
        val, pose, cov, tr = oracle.lm_single(op['coords_2d'][i][m[i]], op['coords_3d'][i][m[i]],
                                              op['coords_2d_istd'][i][m[i]], op['cam_mats'][0], b['init_pose'][i],
                                              _clips(op)[0])
        assert val
        np.testing.assert_array_equal(pose, r['pose'][i])
        np.testing.assert_array_equal(cov, r['cov'][i])
        assert tr == r['tr'][i]


@pytest.mark.parametrize('full', [False, True])
def test_jacobian_against_finite_differences(oracle, full):
    b = synth.make_batch(2, config=3 if full else 2, weights='full' if full else 'diag')
    op = synth.to_op_level(b)
    w = op['w_full'] if full else op['coords_2d_istd']
    x = b['init_pose'][0].astype(np.float64)
    args = (op['coords_2d'][0], op['coords_3d'][0], w[0], op['cam_mats'][0])
    c0, g, jtj = oracle.eval_cost_grad_hess(*args, x, _clips(op)[0], full_w=full)
    num = np.zeros(4)
    for k in range(4):
        h = 1e-6 * max(1.0, abs(x[k]))
        xp, xm = x.copy(), x.copy()
        xp[k] += h
        xm[k] -= h
        num[k] = (oracle.eval_cost_grad_hess(*args, xp, _clips(op)[0], full_w=full)[0]
                  - oracle.eval_cost_grad_hess(*args, xm, _clips(op)[0], full_w=full)[0]) / (2 * h)
    np.testing.assert_allclose(g, num, rtol=1e-5, atol=1e-6 * np.abs(num).max())
    assert np.all(np.linalg.eigvalsh(jtj) > 0)


def test_ground_truth_recovery_noise_free(oracle):
    """Config 1 plumbing: identity covariance, exact correspondences -> the true pose is recovered."""
    rng = np.random.default_rng(5)
    labels, dims, yaw, t = synth.sample_objects(rng, 8)
    pts = synth._points_in_box(rng, dims, 784)
    uv, _ = synth.project(synth.KITTI_K, yaw, t, pts)
    init = np.concatenate([(yaw + 0.05)[:, None], t * 1.02], 1)
    clips = np.array([[0.5, -200.0, 1442.0, -200.0, 575.0]])
    r = oracle.lm_batch(uv, pts, np.ones_like(uv), synth.KITTI_K[None], init, clips)
    assert r['val'].all()
    assert np.abs(r['pose'][:, 0] - yaw).max() < 1e-6
    assert (np.linalg.norm(r['pose'][:, 1:] - t, axis=1) / np.linalg.norm(t, axis=1)).max() < 1e-6


def test_oracle_stops_within_ceres_slack_of_true_minimiser(oracle):
    """Independent check of the restated LM: scipy's MINPACK solution of the same residuals is the true
    minimiser; Ceres' function-tolerance exit may stop up to ~1e-4 (relative translation) short of it."""
    from scipy.optimize import least_squares
    b = synth.make_batch(24, config=2)
    op = synth.to_op_level(b)
    m = _mask(oracle, op['coords_2d_istd'])
    clips = _clips(op)
    r = oracle.lm_batch(op['coords_2d'], op['coords_3d'], op['coords_2d_istd'], op['cam_mats'], b['init_pose'], clips, m)
    K = b['cam_mat'].astype(np.float64)

    def resid(x, p2, p3, w):
        c, s = np.cos(x[0]), np.sin(x[0])
        xc = c * p3[:, 0] + s * p3[:, 2] + x[1]
        yc = p3[:, 1] + x[2]
        z = np.maximum(-s * p3[:, 0] + c * p3[:, 2] + x[3], 0.5)
        u = np.clip(K[0, 0] * xc / z + K[0, 2], clips[0, 1], clips[0, 2])
        v = np.clip(K[1, 1] * yc / z + K[1, 2], clips[0, 3], clips[0, 4])
        return np.concatenate([(u - p2[:, 0]) * w[:, 0], (v - p2[:, 1]) * w[:, 1]])

    for i in range(24):
        a = tuple(x[i][m[i]].astype(np.float64) for x in (op['coords_2d'], op['coords_3d'], op['coords_2d_istd']))
        s = least_squares(resid, r['pose'][i], args=a, xtol=1e-15, ftol=1e-15, gtol=1e-15, method='lm')
        assert np.linalg.norm(s.x[1:] - r['pose'][i, 1:]) / np.linalg.norm(s.x[1:]) < 2e-4
        assert abs(s.x[0] - r['pose'][i, 0]) < 5e-4
        assert 0.5 * np.sum(resid(r['pose'][i], *a) ** 2) == pytest.approx(r['cost'][i], rel=1e-12)


def test_adopt_candidate_switch_moves_result_towards_minimiser(oracle):
    b = synth.make_batch(16, config=2)
    op = synth.to_op_level(b)
    m = _mask(oracle, op['coords_2d_istd'])
    args = (op['coords_2d'], op['coords_3d'], op['coords_2d_istd'], op['cam_mats'], b['init_pose'], _clips(op), m)
    r0 = oracle.lm_batch(*args)
    oracle.lib().pnp_oracle_set_adopt_candidate_on_ftol(1)
    try:
        r1 = oracle.lm_batch(*args)
    finally:
        oracle.lib().pnp_oracle_set_adopt_candidate_on_ftol(0)
    assert (r1['cost'] <= r0['cost'] + 1e-12).all() and (r1['cost'] < r0['cost']).any()
    np.testing.assert_array_equal(r0['stats'][:, 1], r1['stats'][:, 1])


def test_restated_driver_with_opencv_epnp(oracle):
    """u2d_pnp_cpu / pnp_uncert_ref with the reference's EPnP(+RANSAC) initialisation (pnp_uncert_cpu.py:34-58)."""
    b = synth.make_batch(6, config=2)
    op = synth.to_op_level(b)
    args = (op['coords_2d'], op['coords_2d_istd'], op['coords_3d'], op['cam_mats'], op['u_range'], op['v_range'])
    ret, r_vec, t_vec, cov, inl = oracle.pnp_uncert_ref(*args, z_min=0.5, epnp_istd_thres=0.6, inlier_opt_only=True)
    assert ret.all() and r_vec.shape == (6, 1) and t_vec.shape == (6, 3) and cov.shape == (6, 4, 4)
    assert inl.shape == (6, 784) and inl.dtype == bool
    gt = b['gt_pose']
    assert (np.linalg.norm(t_vec - gt[:, 1:], axis=1) / np.linalg.norm(gt[:, 1:], axis=1)).max() < 0.05
    thr = np.full(6, 8.0, np.float32)
    ret2, _, t2, _, inl2 = oracle.pnp_uncert_ref(*args, z_min=0.5, epnp_istd_thres=0.6, epnp_ransac_thres=thr,
                                                 inlier_opt_only=True)
    assert ret2.all() and (inl2.sum(1) <= inl.sum(1)).all()
    # empty batch: pnp_uncert_cpu.py:201-207
    e = oracle.u2d_pnp_cpu(op['coords_2d'][:0], op['coords_2d_istd'][:0], op['coords_3d'][:0], op['cam_mats'],
                           op['u_range'], op['v_range'])
    assert e[0].shape == (0,) and e[1].shape == (0, 1) and e[3].shape == (0, 4, 4) and e[5].shape == (0, 784)


def test_too_few_inliers_uses_all_points(oracle):
    b = synth.make_batch(1, config=2)
    op = synth.to_op_level(b)
    m = np.zeros(784, bool)
    m[:3] = True
    out = oracle.u2d_pnp_cpu_single(op['coords_2d'][0], op['coords_2d_istd'][0], op['coords_3d'][0], m,
                                    op['cam_mats'][0], op['u_range'][0], op['v_range'][0], None,
                                    inlier_opt_only=True, init_pose=b['init_pose'][0])
    assert out[0] and out[5].all()


# ---------------------------------------------------------------- Ceres' own published known answers
# The iteration tables the Ceres Solver tutorial prints for its two introductory problems
# (docs/source/nnls_tutorial.rst: output of examples/helloworld.cc and examples/powell.cc, DENSE_QR, default
# Solver::Options, minimizer_progress_to_stdout).  Columns: iter, cost, cost_change, |gradient|, |step|, tr_ratio,
# tr_radius.  Ceres is not vendored by the reference and there is no network here, so these rows are transcribed
# from the public documentation; they pin the trust-region control flow of the restated minimiser (Jacobi scaling,
# LM diagonal, step acceptance, radius update, termination tests) to Ceres' own output, to every printed digit.
HELLO_WORLD = """
   0  4.512500e+01    0.00e+00    9.50e+00   0.00e+00   0.00e+00  1.00e+04
   1  4.511598e-07    4.51e+01    9.50e-04   9.50e+00   1.00e+00  3.00e+04
   2  5.012552e-16    4.51e-07    3.17e-08   9.50e-04   1.00e+00  9.00e+04
"""
POWELL = """
   0  1.075000e+02    0.00e+00    1.55e+02   0.00e+00   0.00e+00  1.00e+04
   1  5.036190e+00    1.02e+02    2.00e+01   2.16e+00   9.53e-01  3.00e+04
   2  3.148168e-01    4.72e+00    2.50e+00   6.23e-01   9.37e-01  9.00e+04
   3  1.967760e-02    2.95e-01    3.13e-01   3.08e-01   9.37e-01  2.70e+05
   4  1.229900e-03    1.84e-02    3.91e-02   1.54e-01   9.37e-01  8.10e+05
   5  7.687123e-05    1.15e-03    4.89e-03   7.69e-02   9.37e-01  2.43e+06
   6  4.804625e-06    7.21e-05    6.11e-04   3.85e-02   9.37e-01  7.29e+06
   7  3.003028e-07    4.50e-06    7.64e-05   1.92e-02   9.37e-01  2.19e+07
   8  1.877006e-08    2.82e-07    9.54e-06   9.62e-03   9.37e-01  6.56e+07
   9  1.173223e-09    1.76e-08    1.19e-06   4.81e-03   9.37e-01  1.97e+08
  10  7.333425e-11    1.10e-09    1.49e-07   2.40e-03   9.37e-01  5.90e+08
  11  4.584044e-12    6.88e-11    1.86e-08   1.20e-03   9.37e-01  1.77e+09
  12  2.865573e-13    4.30e-12    2.33e-09   6.02e-04   9.37e-01  5.31e+09
  13  1.791438e-14    2.69e-13    2.91e-10   3.01e-04   9.37e-01  1.59e+10
  14  1.120029e-15    1.68e-14    3.64e-11   1.51e-04   9.37e-01  4.78e+10
"""


def _printed(rows):
    """Format like Ceres' progress table so the comparison is on the printed digits."""
    return ['%4d  %.6e    %.2e    %.2e   %.2e   %.2e  %.2e' % (int(r[0]), *r[1:]) for r in rows]


@pytest.mark.parametrize('problem,table', [('hello_world', HELLO_WORLD), ('powell', POWELL)])
def test_minimiser_reproduces_the_ceres_tutorial_tables(oracle, problem, table):
    rows, x, summary = oracle.ceres_tutorial_trace(problem)
    expected = [ln for ln in table.strip('\n').split('\n')]
    got = _printed(rows)
    assert len(got) == len(expected)
    for g, e in zip(got, expected):
        assert g.split() == e.split(), (g, e)
    assert summary[0] == 0  # CONVERGENCE
    if problem == 'hello_world':
        # "x : 0.5 -> 10", "Iterations: 2": a third step ends on the parameter tolerance and is not adopted
        assert '%.6g' % x[0] == '10' and summary[1] == 2
    else:
        # "Final x1 = 0.000146222, x2 = -1.46222e-05, x3 = 2.40957e-05, x4 = 2.40957e-05";
        # "Gradient tolerance reached. Gradient max norm 3.642190e-11 <= 1.000000e-10"
        assert ['%.6g' % v for v in x] == ['0.000146222', '-1.46222e-05', '2.40957e-05', '2.40957e-05']
        assert '%.6e' % summary[2] == '3.642190e-11' and summary[1] == 14


@pytest.fixture(scope='session')
def lm_dense_kat():
    import ctypes
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    src = os.path.join(here, 'harness', 'lm_dense_kat_harness.cpp')
    hdr = os.path.join(os.path.dirname(here), 'monorun_b200', 'csrc', 'lm_dense.cuh')
    out = os.path.join(here, 'harness', 'liblm_dense_kat_harness.so')
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(['/usr/bin/g++', '-O2', '-fPIC', '-std=c++17', '-Wno-unknown-pragmas', '-shared',
                               '-I', os.path.dirname(hdr), '-o', out, src])
    return ctypes.CDLL(out)


@pytest.mark.parametrize('problem,table,n', [('hello_world', HELLO_WORLD, 1), ('powell', POWELL, 4)])
def test_kernel_controller_reproduces_the_ceres_tutorial_tables(lm_dense_kat, problem, table, n):
    """The trust-region controller the fp64 CUDA kernels execute (monorun_b200/csrc/lm_dense.cuh: normal equations +
    Cholesky instead of Ceres' QR), compiled for the host, on the same two problems: cost, cost_change, |gradient| and
    |step| of every iteration as printed in Ceres' documentation (tr_ratio / tr_radius are internal to it)."""
    import ctypes
    fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rows, x, summary = np.zeros((64, 2 + n)), np.zeros(n), np.zeros(3)
    k = getattr(lm_dense_kat, 'lm_dense_kat_' + problem)(fp(rows), 64, fp(x), fp(summary))
    expected = [ln.split() for ln in table.strip('\n').split('\n')]
    assert k == len(expected)
    for i in range(k):
        step = np.linalg.norm(rows[i, 2:] - rows[i - 1, 2:]) if i else 0.0
        change = rows[i - 1, 0] - rows[i, 0] if i else 0.0
        got = ('%d  %.6e  %.2e  %.2e  %.2e' % (i, rows[i, 0], change, rows[i, 1], step)).split()
        assert got == expected[i][:5], (got, expected[i])
    assert summary[0] == 0 and summary[1] == k - 1
    if problem == 'powell':
        assert ['%.6g' % v for v in x] == ['0.000146222', '-1.46222e-05', '2.40957e-05', '2.40957e-05']


def test_kernel_controller_follows_the_oracle_minimiser_on_random_curve_fits(oracle, lm_dense_kat):
    """Normal equations + Cholesky (the kernels' controller) against Householder QR (the oracle's, i.e. Ceres') on 200
    random 4-parameter curve fits from poor starts -- ill-conditioned, ~25 iterations each, a third of them running
    into the 50-iteration limit: same termination, iteration and evaluation counts on >= 97 %, the final cost to 1e-7
    and the parameters to 1e-5 wherever the counts agree (the squared condition number of the normal equations shows
    here; on the Jacobi-scaled PnP problems the two agree to 1e-10, tests/test_noc.py, tests/test_6dof.py)."""
    import ctypes
    fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rng = np.random.default_rng(11)
    same, terms = 0, set()
    for trial in range(200):
        m = int(rng.integers(12, 40))
        t = np.sort(rng.uniform(-1.0, 2.0, m))
        truth = np.array([rng.uniform(0.5, 3.0), rng.uniform(-1.5, 1.2), rng.normal(0, 2.0), rng.normal(0, 1.0)])
        y = truth[0] * np.exp(truth[1] * t) + truth[2] + truth[3] * t + rng.normal(0, 0.05, m)
        x0 = truth + rng.normal(0, 1.0, 4) * np.array([1.0, 0.8, 2.0, 1.0])
        xo, so = oracle.expfit(t, y, x0)
        xk, sk = x0.copy(), np.zeros(4)
        lm_dense_kat.lm_dense_expfit(fp(t), fp(y), m, fp(xk), fp(sk))
        terms.add(int(so[0]))
        if tuple(so[:3]) == tuple(sk[:3]):
            same += 1
            assert np.abs(xk - xo).max() <= 1e-5 * max(1.0, np.abs(xo).max())
            assert abs(sk[3] - so[3]) <= 1e-7 * so[3]
        else:   # a decision on a tolerance boundary: both must still have reached the same cost level
            assert abs(sk[3] - so[3]) <= 1e-3 * so[3]
    assert same >= 194, same
    assert terms == {0, 1}   # both CONVERGENCE and NO_CONVERGENCE (iteration limit) exits were exercised
