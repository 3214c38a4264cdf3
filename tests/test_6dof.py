"""CPU tests of the 6-DoF extension.  No reference exists for it (the reference's rotation vector is hard-wired to
(0, yaw, 0) and ``use_6dof`` is never read), so the oracle (oracle/pnp_6dof_oracle.cpp: dual-number Jacobians through a
restatement of ceres::AngleAxisRotatePoint) is checked against the 4-DoF oracle, finite differences, an independent
minimiser and noise-free recovery; the solver logic the CUDA kernel executes (closed-form rotation derivative + normal
equations, compiled for the host by tests/harness/) is checked against that oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests.sixdof_cases import make_case, oracle_solve, rodrigues

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope='session')
def sd():
    from oracle import sixdof_driver
    sixdof_driver.build()
    return sixdof_driver


@pytest.fixture(scope='session')
def harness6():
    src = os.path.join(HERE, 'harness', 'sixdof_host_harness.cpp')
    hdrs = [os.path.join(ROOT, 'monorun_b200', 'csrc', f) for f in ('pnp_6dof.cuh', 'pnp_6dof_fast.cuh', 'lm_dense.cuh')]
    out = os.path.join(HERE, 'harness', 'libsixdof_host_harness.so')
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(['/usr/bin/g++', '-O2', '-fPIC', '-std=c++17', '-Wno-unknown-pragmas', '-shared',
                               '-I', os.path.dirname(hdrs[0]), '-o', out, src])
    return ctypes.CDLL(out)


def harness_solve(lib, c, full, mask=None, init=None):
    n, p = c['c3'].shape[:2]
    res = np.zeros((n, 48))
    fp = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    m = np.ascontiguousarray(mask, np.uint8) if mask is not None else None
    init = np.ascontiguousarray(c['init'] if init is None else init, np.float32)
    lib.sixdof_host_harness(fp(c['c3']), fp(c['c2']), fp(c['w']), fp(m), fp(np.ascontiguousarray(c['cam'].reshape(-1, 9))),
                            0, fp(c['uv_range']), 0, fp(init), n, p, int(full), ctypes.c_double(0.5), fp(res))
    return res


def _cost(x, c, i, full):
    """Independent numpy statement of 1/2 |r|^2 (Rodrigues matrix, np.clip)."""
    K = c['cam'][0].astype(np.float64)
    z_min, u0, u1, v0, v1 = c['clips'][0]
    q = c['c3'][i].astype(np.float64) @ rodrigues(x[None, :3])[0].T + x[3:]
    z = np.maximum(q[:, 2], z_min)
    du = np.clip(K[0, 0] * q[:, 0] / z + K[0, 2], u0, u1) - c['c2'][i, :, 0]
    dv = np.clip(K[1, 1] * q[:, 1] / z + K[1, 2], v0, v1) - c['c2'][i, :, 1]
    w = c['w'][i].astype(np.float64)
    if full:
        r = np.stack([w[:, 0] * du + w[:, 1] * dv, w[:, 1] * du + w[:, 2] * dv])
    else:
        r = np.stack([w[:, 0] * du, w[:, 1] * dv])
    return 0.5 * (r ** 2).sum()


@pytest.mark.parametrize('full', [False, True])
def test_rotation_about_y_reproduces_the_4dof_objective(oracle, sd, full):
    """At r_vec = (0, yaw, 0) cost, the (yaw, t) gradient and the (yaw, t) block of J^T J equal the 4-DoF oracle's."""
    c = make_case(4, full=full)
    for i in range(4):
        p4 = c['init4'][i].astype(np.float64)
        p6 = np.array([0.0, p4[0], 0.0, p4[1], p4[2], p4[3]])
        c4, g4, h4 = oracle.eval_cost_grad_hess(c['coords_2d_yaw'][i], c['c3'][i], c['w'][i], c['cam'][0], p4,
                                                c['clips'][0], full_w=full)
        c6, g6, h6 = sd.eval_cost_grad_hess(c['coords_2d_yaw'][i], c['c3'][i], c['w'][i], c['cam'][0], p6, c['clips'][0],
                                            full_w=full)
        sel = [1, 3, 4, 5]
        assert abs(c4 - c6) <= 1e-12 * c4
        np.testing.assert_allclose(g6[sel], g4, rtol=1e-10, atol=1e-10 * np.abs(g4).max())
        np.testing.assert_allclose(h6[np.ix_(sel, sel)], h4, rtol=1e-10, atol=1e-10 * np.abs(h4).max())


@pytest.mark.parametrize('full', [False, True])
def test_gradient_matches_finite_differences(sd, full):
    c = make_case(3, full=full, far=True)
    for i in range(3):
        for x in (c['init'][i].astype(np.float64), np.array([1e-9, -2e-9, 1e-9, *c['gt'][i, 3:]])):  # both branches
            cost, grad, jtj = sd.eval_cost_grad_hess(c['c2'][i], c['c3'][i], c['w'][i], c['cam'][0], x, c['clips'][0],
                                                     full_w=full)
            assert abs(cost - _cost(x, c, i, full)) <= 1e-9 * cost
            fd = np.zeros(6)
            for k in range(6):
                h = 1e-6 * max(1.0, abs(x[k]))
                e = np.zeros(6)
                e[k] = h
                fd[k] = (_cost(x + e, c, i, full) - _cost(x - e, c, i, full)) / (2 * h)
            np.testing.assert_allclose(grad, fd, rtol=5e-5, atol=1e-6 * np.abs(fd).max())


def test_noise_free_recovery_of_general_rotations(sd):
    c = make_case(32, noise=False, tilt=0.4, far=True)
    r = oracle_solve(sd, c, False)
    assert r['val'].all()
    np.testing.assert_allclose(r['pose'], c['gt'], atol=2e-5)


def test_minimum_agrees_with_an_independent_minimiser(sd):
    from scipy.optimize import minimize
    c = make_case(5)
    r = oracle_solve(sd, c, False)
    for i in range(5):
        ref = minimize(_cost, r['pose'][i], args=(c, i, False), method='BFGS', options=dict(gtol=1e-8, maxiter=300))
        assert r['cost'][i] <= ref.fun * (1 + 2e-6)
        np.testing.assert_allclose(r['pose'][i], ref.x, rtol=2e-3, atol=3e-3)


@pytest.mark.parametrize('full', [False, True])
@pytest.mark.parametrize('far', [False, True])
def test_kernel_logic_on_the_host_matches_oracle(sd, harness6, full, far):
    c = make_case(48, full=full, far=far)
    r = oracle_solve(sd, c, full)
    h = harness_solve(harness6, c, full)
    same = h[:, 45] == r['stats'][:, 1]
    assert same.mean() >= 0.95
    np.testing.assert_allclose(h[same, :6], r['pose'][same], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(h[:, 44], r['cost'], rtol=1e-5)
    np.testing.assert_array_equal(h[:, 42] > 0, r['val'])
    cov = h[same, 6:42].reshape(-1, 6, 6)
    rel = np.linalg.norm(cov - r['cov'][same], axis=(1, 2)) / np.linalg.norm(r['cov'][same], axis=(1, 2))
    assert rel.max() < 1e-7


def test_kernel_logic_first_order_branch_and_masks(sd, harness6):
    """Start at r_vec = 0 (AngleAxisRotatePoint's first-order branch) with ragged inlier masks."""
    c = make_case(32, tilt=0.05)
    init = c['init'].copy()
    init[:, :3] = 0.0
    rng = np.random.default_rng(4)
    mask = rng.uniform(size=c['c3'].shape[:2]) < rng.uniform(0.3, 1.0, (32, 1))
    r = oracle_solve(sd, c, False, mask=mask, init=init)
    h = harness_solve(harness6, c, False, mask=mask, init=init)
    same = h[:, 45] == r['stats'][:, 1]
    assert same.mean() >= 0.9 and r['val'].all()
    np.testing.assert_allclose(h[same, :6], r['pose'][same], rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize('full', [False, True])
def test_mixed_kernel_controller_is_minimize_cut_at_the_evaluation(harness6, full):
    """mr6::lm_advance (the mixed 6-DoF kernel's controller, run by lane 0 between evaluations) against mrlm::minimize on
    identical numbers -- fp64 cost, the 27 sums rounded to fp32 as the kernel's stash holds them: bit-equal pose, cost,
    iteration and evaluation counts and termination, near and far starts, masked and not, plus a start that fails
    (non-finite) and one that converges at once."""
    for far, masked in ((False, False), (True, False), (True, True)):
        n = 48
        c = make_case(n, full=full, far=far, seed=5)
        p = c['c3'].shape[1]
        init = np.ascontiguousarray(c['init'], np.float32).copy()
        init[0, 5] = np.nan                      # first evaluation not finite: FAILURE, pose returned untouched
        init[1] = c['gt'][1].astype(np.float32)  # (nearly) at the minimum
        mask = None
        if masked:
            rng = np.random.default_rng(3)
            mask = np.ascontiguousarray(rng.uniform(size=(n, p)) < rng.uniform(0.2, 1.0, (n, 1)), np.uint8)
        a, b = np.zeros((n, 10)), np.zeros((n, 11))
        fp = lambda x: x.ctypes.data_as(ctypes.c_void_p) if x is not None else None
        harness6.sixdof_controller_harness(fp(c['c3']), fp(c['c2']), fp(c['w']), fp(mask),
                                           fp(np.ascontiguousarray(c['cam'].reshape(-1, 9))), fp(c['uv_range']), fp(init), n, p,
                                           int(full), ctypes.c_double(0.5), fp(a), fp(b))
        np.testing.assert_array_equal(a[2:], b[2:, :10])
        np.testing.assert_array_equal(a[0, 6:], b[0, 6:10])          # NaN pose: compare the rest
        assert a[0, 9] == 2 and np.isnan(b[0, 5])                    # mrlm::kFailure
        np.testing.assert_array_equal(a[1], b[1, :10])
        assert (b[1:, 10] == 1).all()                                # returned pose = current stash entry
        assert (a[2:, 8] >= 3).all() and (a[2:, 9] == 0).all()       # real solves, all converged
    # wild starts (rotation off by ~1.5 rad, depth scaled 0.3-3x): rejected and invalid steps, shrinking radii, objects that
    # end far from the minimum or fail -- whatever mrlm::minimize does, lm_advance does the same
    n = 64
    c = make_case(n, full=full, far=True, seed=9)
    p = c['c3'].shape[1]
    rng = np.random.default_rng(17)
    init = np.ascontiguousarray(c['init'], np.float32).copy()
    init[:, :3] += rng.normal(0, 1.5, (n, 3)).astype(np.float32)
    init[:, 3:] *= rng.uniform(0.3, 3.0, (n, 1)).astype(np.float32)
    a, b = np.zeros((n, 10)), np.zeros((n, 11))
    fp = lambda x: x.ctypes.data_as(ctypes.c_void_p) if x is not None else None
    harness6.sixdof_controller_harness(fp(c['c3']), fp(c['c2']), fp(c['w']), None,
                                       fp(np.ascontiguousarray(c['cam'].reshape(-1, 9))), fp(c['uv_range']), fp(init), n, p,
                                       int(full), ctypes.c_double(0.5), fp(a), fp(b))
    np.testing.assert_array_equal(a, b[:, :10])
    assert (a[:, 8] > 8).sum() >= n // 4        # long runs are in the sample
    assert (a[:, 6] + 1 > a[:, 8]).any() or (a[:, 9] != 0).any() or (a[:, 8] > 15).any()
