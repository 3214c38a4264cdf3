"""CPU tests of the 7-parameter solvers: the oracle restatement of pnp_noc_uncert /
pnp_noc_cov_uncert (oracle/pnp_noc_oracle.cpp) is checked against the 4-parameter oracle, finite differences and an
independent minimiser; the solver logic the CUDA kernel executes (the __host__ __device__ part of
monorun_b200/csrc/pnp_noc.cuh, compiled for the host by tests/harness/) is checked against the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests.noc_cases import make_case, oracle_solve

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope='session')
def noc_oracle():
    from oracle import noc_driver
    noc_driver.build()
    return noc_driver


@pytest.fixture(scope='session')
def harness():
    src = os.path.join(HERE, 'harness', 'noc_host_harness.cpp')
    hdr = os.path.join(ROOT, 'monorun_b200', 'csrc', 'pnp_noc.cuh')
    out = os.path.join(HERE, 'harness', 'libnoc_host_harness.so')
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(['/usr/bin/g++', '-O2', '-fPIC', '-std=c++17', '-Wno-unknown-pragmas', '-shared',
                               '-I', os.path.dirname(hdr), '-o', out, src])
    return ctypes.CDLL(out)


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def harness_solve(lib, c, delta, full, mask=None):
    n, p = c['noc'].shape[:2]
    res = np.zeros((n, 12))
    m = np.ascontiguousarray(mask, np.uint8) if mask is not None else None
    cam = np.ascontiguousarray(c['cam'].reshape(-1, 9))
    lib.noc_host_harness(_fp(c['noc']), _fp(c['c2']), _fp(c['w']), _fp(m), _fp(c['logdim']), _fp(c['logdim_wgt']),
                         _fp(cam), 0, _fp(c['uv_range']), 0, _fp(c['init']), n, p, int(full),
                         ctypes.c_double(0.5), ctypes.c_double(delta), _fp(res))
    return res


def _robust_cost(x, c, i, delta, full):
    """Independent numpy statement of the objective: 1/2 sum_blocks huber(|r_block|^2) (pnp_uncert_cpu.cpp:122-148,
    :189-217, :77-104; ceres::HuberLoss)."""
    K = c['cam'][0].astype(np.float64)
    z_min, u0, u1, v0, v1 = c['clips'][0]
    S = c['noc'][i].astype(np.float64) * np.exp(x[:3])
    cs, sn = np.cos(x[3]), np.sin(x[3])
    xc = cs * S[:, 0] + sn * S[:, 2] + x[4]
    yc = S[:, 1] + x[5]
    zc = np.maximum(-sn * S[:, 0] + cs * S[:, 2] + x[6], z_min)
    pu = np.clip(K[0, 0] * xc / zc + K[0, 2], u0, u1)
    pv = np.clip(K[1, 1] * yc / zc + K[1, 2], v0, v1)
    du, dv = pu - c['c2'][i, :, 0], pv - c['c2'][i, :, 1]
    w = c['w'][i].astype(np.float64)
    if full:
        r0, r1 = w[:, 0] * du + w[:, 1] * dv, w[:, 1] * du + w[:, 2] * dv
    else:
        r0, r1 = w[:, 0] * du, w[:, 1] * dv
    s = np.concatenate([r0 * r0 + r1 * r1,
                        [np.sum((c['logdim_wgt'][i].astype(np.float64) * (x[:3] - c['logdim'][i])) ** 2)]])
    rho = np.where(s > delta * delta, 2 * delta * np.sqrt(s) - delta * delta, s)
    return 0.5 * rho.sum()


def test_stiff_prior_without_loss_reduces_to_the_4_parameter_oracle(oracle, noc_oracle):
    """With the dimensions pinned by a stiff prior at the true values and delta -> inf, the 7-parameter problem is
    the 4-parameter one: same pose, same iteration / evaluation counts, same cost."""
    for full in (False, True):
        c = make_case(32, full=full, prior_sd=0.0, prior_wgt=1e6)
        r7 = oracle_solve(noc_oracle, c, 1e9, full, threads=1)
        metric = c['noc'].astype(np.float64) * np.exp(c['logdim'].astype(np.float64))[:, None, :]
        r4 = oracle.lm_batch(c['c2'], metric, c['w'], c['cam'], c['init'][:, 3:], c['clips'], full_w=full)
        np.testing.assert_allclose(r7['dimpose'][:, 3:], r4['pose'], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(r7['dimpose'][:, :3], c['logdim'], atol=1e-7)
        np.testing.assert_array_equal(r7['stats'], r4['stats'])
        np.testing.assert_allclose(r7['cost'], r4['cost'], rtol=1e-8)


@pytest.mark.parametrize('full', [False, True])
def test_gradient_matches_finite_differences_of_the_robust_cost(noc_oracle, full):
    c = make_case(4, full=full, far=True)
    for delta in (0.8, 3.0):
        for i in range(4):
            x = c['init'][i].astype(np.float64)
            cost, grad, jtj = noc_oracle.noc_eval(c['c2'][i], c['noc'][i], c['w'][i], c['logdim'][i],
                                                  c['logdim_wgt'][i], c['cam'][0], x, c['clips'][0], delta, full)
            assert abs(cost - _robust_cost(x, c, i, delta, full)) <= 1e-9 * cost
            fd = np.zeros(7)
            for k in range(7):
                h = 1e-6 * max(1.0, abs(x[k]))
                e = np.zeros(7)
                e[k] = h
                fd[k] = (_robust_cost(x + e, c, i, delta, full) - _robust_cost(x - e, c, i, delta, full)) / (2 * h)
            np.testing.assert_allclose(grad, fd, rtol=2e-4, atol=1e-4 * np.abs(fd).max())
            assert np.allclose(jtj, jtj.T) and np.all(np.linalg.eigvalsh(jtj) > -1e-9)


@pytest.mark.parametrize('full', [False, True])
def test_minimum_agrees_with_an_independent_minimiser(noc_oracle, full):
    from scipy.optimize import minimize
    c = make_case(6, full=full)
    delta = 1.5
    r = oracle_solve(noc_oracle, c, delta, full)
    assert r['val'].all()
    for i in range(6):
        ref = minimize(_robust_cost, r['dimpose'][i] + 1e-3, args=(c, i, delta, full), method='BFGS',
                       options=dict(gtol=1e-9, maxiter=500))
        # Ceres stops on its function tolerance (1e-6 relative), so the oracle sits within that of the optimum
        assert r['cost'][i] <= ref.fun * (1 + 2e-6)
        assert abs(_robust_cost(r['dimpose'][i], c, i, delta, full) - r['cost'][i]) <= 1e-9 * r['cost'][i]
        # the depth / size direction is nearly flat, so positions agree to ~1e-3 relative at that cost level
        np.testing.assert_allclose(r['dimpose'][i], ref.x, rtol=2e-3, atol=5e-3)


def test_solver_refines_the_dimensions(noc_oracle):
    c = make_case(64, prior_sd=0.15, prior_wgt=2.0)
    r = oracle_solve(noc_oracle, c, 2.0, False)
    err_prior = np.abs(c['logdim'] - np.log(c['dims'])).mean()
    err_post = np.abs(r['dimpose'][:, :3] - np.log(c['dims'])).mean()
    assert r['val'].all() and err_post < 0.8 * err_prior


def test_single_call_abi_equals_batch(noc_oracle):
    for full in (False, True):
        c = make_case(4, full=full)
        r = oracle_solve(noc_oracle, c, 1.5, full)
        for i in range(4):
            val, x = noc_oracle.noc_single(c['c2'][i], c['noc'][i], c['w'][i], c['logdim'][i], c['logdim_wgt'][i],
                                           c['cam'][0], c['init'][i], c['clips'][0], 1.5, full_w=full)
            assert val
            np.testing.assert_array_equal(x, r['dimpose'][i])


@pytest.mark.parametrize('full', [False, True])
@pytest.mark.parametrize('mode,cfg', [('S0', 2), ('S1', 3)])
def test_kernel_logic_on_the_host_matches_oracle(noc_oracle, harness, full, mode, cfg):
    """The product's point functor, Huber corrector and trust-region controller (normal equations + Cholesky) against
    the oracle (explicit Jacobian + Householder QR): same decisions, parameters to 1e-8."""
    c = make_case(48, full=full, mode=mode, cfg=cfg)
    for delta in (0.5, 1.5, 1e9):
        r = oracle_solve(noc_oracle, c, delta, full)
        h = harness_solve(harness, c, delta, full)
        np.testing.assert_array_equal(h[:, 10], r['stats'][:, 1])
        np.testing.assert_array_equal(h[:, 8], r['stats'][:, 0])
        np.testing.assert_array_equal(h[:, 11], r['stats'][:, 3])
        np.testing.assert_allclose(h[:, :7], r['dimpose'], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(h[:, 9], r['cost'], rtol=1e-11)
        assert h[:, 7].all() and r['val'].all()


def test_kernel_logic_far_start_and_masks(noc_oracle, harness):
    """Far initialisation (rejected steps, clipped points) and ragged inlier masks."""
    c = make_case(64, far=True, seed=5)
    rng = np.random.default_rng(3)
    mask = rng.uniform(size=c['noc'].shape[:2]) < rng.uniform(0.2, 1.0, (64, 1))
    mask[0] = False
    mask[0, :3] = True  # an under-determined object: 3 points, 7 unknowns (the prior keeps it solvable)
    r = oracle_solve(noc_oracle, c, 1.0, False, mask=mask)
    h = harness_solve(harness, c, 1.0, False, mask=mask)
    assert (r['stats'][:, 1] > r['stats'][:, 2]).any(), 'no rejected step in the case'
    same = h[:, 10] == r['stats'][:, 1]
    assert same.mean() >= 0.95
    np.testing.assert_allclose(h[same, :7], r['dimpose'][same], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(h[:, 7] > 0, r['val'])
    # decisions that differ (normal equations vs QR at the function tolerance) still end at the same cost
    np.testing.assert_allclose(h[:, 9], r['cost'], rtol=1e-4)
