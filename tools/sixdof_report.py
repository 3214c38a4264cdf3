"""6-DoF solver: the MIXED kernel (pnp_6dof_fast.cuh) against the fp64 kernel and the CPU oracle on seeded cases, and
the timing of both kernels on 8192 objects x 784 points (the case of tools/bench_noc.py).  Prints one JSON line per leg."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monorun_b200 import pnp  # noqa: E402
from oracle import sixdof_driver as sd  # noqa: E402
from tests.sixdof_cases import make_case, oracle_solve, rodrigues  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def solve(c, full, precision, mask=None, layout='planar'):
    c3, c2, w = dev(c['c3']), dev(c['c2']), dev(c['w'])
    if layout == 'planar':
        c3, c2, w = (t.permute(0, 2, 1).contiguous() for t in (c3, c2, w))
    res = pnp.solve_6dof_batched(c3, c2, w, dev(c['cam']), dev(c['uv_range']), dev(c['init']),
                                 dev(mask) if mask is not None else None, layout=layout,
                                 weight_mode='full' if full else 'istd', precision=precision)
    torch.cuda.synchronize()
    return res.cpu().numpy()


def errors(g, pose, cov):
    t_rel = np.linalg.norm(g[:, 3:6] - pose[:, 3:], axis=1) / np.linalg.norm(pose[:, 3:], axis=1)
    Rg, Rr = rodrigues(g[:, :3]), rodrigues(pose[:, :3])
    cosang = np.clip((np.einsum('nij,nij->n', Rg, Rr) - 1) / 2, -1, 1)
    rot = np.arccos(cosang)
    c = g[:, 6:42].reshape(-1, 6, 6)
    cov_rel = np.linalg.norm(c - cov, axis=(1, 2)) / np.linalg.norm(cov, axis=(1, 2))
    return t_rel, rot, cov_rel


def main():
    sd.build()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    for full in ((False, True) if n else ()):
        for far in (False, True):
            for masked in (False, True):
                c = make_case(n, full=full, far=far)
                mask = None
                if masked:
                    rng = np.random.default_rng(4)
                    mask = rng.uniform(size=c['c3'].shape[:2]) < rng.uniform(0.3, 1.0, (n, 1))
                r = oracle_solve(sd, c, full, mask=mask)
                g64 = solve(c, full, 'fp64', mask)
                gm = solve(c, full, 'mixed', mask)
                gi = solve(c, full, 'mixed', mask, layout='interleaved')
                ok = r['val'] & (g64[:, 42] > 0)
                t_rel, rot, cov_rel = errors(gm[ok], r['pose'][ok], r['cov'][ok])
                t64, rot64, cov64 = errors(gm[ok], g64[ok, :6], g64[ok, 6:42].reshape(-1, 6, 6))
                same_o = gm[:, 45] == r['stats'][:, 1]
                same_k = gm[:, 45] == g64[:, 45]
                print(json.dumps(dict(
                    leg='parity', full=full, far=far, masked=masked, n=n, valid_equal=bool(((gm[:, 42] > 0) == r['val']).all()),
                    same_evals_vs_oracle=float(same_o.mean()), same_evals_vs_fp64_kernel=float(same_k.mean()),
                    t_rel_max=float(t_rel.max()), rot_max=float(rot.max()), cov_rel_max=float(cov_rel.max()),
                    t_rel_max_same=float(t_rel[same_o[ok]].max()), rot_max_same=float(rot[same_o[ok]].max()),
                    t_rel_p999=float(np.quantile(t_rel, 0.999)), n_t_over_1e4=int((t_rel > 1e-4).sum()), n_rot_over_1e3=int((rot > 1e-3).sum()),
                    vs_fp64_kernel=dict(t_rel_max=float(t64.max()), rot_max=float(rot64.max()), cov_rel_max=float(cov64.max())),
                    cost_rel_max=float((np.abs(gm[:, 44] - r['cost']) / r['cost']).max()),
                    layouts_identical=bool(np.array_equal(gm, gi)))), flush=True)
    # timing: 8192 objects x 784 points, all points used, close start (tools/bench_noc.py's case)
    big = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    for full in (False, True):
        c = make_case(big, full=full, cfg=3, mode='S1')
        d = {k: dev(c[k]) for k in ('c3', 'c2', 'w', 'cam', 'uv_range', 'init')}
        pl = [d[k].permute(0, 2, 1).contiguous() for k in ('c3', 'c2', 'w')]
        out = {}
        for prec in (('mixed',) if 'mixedonly' in sys.argv else ('fp64', 'mixed')):
            def run():
                return pnp.solve_6dof_batched(pl[0], pl[1], pl[2], d['cam'], d['uv_range'], d['init'], layout='planar',
                                              weight_mode='full' if full else 'istd', precision=prec)
            for _ in range(3):
                res = run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                res = run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            out[prec] = res.cpu().numpy()
            bytes_alg = big * (784 * 4 * (8 if full else 7) + 96)
            print(json.dumps(dict(leg='timing', full=full, precision=prec, n=big, ms=ms, objects_per_s=big / ms * 1e3,
                                  gb_per_s=bytes_alg / ms / 1e6, mean_cost_evals=float(out[prec][:, 45].mean()))), flush=True)
        if 'fp64' not in out:
            continue
        t_rel, rot, cov_rel = errors(out['mixed'], out['fp64'][:, :6], out['fp64'][:, 6:42].reshape(-1, 6, 6))
        print(json.dumps(dict(leg='timing-parity', full=full, same_evals=float((out['mixed'][:, 45] == out['fp64'][:, 45]).mean()),
                              t_rel_max=float(t_rel.max()), rot_max=float(rot.max()), cov_rel_max=float(cov_rel.max()))), flush=True)


if __name__ == '__main__':
    main()
