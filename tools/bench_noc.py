"""Times mrpnp_solve_noc (7-parameter Huber solvers, fp64 kernel) on the B200 next to the CPU oracle.
Usage: python tools/bench_noc.py [--n 8192] [--out gpurun_out/noc_bench.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monorun_b200 import pnp  # noqa: E402
from tests.noc_cases import make_case, oracle_solve  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=8192)
    ap.add_argument('--cpu-sample', type=int, default=512)
    ap.add_argument('--out', default='gpurun_out/noc_bench.json')
    a = ap.parse_args()
    out = {}
    for full in (False, True):
        c = make_case(a.n, full=full, mode='S1', cfg=3)
        d = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in c.items()
             if k in ('noc', 'c2', 'w', 'logdim', 'logdim_wgt', 'cam', 'uv_range', 'init')}
        planar = [d[k].permute(0, 2, 1).contiguous() for k in ('noc', 'c2', 'w')]

        def run():
            return pnp.solve_noc_batched(planar[0], planar[1], planar[2], d['logdim'], d['logdim_wgt'], d['cam'],
                                         d['uv_range'], d['init'], layout='planar',
                                         weight_mode='full' if full else 'istd', huber_delta=1.5)
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
        res = res.cpu().numpy()
        from oracle import noc_driver as nd
        m = a.cpu_sample
        cs = {k: (v[:m] if isinstance(v, np.ndarray) and v.shape[0] == a.n else v) for k, v in c.items()}
        t0 = time.perf_counter()
        r = oracle_solve(nd, cs, 1.5, full, threads=0)
        cpu_s = time.perf_counter() - t0
        same = (res[:m, 10] == r['stats'][:, 1]).mean()
        out['full' if full else 'diag'] = dict(
            n=a.n, ms=ms, objects_per_s=a.n / ms * 1e3, mean_cost_evals=float(res[:, 10].mean()),
            cpu_objects_per_s=m / cpu_s, cpu_threads=os.cpu_count(), cpu_sample=m, same_decisions=float(same),
            max_param_diff=float(np.abs(res[:m, :7] - r['dimpose']).max()), valid=float((res[:, 7] > 0).mean()))
        print(out)
    # 6-DoF extension (mrpnp_solve_6dof)
    from tests.sixdof_cases import make_case as make6, oracle_solve as oracle6
    from oracle import sixdof_driver as sd
    for full in (False, True):
        c = make6(a.n, full=full, cfg=3, mode='S1')
        d = {k: torch.from_numpy(np.ascontiguousarray(c[k])).cuda() for k in ('c3', 'c2', 'w', 'cam', 'uv_range', 'init')}
        pl = [d[k].permute(0, 2, 1).contiguous() for k in ('c3', 'c2', 'w')]

        def run6(precision='fp64'):
            return pnp.solve_6dof_batched(pl[0], pl[1], pl[2], d['cam'], d['uv_range'], d['init'], layout='planar',
                                          weight_mode='full' if full else 'istd', precision=precision)

        def time6(precision):
            for _ in range(3):
                res = run6(precision)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                res = run6(precision)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / 10, res
        ms_mixed, res_mixed = time6('mixed')
        ms, res = time6('fp64')
        res, rm = res.cpu().numpy(), res_mixed.cpu().numpy()
        m = a.cpu_sample
        cs = {k: (v[:m] if isinstance(v, np.ndarray) and v.shape[0] == a.n else v) for k, v in c.items()}
        t0 = time.perf_counter()
        r = oracle6(sd, cs, full, threads=0)
        cpu_s = time.perf_counter() - t0
        same = (res[:m, 45] == r['stats'][:, 1]) & (np.abs(res[:m, 44] - r['cost']) <= 1e-9 * r['cost'])
        out['6dof_' + ('full' if full else 'diag')] = dict(
            n=a.n, ms=ms, objects_per_s=a.n / ms * 1e3, mean_cost_evals=float(res[:, 45].mean()),
            cpu_objects_per_s=m / cpu_s, cpu_threads=os.cpu_count(), cpu_sample=m, same_lm_paths=float(same.mean()),
            max_pose_diff_same_paths=float(np.abs(res[:m, :6] - r['pose'])[same].max()), valid=float((res[:, 42] > 0).mean()),
            mixed_ms=ms_mixed, mixed_objects_per_s=a.n / ms_mixed * 1e3,
            mixed_same_evals_as_fp64=float((rm[:, 45] == res[:, 45]).mean()),
            mixed_max_pose_diff_same_evals=float(np.abs(rm[:, :6] - res[:, :6])[rm[:, 45] == res[:, 45]].max()))
    print({k: v for k, v in out.items() if k.startswith('6dof')})

    # second-order covariance pass (mrpnp_exact_hessian): one read of the correspondences at the final pose
    c = make_case(a.n, mode='S1', cfg=3)
    metric = torch.from_numpy((c['coords_3d']).astype(np.float32)).cuda().permute(0, 2, 1).contiguous()
    c2 = torch.from_numpy(c['c2']).cuda().permute(0, 2, 1).contiguous()
    w = torch.from_numpy(c['w']).cuda().permute(0, 2, 1).contiguous()
    cam, rg = torch.from_numpy(c['cam']).cuda(), torch.from_numpy(c['uv_range']).cuda()
    rows = torch.zeros((a.n, 24), device='cuda')
    rows[:, :4] = torch.from_numpy(c['init_pose'].astype(np.float32)).cuda()
    rows[:, 20] = 1
    for layout, t3, t2, tw in (('planar', metric, c2, w),
                               ('interleaved', metric.permute(0, 2, 1).contiguous(), c2.permute(0, 2, 1).contiguous(),
                                w.permute(0, 2, 1).contiguous())):
        def run_xh():
            return pnp.exact_hessian(t3, t2, tw, cam, rg, rows, None, layout=layout, weight_mode='istd', rows=rows,
                                     return_hessian=False)
        for _ in range(3):
            run_xh()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run_xh()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        nbytes = a.n * t3.shape[1 if layout == 'interleaved' else 2] * 7 * 4
        out['exact_hessian_' + layout] = dict(n=a.n, ms=ms, objects_per_s=a.n / ms * 1e3, bytes=nbytes,
                                              gb_per_s=nbytes / ms / 1e6)
    print({k: v for k, v in out.items() if k.startswith('exact')})
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
