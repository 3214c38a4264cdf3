#!/usr/bin/env python
"""bench.py -- throughput of the batched uncertainty-PnP hot path (BASELINE.json metric: PnP objects/s at 28x28
correspondences, % of the HBM roofline, next to the reference-equivalent CPU path on the same box).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload full|diag] [--impl ours|reference]

A "step" is one pass of the hot path over one batch: 8192 objects per GPU (BASELINE.json configs[2]; configs[4]
shards 65,536 objects over 8 GPUs = the same 8192 per GPU, so scaling is weak), each with 784 correspondences,
full 2x2 per-pixel covariance and pose-covariance output.  One solver launch per step (precision 'fast': the objects the
fp32 path hands back are solved by the exact fp64 routine inside the same launch); for N > 1 the step also contains the
gather of the [N_local, 24] result rows (peer-to-peer stores from the kernel + completion flags, or one NCCL all-gather).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OBJ_PER_GPU = 8192
ALG_BYTES = {'full': 8 * 784 * 4 + 96, 'diag': 7 * 784 * 4 + 96}  # reads + 96 B result row
METRIC = 'pnp_objects_per_s'
UNIT = 'objects/s'


def measured_peak():
    try:
        d = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json, STREAM-style copy, burst)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(workload):
    """dram__bytes_read+write per launch from the committed ncu capture of this workload, if present."""
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))[workload]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '100'], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        load = [s for s in sm if mx and s > 0.5 * mx] or sm
        return {'sm_mhz': float(np.median(load)) if load else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def make_inputs(workload, n, config, rank, nsets):
    from monorun_b200 import synth
    sets = []
    for s in range(nsets):
        b = synth.make_batch(n, config=config, rank=rank * 16 + s, weights=workload, mode='S1',
                             classes=(0, 1, 2) if workload == 'full' else (0,))
        sets.append(b)
    return sets


def cpu_threads():
    """Host threads for the CPU legs: every core this process may run on, stated explicitly -- torchrun exports
    OMP_NUM_THREADS=1, which must not shrink the reference arm."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_pack(od, b, workload, max_objects=None):
    """fp64 buffers of the oracle for (a prefix of) one batch, packed ONCE outside any timed loop."""
    from monorun_b200 import synth
    op = synth.to_op_level(b)
    full = workload == 'full'
    w = op['w_full'] if full else op['coords_2d_istd']
    n = w.shape[0] if max_objects is None else min(max_objects, w.shape[0])
    mask = od.istd_inlier_masks(w[:n][..., [0, 2]] if full else w[:n], 0.6)
    mask[mask.sum(1) <= 4] = True
    clips = np.array([[0.5, op['u_range'][0, 0], op['u_range'][0, 1], op['v_range'][0, 0], op['v_range'][0, 1]]])
    pk = od.pack_lm_batch(op['coords_2d'][:n], op['coords_3d'][:n], w[:n], op['cam_mats'], b['init_pose'][:n], clips, mask,
                          full_w=full)
    return pk, n, op, mask


def cpu_lm_rate(od, b, workload, threads, min_seconds, max_objects=None):
    """Oracle (restated Ceres LM + covariance, fp64) on the same workload: objects/s with `threads` OpenMP threads;
    the timed region is the native batched call only."""
    pk, n, _, _ = cpu_pack(od, b, workload, max_objects)
    od.lm_batch_packed(pk, with_pose_cov=True, threads=threads)  # warm-up (page-in, thread pool)
    done, t0 = 0, time.perf_counter()
    while True:
        od.lm_batch_packed(pk, with_pose_cov=True, threads=threads)
        done += n
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return done / dt, done, dt


def cpu_full_driver_rate(od, b, min_seconds, max_objects=256):
    """SURVEY 8d leg (ii): the reference's whole per-object CPU path, single thread -- istd inlier test, OpenCV
    EPnP-RANSAC (30 iterations, threshold 0.2 x RoI height) for the start and the inlier refinement, then the native LM
    (restated driver of pnp_uncert_cpu.py:11-125, 128-209; diagonal weights: the reference's Python path has no other)."""
    from monorun_b200 import synth
    op = synth.to_op_level(b)
    n = min(max_objects, op['coords_2d'].shape[0])
    v = op['coords_2d'][:n, :, 1]
    thres = 0.2 * (v[:, -1] - v[:, 0])   # uncert_prop_pnp_optimizer.py:86-88: v of the last row minus v of the first
    args = (op['coords_2d'][:n], op['coords_2d_istd'][:n], op['coords_3d'][:n], op['cam_mats'], op['u_range'], op['v_range'])
    kw = dict(z_min=0.5, epnp_istd_thres=0.6, epnp_ransac_thres=thres, inlier_opt_only=True)
    od.pnp_uncert_ref(*(a[:8] if a.shape[0] == n else a for a in args), **dict(kw, epnp_ransac_thres=thres[:8]))
    done, t0 = 0, time.perf_counter()
    while True:
        od.pnp_uncert_ref(*args, **kw)
        done += n
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return done / dt, done, dt


def base_config(args, n_local, world):
    """The workload description both arms print (the driver compares the two `config` objects)."""
    return {'workload': workload_name(args.workload), 'objects_per_gpu': n_local, 'points_per_object': 784,
            'objects_per_step': n_local * world}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  Its native op needs ceres-solver 1.14
    (absent, not installable), so this is the oracle port on every host core this process may use; a step is the same
    batch as the GPU arm's (objects_per_gpu x n_gpus objects of the same generator), packed once outside the timed
    region."""
    if rank != 0:
        return
    from oracle import pnp_driver as od
    od.build()
    threads = cpu_threads()
    n_local = objects_per_gpu(args, world)
    b = make_inputs(args.workload, n_local, 3 if args.workload == 'full' else 2, 0, 1)[0]
    pk, n, _, _ = cpu_pack(od, b, args.workload)
    steps = max(1, args.steps)
    for _ in range(max(min(args.warmup, 2), 1)):
        od.lm_batch_packed(pk, with_pose_cov=True, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        od.lm_batch_packed(pk, with_pose_cov=True, threads=threads)
    dt = time.perf_counter() - t0
    value = steps * n / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': args.warmup, 'ms_per_step': dt / steps * 1e3, 'higher_is_better': True,
        'scaling': 'strong' if args.total_objects else 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': base_config(args, n_local, world),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': f'{n} objects/step (one GPU\'s share of the same workload, same generator and seed), oracle LM + '
                                   f'covariance from the shared init, {threads} OpenMP threads stated explicitly, fp64 buffers '
                                   'packed once outside the timed region (reference native op needs ceres 1.14: absent)'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    emit(line)


def objects_per_gpu(args, world):
    if args.total_objects:
        if args.total_objects % world:
            raise SystemExit('--total-objects must be a multiple of the number of GPUs')
        return args.total_objects // world
    return OBJ_PER_GPU


def workload_name(w):
    return ('cfg3: 8192 mixed-class ROIs/GPU x 28x28 corr, full 2x2 per-pixel covariance, pose-cov out' if w == 'full'
            else 'cfg2-style: 8192 car ROIs/GPU x 28x28 corr, diagonal covariance (log-std in), pose-cov out')


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else a library prints there (NCCL's version banner ...)
    was redirected to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + '\n').encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--workload', choices=['full', 'diag'], default='full')
    ap.add_argument('--precision', choices=['fast', 'mixed', 'fp64'], default='fast')
    ap.add_argument('--impl', choices=['ours', 'reference'], default='ours')
    ap.add_argument('--streams', type=int, default=2, help='CUDA streams the K timed steps alternate between')
    ap.add_argument('--gather-lag', type=int, default=1,
                    help='fused gather: steps (per stream) between a solve and the wait for its rows')
    ap.add_argument('--gather', choices=['fused', 'fused-barrier', 'nccl'], default='fused',
                    help='N > 1: result rows by peer-to-peer stores from the kernel + symmetric-memory barrier, or NCCL all-gather')
    ap.add_argument('--large-batch', type=int, default=4,
                    help='N=1: also time one launch over this many concatenated batches (details.large_batch); 0/1 = skip')
    ap.add_argument('--no-sixdof', action='store_true', help='N=1: skip the 6-DoF leg (details.six_dof)')
    ap.add_argument('--total-objects', type=int, default=0,
                    help='strong scaling (BASELINE configs[4]: 65536): this many objects in total, split over the GPUs; '
                         'default 0 = 8192 objects per GPU (weak scaling)')
    ap.add_argument('--sustain-seconds', type=float, default=1.0,
                    help='after the K timed steps, keep solving back to back for this long and report the sustained rate with its clocks')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from monorun_b200 import pnp
    from monorun_b200 import dist as mdist
    from monorun_b200 import _native
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device; there is no CPU fallback')
    _native.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    full = args.workload == 'full'
    n_local = objects_per_gpu(args, world)
    n_total = n_local * world

    # ---- synthetic inputs: two alternating sets per rank (2 x 206 MB > 126 MB L2 at 8192 objects) ----
    sets = make_inputs(args.workload, n_local, 3 if full else 2, rank, 2)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    dsets = []
    for b in sets:
        ih, iw = b['img_shape']
        dsets.append(dict(c3=t(b['coords_3d']), c2=t(b['coords_2d']), w=t(b['w_full'] if full else b['logstd']),
                          cam=t(b['cam_mat'][None]), rng=torch.tensor([[-200., iw + 200., -200., ih + 200.]], device=dev),
                          init=t(b['init_pose'])))
    kw = dict(layout='planar', weight_mode='full' if full else 'logstd', precision=args.precision,
              cov_mode='pipeline', return_inlier_mask=False)

    # N > 1, --gather fused: the kernel stores the rows into every rank's symmetric buffer and raises completion flags;
    # the wait for step i (a one-warp kernel) is issued `--gather-lag` steps later on the same stream, so that no rank's
    # next solve waits for the slowest rank's current one.  streams x (lag + 1) buffers rotate.
    gathers, gather_note = [], ''
    nstreams = max(1, args.streams)
    lag = max(0, args.gather_lag) if args.gather == 'fused' else 0
    if world > 1 and args.gather in ('fused', 'fused-barrier'):
        try:
            gathers = [mdist.FusedGather(n_total, dev, signal='flags' if args.gather == 'fused' else 'barrier')
                       for _ in range(nstreams * (lag + 1))]
            ok = torch.ones(1, device=dev)
        except Exception as exc:  # symmetric memory unavailable on this box: NCCL all-gather instead, and say so
            gathers, gather_note, ok = [], f' (symmetric memory unavailable: {type(exc).__name__})', torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # every rank takes the same path
        if ok.item() == 0:
            gathers = []
    pending = []   # (stream index, FusedGather) of the solves whose rows have not been waited for yet

    def submit(i):
        """Step i on the current stream: solve (+ gather).  With the fused gather the wait for the rows of an EARLIER
        step of this stream is issued here; drain() issues the remaining waits."""
        d = dsets[i % 2]
        if gathers:
            fg = gathers[i % len(gathers)]
            pnp.solve_batched(d['c3'], d['c2'], d['w'], d['cam'], d['rng'], init_pose=d['init'], **kw, **fg.solve_kwargs())
            pending.append((i % nstreams, fg))
            mine = [k for k, (st, _) in enumerate(pending) if st == i % nstreams]
            if len(mine) > lag:
                return pending.pop(mine[0])[1].finish()
            return None
        rows, _, _ = pnp.solve_batched(d['c3'], d['c2'], d['w'], d['cam'], d['rng'], init_pose=d['init'], **kw)
        if world > 1:
            rows = mdist.all_gather_rows(rows, n_total)
        return rows

    def drain(streams=None):
        rows = None
        while pending:
            st, fg = pending.pop(0)
            if streams is None:
                rows = fg.finish()
            else:
                with torch.cuda.stream(streams[st]):
                    rows = fg.finish()
        return rows

    def step(i):
        rows = submit(i)
        last = drain()
        return last if last is not None else rows

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- multi-GPU: the fused gather (rows stored peer-to-peer by the kernel) must return, bit for bit, what one NCCL
    #      all-gather of the same solve returns -- checked on every rank before anything is timed ----
    gather_check = None
    if world > 1:
        d = dsets[0]
        rows_l, _, _ = pnp.solve_batched(d['c3'], d['c2'], d['w'], d['cam'], d['rng'], init_pose=d['init'], **kw)
        ref_rows = mdist.all_gather_rows(rows_l, n_total)
        got = step(0)
        fence()
        same = torch.equal(got, ref_rows) and torch.equal(ref_rows[rank * n_local:(rank + 1) * n_local], rows_l)
        flag = torch.tensor([1.0 if same else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_check = {'method': ('fused peer-to-peer stores + ' + ('completion flags' if args.gather == 'fused' else 'barrier')) if gathers else 'nccl all_gather_into_tensor',
                        'against': 'one NCCL all-gather of the same solve', 'rows': n_total,
                        'result': 'bitwise' if flag.item() == 1.0 else 'MISMATCH'}
        if flag.item() != 1.0:
            raise SystemExit(f'gather check failed: {gather_check}')

    # ---- parity against the oracle before any number counts: EVERY object of rank 0's first batch ----
    parity = None
    if rank == 0:
        from monorun_b200 import synth
        from oracle import pnp_driver as od
        od.build()
        b = sets[0]
        d = dsets[0]
        hb0 = pnp.handed_back_count(dev) if args.precision == 'fast' else 0
        rows, inl, r64 = pnp.solve_batched(d['c3'], d['c2'], d['w'], d['cam'], d['rng'], init_pose=d['init'],
                                           **dict(kw, return_inlier_mask=True, return_fp64=True))
        handed_back = (pnp.handed_back_count(dev) - hb0) if args.precision == 'fast' else None
        op = synth.to_op_level(b)
        wgt = op['w_full'] if full else op['coords_2d_istd']
        clips = np.array([[0.5, op['u_range'][0, 0], op['u_range'][0, 1], op['v_range'][0, 0], op['v_range'][0, 1]]])
        ref = od.lm_batch(op['coords_2d'], op['coords_3d'], wgt, op['cam_mats'], b['init_pose'], clips,
                          inl.cpu().numpy(), full_w=full, threads=cpu_threads())
        r = r64.cpu().numpy()
        t_err = np.linalg.norm(r[:, 1:4] - ref['pose'][:, 1:], axis=1) / np.linalg.norm(ref['pose'][:, 1:], axis=1)
        y_err = np.abs((r[:, 0] - ref['pose'][:, 0] + np.pi) % (2 * np.pi) - np.pi)
        off = int(((t_err >= 1e-4) | (y_err >= 1e-3)).sum())
        parity = {'objects': int(n_local), 'objects_outside_tolerance': off, 'tolerance': 'translation 1e-4 relative, yaw 1e-3 rad',
                  'max_rel_translation_err': float(t_err.max()), 'max_yaw_err_rad': float(y_err.max()),
                  'different_evaluation_counts': int((r[:, 6].astype(int) != ref['stats'][:, 1]).sum()),
                  'handed_to_exact_routine': handed_back, 'valid': float(rows[:, 20].mean().item())}
        if off:
            raise SystemExit(f'parity check failed: {parity}')

    # ---- (1) the kernel alone: K serialized launches on one stream, CUDA events around each (roofline source) ----
    # The parity check above left the GPU idle for seconds: besides the W warm-up steps, keep it busy for a quarter of a
    # second (untimed) so that the launches below run at the clocks of a loaded device, as they do inside a long job.
    # (Local solves only: a time-based loop runs a different number of iterations on every rank, so it must not
    # contain anything the other ranks take part in.)
    t_warm = time.perf_counter()
    i = 0
    while i < args.warmup or time.perf_counter() - t_warm < 0.25:
        d = dsets[i % 2]
        pnp.solve_batched(d['c3'], d['c2'], d['w'], d['cam'], d['rng'], init_pose=d['init'], **kw)
        i += 1
        if i % 64 == 0:
            torch.cuda.synchronize(dev)
    fence()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    # two untimed launches go first, without a synchronise in between: an event pair recorded around a launch on an EMPTY
    # queue also measures the host's launch latency (~40 us here, first sample of profiles/r02_bench_*), not the kernel
    for i in range(2):
        d = dsets[i % 2]
        pnp.solve_batched(d['c3'], d['c2'], d['w'], d['cam'], d['rng'], init_pose=d['init'], **kw)
    for i in range(args.steps):
        d = dsets[i % 2]
        kev[i][0].record()
        rows, _, _ = pnp.solve_batched(d['c3'], d['c2'], d['w'], d['cam'], d['rng'], init_pose=d['init'], **kw)
        kev[i][1].record()
    fence()
    kernel_us = [1e3 * a.elapsed_time(b) for a, b in kev]
    kernel_ms = float(np.mean(kernel_us)) * 1e-3

    # ---- (1b) the same kernel on a 4x larger batch (informative, not the headline): with one warp per object the last
    #      few long Levenberg-Marquardt runs of a launch leave most SMs idle; the larger batch shows the rate the kernel
    #      sustains per object when that ramp-down is amortised ----
    large = None
    if world == 1 and args.large_batch > 1 and not args.total_objects:
        reps = args.large_batch
        cat = {k: torch.cat([dsets[j % 2][k] for j in range(reps)]) for k in ('c3', 'c2', 'w', 'init')}
        for _ in range(2):
            pnp.solve_batched(cat['c3'], cat['c2'], cat['w'], dsets[0]['cam'], dsets[0]['rng'], init_pose=cat['init'], **kw)
        fence()
        lev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(6)]
        for a, b in lev:
            a.record()
            pnp.solve_batched(cat['c3'], cat['c2'], cat['w'], dsets[0]['cam'], dsets[0]['rng'], init_pose=cat['init'], **kw)
            b.record()
        fence()
        lus = float(np.mean([1e3 * a.elapsed_time(b) for a, b in lev]))
        large = {'objects_per_launch': reps * n_local, 'us_per_launch': lus, 'us_per_8192_objects': lus / reps,
                 'hbm_roofline_frac': ALG_BYTES[args.workload] * reps * n_local / (lus * 1e-6) / 1e9 / measured_peak()[0]}
        del cat

    # ---- (1c) the 6-DoF extension (the north star's wording; the reference solves 4 DoF) on the same correspondences,
    #      started from (0, yaw, 0, t): mixed kernel, every point used (informative, not the headline) ----
    six = None
    if world == 1 and not args.total_objects and not args.no_sixdof:
        try:
            d = dsets[0]
            zero = torch.zeros_like(d['init'][:, :1])
            init6 = torch.cat([zero, d['init'][:, :1], zero, d['init'][:, 1:4]], 1).contiguous()
            run6 = lambda: pnp.solve_6dof_batched(d['c3'], d['c2'], d['w'], d['cam'], d['rng'], init6, layout='planar',
                                                  weight_mode='full' if full else 'logstd', precision='mixed')
            for _ in range(2):
                r6 = run6()
            fence()
            sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
            for a, b in sev:
                a.record()
                r6 = run6()
                b.record()
            fence()
            sus = float(np.mean([1e3 * a.elapsed_time(b) for a, b in sev]))
            six = {'kernel': 'mr6::pnp_6dof_mixed_kernel (fp64 cost chain, fp32 normal equations)', 'objects_per_launch': n_local,
                   'us_per_launch': sus, 'objects_per_s': n_local / (sus * 1e-6),
                   'valid_fraction': float((r6[:, 42] > 0).double().mean()), 'mean_cost_evaluations': float(r6[:, 45].mean()),
                   'hbm_roofline_frac': ALG_BYTES[args.workload] * n_local / (sus * 1e-6) / 1e9 / measured_peak()[0]}
        except Exception as exc:  # informative leg: never take the headline down with it
            six = {'error': f'{type(exc).__name__}: {exc}'[:300]}

    # ---- (2) device-resident throughput: exactly K steps between fences.  Consecutive steps are independent
    #      batches, so they alternate between `--streams` CUDA streams: the ramp-down of one persistent launch (a
    #      few long Levenberg-Marquardt runs on otherwise idle SMs) overlaps the ramp-up of the next one. ----
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
    main = torch.cuda.current_stream(dev)
    for i in range(args.warmup):
        with torch.cuda.stream(streams[i % nstreams]):
            submit(i)
    drain(streams)
    fence()
    launches0 = pnp.launch_count(dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    fence()
    ev[0].record(main)
    for st in streams:
        st.wait_event(ev[0])
    rows = None
    for i in range(args.steps):
        with torch.cuda.stream(streams[i % nstreams]):
            r = submit(i)
            rows = r if r is not None else rows
    r = drain(streams)   # the waits still outstanding: every step's rows are complete inside the timed region
    rows = r if r is not None else rows
    for st in streams:
        done = torch.cuda.Event()
        done.record(st)
        main.wait_event(done)
    ev[1].record(main)
    fence()
    launches = pnp.launch_count(dev) - launches0
    ms_total = ev[0].elapsed_time(ev[1])
    iters = rows[:n_local, 21] if world == 1 else rows[rank * n_local:(rank + 1) * n_local, 21]
    hist = torch.bincount(iters.to(torch.int64).clamp(0, 63)).cpu().tolist()
    valid_frac = float(rows[:, 20].mean().item())

    # ---- (2b) the same loop kept up for --sustain-seconds: does the rate hold under sustained clocks? ----
    sustained = None
    if args.sustain_seconds > 0:
        per_step = ms_total / args.steps * 1e-3
        k_sus = int(min(max(args.sustain_seconds / max(per_step, 1e-6), args.steps), 200000))
        if world > 1:   # every rank must run the SAME number of steps: they all take part in each step's gather
            kt = torch.tensor([k_sus], device=dev, dtype=torch.int64)
            dist.broadcast(kt, src=0)
            k_sus = int(kt.item())
        fence()
        ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev2[0].record(main)
        for st in streams:
            st.wait_event(ev2[0])
        for i in range(k_sus):
            with torch.cuda.stream(streams[i % nstreams]):
                submit(i)
        drain(streams)
        for st in streams:
            done = torch.cuda.Event()
            done.record(st)
            main.wait_event(done)
        ev2[1].record(main)
        fence()
        sus_ms = torch.tensor([ev2[0].elapsed_time(ev2[1])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(sus_ms, op=dist.ReduceOp.MAX)
        sustained = {'steps': k_sus, 'seconds': sus_ms.item() * 1e-3, 'value': n_total * k_sus / (sus_ms.item() * 1e-3), 'unit': UNIT}
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the host-buffer C ABI call: pinned host inputs, H2D + kernel + D2H every step ----
    hsets = []
    for b in sets:
        ih, iw = b['img_shape']
        hsets.append(dict(c3=torch.from_numpy(b['coords_3d']).pin_memory(), c2=torch.from_numpy(b['coords_2d']).pin_memory(),
                          w=torch.from_numpy(b['w_full'] if full else b['logstd']).pin_memory(),
                          cam=torch.from_numpy(b['cam_mat'][None].copy()),
                          rng=torch.tensor([[-200., iw + 200., -200., ih + 200.]]),
                          init=torch.from_numpy(b['init_pose']).pin_memory(),
                          out=torch.empty((n_local, pnp.RESULT_STRIDE)).pin_memory()))
    hkw = dict(device=local_rank, layout='planar', weight_mode='full' if full else 'logstd', precision=args.precision)
    e2e_steps = max(3, min(args.steps, 10))
    for i in range(2):
        h = hsets[i % 2]
        pnp.solve_host(h['c3'], h['c2'], h['w'], h['cam'], h['rng'], h['init'], result=h['out'], **hkw)
    fence()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        h = hsets[i % 2]
        pnp.solve_host(h['c3'], h['c2'], h['w'], h['cam'], h['rng'], h['init'], result=h['out'], **hkw)
    fence()
    e2e_s = time.perf_counter() - t0
    h2d = int(sum(hsets[0][k].numel() * 4 for k in ('c3', 'c2', 'w', 'init', 'cam', 'rng')))
    d2h = int(hsets[0]['out'].numel() * 4)

    # ---- max over ranks ----
    tv = torch.tensor([ms_total, kernel_ms, e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms, e2e_s = tv.tolist()

    if rank == 0:
        peak, peak_src = measured_peak()
        alg = ALG_BYTES[args.workload] * n_local
        achieved = alg / (kernel_ms * 1e-3) / 1e9
        cfg = base_config(args, n_local, world)   # identical in both arms; everything else about this run is in `details`
        details = {}
        details.update({
            'precision': args.precision, 'init': 'ground truth perturbed (5e-2 rad, 2% depth), shared with the oracle',
            'l2': 'two alternating input sets of %.0f MB each (%s 126 MB L2)' % (alg / 1e6, '>' if alg > 126e6 else 'NOT larger than the'),
            'streams': nstreams, 'serialized_ms_per_step': kernel_ms,
            'serialized_us_per_launch': [round(v, 1) for v in kernel_us],
            'large_batch': large,
            'six_dof': six,
            'parallelism': f'objects sharded contiguously over {world} GPU(s)' + (
                '' if world == 1 else (', result rows stored peer-to-peer into every rank\'s symmetric buffer by the kernel; completion flags '
                                       f'raised by the kernel, waited for {lag} step(s) later on the same stream (1 one-warp launch per step)'
                                       if args.gather == 'fused' else
                                       ', result rows stored peer-to-peer into every rank\'s symmetric buffer by the kernel + 1 symmetric-memory barrier per step')
                if gathers else ', 1 NCCL all-gather of [N,24] rows per step' + gather_note)})
        if world > 1:
            details['p2p_bytes_per_step_per_gpu'] = int(n_local * 96 * (world - 1))
            if gathers and args.gather == 'fused':
                details['gather_flag_timeouts'] = pnp.gather_timeouts(dev)
                if details['gather_flag_timeouts']:
                    raise SystemExit(f"fused gather: {details['gather_flag_timeouts']} flag waits timed out")
        line = {
            'metric': METRIC, 'value': n_total * args.steps / (ms_total * 1e-3), 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_total / args.steps,
            'higher_is_better': True, 'scaling': 'strong' if args.total_objects else 'weak', 'vs_baseline': None,
            'dtype': {'fast': 'f32 (residuals evaluated once in f64, then tracked in packed f32; f64 covariance; borderline objects in f64)',
                      'mixed': 'f64 residual/cost + f32 Jacobian', 'fp64': 'f64'}[args.precision], 'data': 'synthetic',
            'config': cfg, 'details': details,
            'clocks': clocks,
            'e2e': {'value': n_total * e2e_steps / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': h2d * world,
                    'd2h_bytes_per_step': d2h * world, 'steps': e2e_steps,
                    'path': 'mrpnp_solve_host: pinned host tensors -> chunked H2D overlapped with the kernel -> D2H of result rows'},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': (ncu_traffic(args.workload) * n_local / OBJ_PER_GPU) if ncu_traffic(args.workload) else None, 'peak_source': peak_src,
                         'kernel': 'mrpnp::pnp_lm_fast_kernel' if args.precision == 'fast' else 'mrpnp::pnp_lm_kernel',
                         'kernel_ms': kernel_ms,
                         'algorithmic_bytes_per_object': ALG_BYTES[args.workload], 'objects_per_launch': n_local},
            'lm_iterations_histogram': hist, 'valid_fraction': valid_frac, 'parity_check': parity,
        }
        if sustained:
            line['sustained'] = sustained
        if gather_check:
            line['gather_check'] = gather_check
        if world == 1 and not args.no_cpu_baseline:
            from oracle import pnp_driver as od
            from monorun_b200 import synth
            threads = cpu_threads()
            all_rate, done, dt = cpu_lm_rate(od, sets[0], args.workload, threads, 6.0)
            one_rate, done1, dt1 = cpu_lm_rate(od, sets[0], args.workload, 1, 4.0, max_objects=2048)
            diag_b = sets[0] if not full else synth.make_batch(256, config=2, rank=0, weights='diag', mode='S1')
            drv_rate, done2, dt2 = cpu_full_driver_rate(od, diag_b, 5.0)
            line['cpu_baseline'] = {
                'value': all_rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                'sample': f'{done} object solves of this workload in {dt:.1f} s: oracle (restated Ceres-1.14 LM + covariance, '
                          f'fp64) from the shared init, OpenMP over objects on {threads} host threads of {os.cpu_count()}, fp64 buffers '
                          f'packed outside the timed region; single thread (what the reference does, pnp_uncert_cpu.py:180-191): '
                          f'{one_rate:.0f} objects/s ({done1} solves in {dt1:.1f} s); the reference\'s whole per-object driver on one '
                          f'thread (istd test + OpenCV EPnP-RANSAC + LM + covariance, pnp_uncert_cpu.py:11-125, diagonal weights): '
                          f'{drv_rate:.0f} objects/s ({done2} objects in {dt2:.1f} s)',
                'single_thread_value': one_rate, 'single_thread_full_driver_value': drv_rate}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
