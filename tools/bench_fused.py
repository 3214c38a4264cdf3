"""Head->PnP boundary, fused vs unfused, on one GPU.

unfused = the reference's launch sequence on device tensors (NOCCoder.decode, decode_logstd, RoI grid, then the
PnP launch, monorun_roi_head.py:513-529);  fused = one mrpnp_solve_dense launch on the head's raw maps.
Prints one JSON line; CUDA-event timing, two alternating input sets (> L2), W warm-up + K timed steps each.
    python tools/bench_fused.py [--objects 8192] [--steps 30] [--warmup 5]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monorun_b200 import coders, pnp, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--objects', type=int, default=8192)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--precision', default='fast')
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    cc, pc = coders.NOCCoder(synth.NOC_MEANS, synth.NOC_STDS), coders.DistanceInvarProjErrorCoder()
    sets = []
    for s in range(2):
        b = synth.make_batch(a.objects, config=3, rank=16 + s, weights='diag', mode='S1')
        raw = synth.to_head_raw(b, rng=np.random.default_rng(s))
        ih, iw = b['img_shape']
        sets.append(dict(noc=t(raw['noc_pred']), ls=t(raw['proj_logstd']), rois=t(raw['rois']), dims=t(raw['dims']),
                         dv=t(raw['dims_var']), cam=t(b['cam_mat'][None]), init=t(b['init_pose']),
                         rng=torch.tensor([[-200., iw + 200., -200., ih + 200.]], device=dev)))

    def unfused(d):
        c3, c3v = cc.decode(d['noc'], None, d['dims'], d['dv'], False)
        ls = pc.decode_logstd(d['ls'], c3v, None)
        c2 = coders.coords_2d_from_rois(d['rois'], 28).contiguous()
        return pnp.solve_batched(c3, c2, ls, d['cam'], d['rng'], init_pose=d['init'], layout='planar',
                                 weight_mode='logstd', precision=a.precision, return_inlier_mask=False)[0]

    def fused(d):
        return pnp.solve_dense(d['noc'], d['ls'], d['rois'], d['dims'], d['dv'], d['cam'], d['rng'],
                               noc_mean=cc.target_means, noc_std=cc.target_stds,
                               focal_gain=pc.ref_focal_y * pc.epistemic_std_gain,
                               scaling_denominator=pc.scaling_denomitor, init_pose=d['init'], precision=a.precision,
                               return_inlier_mask=False)[0]

    out = {}
    for name, fn in (('unfused', unfused), ('fused', fused)):
        for i in range(a.warmup):
            fn(sets[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            rows = fn(sets[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        out[name] = {'ms_per_step': ms, 'objects_per_s': a.objects / ms * 1e3, 'valid': float(rows[:, 20].mean())}
    ru, rf = unfused(sets[0]), fused(sets[0])
    t_rel = ((ru[:, 1:4] - rf[:, 1:4]).norm(dim=1) / ru[:, 1:4].norm(dim=1))
    out['agreement'] = {'median_rel_translation': float(t_rel.median()), 'p999_rel_translation': float(t_rel.quantile(0.999))}
    out['config'] = {'objects': a.objects, 'points_per_object': 784, 'precision': a.precision, 'steps': a.steps,
                     'warmup': a.warmup, 'input_bytes_per_object': {'unfused_pnp_launch': 7 * 784 * 4 + 16,
                                                                    'fused': 5 * 784 * 4 + 16 + 16 + 24}}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
