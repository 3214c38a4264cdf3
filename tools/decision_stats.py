"""GPU diagnostic: how often does a non-fp64 precision stop at a different LM step than the fp64 kernel (which
reproduces the oracle decision for decision), and how far are such objects from the north-star tolerance?
Usage: python tools/decision_stats.py [n] [precisions]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monorun_b200 import synth, pnp


def run(n, cfg, weights, prec, rank=0, bands=None):
    b = synth.make_batch(n, config=cfg, rank=rank, weights=weights, mode='S1', classes=(0, 1, 2) if weights == 'full' else (0,))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    full = weights == 'full'
    ih, iw = b['img_shape']
    rng = torch.tensor([[-200., iw + 200., -200., ih + 200.]], device='cuda')
    args = (t(b['coords_3d']), t(b['coords_2d']), t(b['w_full'] if full else b['logstd']), t(b['cam_mat'][None]), rng)
    kw = dict(init_pose=t(b['init_pose']), layout='planar', weight_mode='full' if full else 'logstd', return_fp64=True)
    out = {}
    for p in ['fp64'] + prec:
        res, inl, r64 = pnp.solve_batched(*args, precision=p, decision_bands=bands if p == 'fast' else None, **kw)
        torch.cuda.synchronize()
        out[p] = (res.cpu().numpy(), r64.cpu().numpy(), inl.cpu().numpy())
    ref = out['fp64'][1]
    line = {'n': n, 'cfg': cfg, 'weights': weights, 'rank': rank, 'bands': bands}
    for p in prec:
        r = out[p][1]
        terr = np.linalg.norm(r[:, 1:4] - ref[:, 1:4], axis=1) / np.linalg.norm(ref[:, 1:4], axis=1)
        yerr = np.abs((r[:, 0] - ref[:, 0] + np.pi) % (2 * np.pi) - np.pi)
        diff = r[:, 6] != ref[:, 6]
        off = (terr >= 1e-4) | (yerr >= 1e-3)
        line[p] = {'diff_evals': int(diff.sum()), 'off_tolerance': int(off.sum()), 'max_t': float(terr.max()),
                   'max_yaw': float(yerr.max()), 'off_and_same_evals': int((off & ~diff).sum()),
                   'mask_diff_points': int((out[p][2] != out['fp64'][2]).sum()),
                   'valid': float(out[p][0][:, 20].mean()),
                   'term_hist': np.bincount(r[:, 7].astype(int)).tolist(),
                   'evals_hist': np.bincount(r[:, 6].astype(int)).tolist()}
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    prec = sys.argv[2].split(',') if len(sys.argv) > 2 else ['mixed', 'fast']
    bands = [None]
    if len(sys.argv) > 3:
        bands = [None if b == 'default' else tuple(float(v) for v in b.split(':')) for b in sys.argv[3].split(',')]
    for b in bands:
        for rank in (0, 1, 2, 3):
            run(n, 3, 'full', prec, rank, b)
            run(n, 2, 'diag', prec, rank, b)
