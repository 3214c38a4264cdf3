"""Scratch: small solves of every kernel variant for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monorun_b200 import synth, pnp
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for weights, cfg in (('diag', 2), ('full', 3)):
    b = synth.make_batch(40, config=cfg, weights=weights, mode='S1')
    op = synth.to_op_level(b)
    full = weights == 'full'
    ih, iw = b['img_shape']
    uvr = torch.tensor([[-200., iw + 200., -200., ih + 200.]], device='cuda')
    for prec in ('fp64', 'mixed', 'fast'):
        for layout in ('planar', 'interleaved'):
            if layout == 'planar':
                args = (t(b['coords_3d']), t(b['coords_2d']), t(b['w_full'] if full else b['logstd']))
                wm = 'full' if full else 'logstd'
            else:
                args = (t(op['coords_3d']), t(op['coords_2d']), t(op['w_full'] if full else op['coords_2d_istd']))
                wm = 'full' if full else 'istd'
            for init in (t(b['init_pose']), None):
                res, inl, _ = pnp.solve_batched(*args, t(b['cam_mat'][None]), uvr, init_pose=init, layout=layout,
                                                weight_mode=wm, precision=prec)
    torch.cuda.synchronize()
    print(weights, 'valid', res[:, 20].mean().item())
# unaligned / small shapes and the clip fallback
b = synth.make_batch(9, config=2, roi=7)
op = synth.to_op_level(b)
res, _, _ = pnp.solve_batched(t(op['coords_3d']), t(op['coords_2d']), t(op['coords_2d_istd']), t(b['cam_mat'][None]),
                              torch.tensor([[550., 650., 150., 220.]], device='cuda'), init_pose=t(b['init_pose']),
                              layout='interleaved', weight_mode='istd', precision='fast')
torch.cuda.synchronize()
print('clip case valid', res[:, 20].mean().item())
