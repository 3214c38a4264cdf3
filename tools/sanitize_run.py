"""Scratch: small solves of every kernel variant for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monorun_b200 import synth, pnp
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
if os.environ.get('MRSAN_ONLY') != '6dof':
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
    # round 2: the redo phase (whole CTAs solving handed-back objects, evaluations split over the warps): every object handed
    # back, few and many objects per CTA; the consensus prune; the score stage; the flag kernels
    for n_obj in (5, 300):
        b = synth.make_batch(n_obj, config=3, weights='full', mode='S1')
        ih, iw = b['img_shape']
        uvr = torch.tensor([[-200., iw + 200., -200., ih + 200.]], device='cuda')
        log = torch.zeros(n_obj, dtype=torch.int32, device='cuda')
        res, _, _ = pnp.solve_batched(t(b['coords_3d']), t(b['coords_2d']), t(b['w_full']), t(b['cam_mat'][None]), uvr, init_pose=t(b['init_pose']),
                                      layout='planar', weight_mode='full', precision='fast', decision_bands=(0.0, 1e9, 0.0), hand_back_log=log)
        torch.cuda.synchronize()
        print('redo phase', n_obj, 'handed', int((log != 0).sum()), 'valid', res[:, 20].mean().item())
    b = synth.make_batch(40, config=2, weights='diag', mode='S1')
    ih, iw = b['img_shape']
    uvr = torch.tensor([[-200., iw + 200., -200., ih + 200.]], device='cuda')
    thr = torch.full((40,), 6.0, device='cuda')
    res, _, _ = pnp.solve_batched(t(b['coords_3d']), t(b['coords_2d']), t(b['logstd']), t(b['cam_mat'][None]), uvr, layout='planar',
                                  weight_mode='logstd', precision='fast', ransac_thres=thr)
    rows = res.clone()
    w1, b1 = torch.randn(1024, 17, device='cuda') * 0.1, torch.randn(1024, device='cuda') * 0.1
    w2t, b2 = torch.randn(1024, 256, device='cuda') * 0.05, torch.randn(256, device='cuda') * 0.1
    w3, b3 = torch.randn(256, device='cuda') * 0.1, torch.zeros(1, device='cuda')
    sc = pnp.score_stage(rows, torch.rand(40, 3, device='cuda') + 1, torch.randn(40, 1024, device='cuda'), w1, b1, w2t, b2, w3, b3,
                         det_scores=torch.rand(40, device='cuda'))
    torch.cuda.synchronize()
    print('consensus + score stage', res[:, 20].mean().item(), float(sc[0].mean()))
# the 6-DoF solve, both kernels: masks, every weight mode, both layouts (MRSAN_ONLY=6dof runs this part alone)
from tests.sixdof_cases import make_case as make6
c = make6(24, full=False, tilt=0.05)
rng = np.random.default_rng(4)
mask = t(rng.uniform(size=c['c3'].shape[:2]) < rng.uniform(0.3, 1.0, (24, 1)))
cf = make6(24, full=True)
for prec in ('fp64', 'mixed'):
    for cc, wm, m in ((c, 'istd', mask), (c, 'istd', None), (cf, 'full', None), (cf, 'full', mask)):
        r6 = pnp.solve_6dof_batched(t(cc['c3']), t(cc['c2']), t(cc['w']), t(cc['cam']), t(cc['uv_range']), t(cc['init']), m,
                                    layout='interleaved', weight_mode=wm, precision=prec)
    r6 = pnp.solve_6dof_batched(t(c['c3']).permute(0, 2, 1).contiguous(), t(c['c2']).permute(0, 2, 1).contiguous(),
                                (-torch.log(t(c['w']) * 10.0)).permute(0, 2, 1).contiguous(), t(c['cam']), t(c['uv_range']),
                                t(c['init']), mask, layout='planar', weight_mode='logstd', precision=prec)
    torch.cuda.synchronize()
    print('6dof', prec, 'valid', float((r6[:, 42] > 0).double().mean()))
