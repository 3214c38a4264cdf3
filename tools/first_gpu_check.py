"""Scratch: first parity + timing run on the GPU box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monorun_b200 import synth, pnp
from oracle import pnp_driver as od

def parity(n=256, config=2, weights='diag', mode='S0', precision='fp64'):
    b = synth.make_batch(n, config=config, weights=weights, mode=mode)
    op = synth.to_op_level(b)
    full = weights == 'full'
    w = op['w_full'] if full else op['coords_2d_istd']
    mask = od.istd_inlier_masks(w[..., [0, 2]] if full else w, 0.6)
    cnt = mask.sum(1); mask[cnt <= 4] = True
    clips = np.array([[0.5, op['u_range'][0,0], op['u_range'][0,1], op['v_range'][0,0], op['v_range'][0,1]]])
    ref = od.lm_batch(op['coords_2d'], op['coords_3d'], w, op['cam_mats'], b['init_pose'], clips, mask, full_w=full, threads=0)
    dev = 'cuda'
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    uvr = torch.tensor([[op['u_range'][0,0], op['u_range'][0,1], op['v_range'][0,0], op['v_range'][0,1]]], device=dev)
    res, inl, r64 = pnp.solve_batched(t(op['coords_3d']), t(op['coords_2d']), t(w), t(op['cam_mats']), uvr,
        init_pose=t(b['init_pose']), inlier_mask=t(mask), layout='interleaved', weight_mode='full' if full else 'istd',
        precision=precision, return_fp64=True)
    torch.cuda.synchronize()
    r64 = r64.cpu().numpy(); res = res.cpu().numpy()
    terr = np.linalg.norm(r64[:,1:4]-ref['pose'][:,1:],axis=1)/np.linalg.norm(ref['pose'][:,1:],axis=1)
    yerr = np.abs(r64[:,0]-ref['pose'][:,0])
    same_evals = (r64[:,6].astype(int)==ref['stats'][:,1]).mean()
    print(f'[{precision} {weights} {mode}] n={n} t_rel max {terr.max():.3e} med {np.median(terr):.3e} yaw max {yerr.max():.3e} '
          f'same #evals {same_evals:.4f} valid {res[:,20].mean():.3f} evals hist {np.bincount(r64[:,6].astype(int))} '
          f'cost rel {np.abs(r64[:,4]-ref["cost"]).max()/ref["cost"].max():.2e}')
    # own mask (istd test on device)
    res2, inl2, _ = pnp.solve_batched(t(op['coords_3d']), t(op['coords_2d']), t(w), t(op['cam_mats']), uvr,
        init_pose=t(b['init_pose']), layout='interleaved', weight_mode='full' if full else 'istd', precision=precision)
    print('   mask mismatches', int((inl2.cpu().numpy()!=mask).sum()), 'of', mask.size)
    # planar/logstd path + linear init
    if not full:
        res3, inl3, r3 = pnp.solve_batched(t(b['coords_3d']), t(b['coords_2d']), t(b['logstd']), t(op['cam_mats']), uvr,
            layout='planar', weight_mode='logstd', precision=precision, return_fp64=True)
        r3 = r3.cpu().numpy()
        terr3 = np.linalg.norm(r3[:,1:4]-ref['pose'][:,1:],axis=1)/np.linalg.norm(ref['pose'][:,1:],axis=1)
        print(f'   planar+logstd+linear-init: t_rel vs oracle(gt-perturbed init) max {terr3.max():.3e} med {np.median(terr3):.3e} valid {res3[:,20].mean():.3f} evals {np.bincount(r3[:,6].astype(int))}')

def timing(n, weights='diag', mode='S1', precision='fp64', config=2, reps=20):
    b = synth.make_batch(n, config=config, weights=weights, mode=mode)
    dev='cuda'
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    full = weights=='full'
    c3, c2 = t(b['coords_3d']), t(b['coords_2d'])
    w = t(b['w_full']) if full else t(b['logstd'])
    cam = t(b['cam_mat'][None]); ih, iw = b['img_shape']
    uvr = torch.tensor([[-200., iw+200., -200., ih+200.]], device=dev)
    init = t(b['init_pose'])
    kw = dict(layout='planar', weight_mode='full' if full else 'logstd', precision=precision, return_inlier_mask=False)
    for _ in range(3): pnp.solve_batched(c3,c2,w,cam,uvr,init_pose=init,**kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): res,_,_ = pnp.solve_batched(c3,c2,w,cam,uvr,init_pose=init,**kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    bytes_obj = (8 if full else 7)*784*4+96
    its = res[:,21].cpu().numpy()
    print(f'[time {precision} {weights} {mode}] n={n}: {ms*1e3:.1f} us  {n/ms*1e3:.3e} obj/s  {n*bytes_obj/ms/1e6:.1f} GB/s  mean LM iters {its.mean():.2f}')
    e0.record()
    for _ in range(reps): res,_,_ = pnp.solve_batched(c3,c2,w,cam,uvr,**kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    print(f'      linear-init: {ms*1e3:.1f} us  {n/ms*1e3:.3e} obj/s valid {res[:,20].mean().item():.3f}')

if __name__ == '__main__':
    print(torch.cuda.get_device_name(0))
    for prec in ('fp64', 'mixed'):
        parity(512, 2, 'diag', 'S0', prec)
        parity(512, 2, 'diag', 'S1', prec)
        parity(512, 3, 'full', 'S0', prec)
    for prec in ('fp64', 'mixed'):
        for n in (1024, 8192, 32768):
            timing(n, 'diag', 'S1', prec)
        timing(8192, 'diag', 'S0', prec)
        timing(8192, 'full', 'S0', prec, config=3)
