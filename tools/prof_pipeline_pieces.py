"""Scratch: split the time of MonoRUnRoIHead.forward_3d(fused=True, native_head=True)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import monorun_b200
from monorun_b200 import synth, pnp
from tests.test_host import _roi_head_cfg
torch.manual_seed(0)
head = monorun_b200.build_head(_roi_head_cfg()).cuda().eval(); head.init_weights()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = synth.make_batch(n, config=3, mode='S1')
dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
rois = torch.cat([torch.zeros(n, 1), torch.from_numpy(b['boxes'])], 1).cuda()
feats = torch.randn(n, 256, 14, 14, device='cuda'); latent = torch.randn(n, 16, device='cuda')
labels, dims = dev(b['labels']), dev(b['dims']); dims_var = torch.full((n, 3), 1e-3, device='cuda'); cam = dev(b['cam_mat'][None])
img_shapes = cam.new_tensor((375, 1242))[None]
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3, out
with torch.no_grad():
    ms_all, out = timed(lambda: head.forward_3d(feats, rois, labels, latent, dims, dims_var, cam, (375, 1242), fused=True, native_head=True))
    ms_head, all_pred = timed(lambda: head.noc_head.forward_all(feats, latent, False, native=True))
    nh = head.noc_head
    ms_pose, _ = timed(lambda: head.pose_head.forward_fused(all_pred, None, rois, dims, dims_var, cam, img_shapes, nh.coord_coder,
                       head.projection_head.proj_error_coder, labels=labels, num_classes=nh.num_classes))
    ms_solve, (res, inl) = timed(lambda: pnp.solve_dense(all_pred, None, rois, dims, dims_var, cam, head.pose_head._uv_range(img_shapes),
                       noc_mean=nh.coord_coder.target_means, noc_std=nh.coord_coder.target_stds, focal_gain=722.0, scaling_denominator=173.28,
                       labels=labels, num_classes=nh.num_classes))
    ms_solve_nomask, _ = timed(lambda: pnp.solve_dense(all_pred, None, rois, dims, dims_var, cam, head.pose_head._uv_range(img_shapes),
                       noc_mean=nh.coord_coder.target_means, noc_std=nh.coord_coder.target_stds, focal_gain=722.0, scaling_denominator=173.28,
                       labels=labels, num_classes=nh.num_classes, return_inlier_mask=False))
print(f'n={n}: forward_3d {ms_all:.3f} ms | head {ms_head:.3f} | pose_head.forward_fused {ms_pose:.3f} | solve_dense {ms_solve:.3f} | solve_dense without mask unpack {ms_solve_nomask:.3f}')
