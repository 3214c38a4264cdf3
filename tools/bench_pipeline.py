"""RoI features -> poses -> scores -> 3-D NMS through the native path only (libmonorun_head + libmonorun_pnp):
where the time of the hot sequence of MonoRUnRoIHead.simple_test (monorun_roi_head.py:509-565) goes on B200.

    python tools/bench_pipeline.py [--rois 1024] [--images 16] [--steps 20]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import monorun_b200
from monorun_b200 import synth, pnp, dense_head as dh
from tests.test_host import _roi_head_cfg

ap = argparse.ArgumentParser()
ap.add_argument('--rois', type=int, default=1024)
ap.add_argument('--images', type=int, default=16)
ap.add_argument('--steps', type=int, default=20)
a = ap.parse_args()
torch.manual_seed(0)
head = monorun_b200.build_head(_roi_head_cfg()).cuda().eval()
head.init_weights()
n = a.rois
b = synth.make_batch(n, config=3, mode='S1')
dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
rois = torch.cat([torch.zeros(n, 1), torch.from_numpy(b['boxes'])], 1).cuda()
feats = torch.randn(n, 256, 14, 14, device='cuda')
latent = torch.randn(n, 16, device='cuda')
labels, dims = dev(b['labels']), dev(b['dims'])
dims_var = torch.full((n, 3), 1e-3, device='cuda')
cam = dev(b['cam_mat'][None])
reg = torch.randn(n, 1024, device='cuda')
det = torch.rand(n, device='cuda')
offsets = np.linspace(0, n, a.images + 1).astype(np.int32).tolist()

# The dense head runs on random-init weights (its output is then meaningless as a correspondence map: a solver fed
# with it sees degenerate problems), so the PnP stage is timed on a consistent synthetic head output of the same layout:
# all_pred [N, 5*C, 28, 28] = NOC maps of C classes, then log-std maps of C classes (fcn_noc_decoder.py:242-267).
raw = synth.to_head_raw(b, rng=np.random.default_rng(5))
C = head.noc_head.num_classes
all_pred = torch.zeros(n, 5 * C, 28, 28, device='cuda')
lab = labels.long()
idx = torch.arange(n, device='cuda')
for c in range(3):
    all_pred[idx, 3 * lab + c] = dev(raw['noc_pred'])[:, c]
for c in range(2):
    all_pred[idx, 3 * C + 2 * lab + c] = dev(raw['proj_logstd'])[:, c]
rois4 = dev(raw['rois'])   # (0, x1, y1, x2, y2)
dimsr, dims_varr = dev(raw['dims']), dev(raw['dims_var'])
img_shapes = cam.new_tensor((375, 1242))[None]
nh, ph = head.noc_head, head.pose_head

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def run():
    t = [ev()]
    with torch.no_grad():
        head_out = nh.forward_all(feats, latent, False, native=True)          # 10 launches (tcgen05 convs + CARAFE)
        t.append(ev())
        ret_val, yaw, t_vec, cov, _ = ph.forward_fused(all_pred, None, rois4, dimsr, dims_varr, cam, img_shapes, nh.coord_coder,
                                                       head.projection_head.proj_error_coder, labels=lab, num_classes=C)
        t.append(ev())
        rows = torch.cat([yaw, t_vec, cov.reshape(n, 16), ret_val.float()[:, None], torch.zeros(n, 3, device='cuda')], 1)
        scores, bbox, cal = head.forward_scores(rows, reg, dimsr, det_scores=det, cov_correction=True, calib_scoring=True)
        t.append(ev())
        keep = head.nms_3d(bbox, lab, offsets)
        t.append(ev())
    return t, keep, ret_val

for _ in range(3): run()
torch.cuda.synchronize()
acc = np.zeros(4)
for _ in range(a.steps):
    t, keep, ret_val = run()
    torch.cuda.synchronize()
    acc += [t[i].elapsed_time(t[i + 1]) for i in range(4)]
acc /= a.steps
print(json.dumps({'rois': n, 'images': a.images,
                  'ms': {'dense head (10 launches, random-init weights)': acc[0],
                         'fused head->PnP solve incl. on-device initialiser (1 launch, synthetic head output)': acc[1],
                         'score stage (1024 RoIs: 2 launches + 3 library GEMMs; <= 384 RoIs: 1 launch)': acc[2], '3-D NMS (1 launch)': acc[3], 'total': acc.sum()},
                  'rois_per_s': n / acc.sum() * 1e3, 'valid_poses': float(ret_val.float().mean().item()),
                  'kept_after_nms': int(keep.sum().item())}))

# ---- the same sequence captured into ONE CUDA graph (launch-bound regime: one image's worth of RoIs) ----
from monorun_b200.graph import GraphedSequence
off_t = torch.tensor(offsets, dtype=torch.int32, device='cuda')
max_group = max(b - a_ for a_, b in zip(offsets[:-1], offsets[1:]))

def sequence(feats_, latent_, all_pred_, rois_, dims_, dims_var_, reg_, det_, lab_):
    head_out = nh.forward_all(feats_, latent_, False, native=True)
    ret_val, yaw, t_vec, cov, _ = ph.forward_fused(all_pred_, None, rois_, dims_, dims_var_, cam, img_shapes, nh.coord_coder,
                                                   head.projection_head.proj_error_coder, labels=lab_, num_classes=C)
    rows = torch.cat([yaw, t_vec, cov.reshape(n, 16), ret_val.float()[:, None], torch.zeros(n, 3, device='cuda')], 1)
    scores, bbox, cal = head.forward_scores(rows, reg_, dims_, det_scores=det_, cov_correction=True, calib_scoring=True)
    keep = head.nms_3d(bbox, lab_, off_t, max_group=max_group)
    return head_out, bbox, keep

ins = (feats, latent, all_pred, rois4, dimsr, dims_varr, reg, det, lab)
try:
    g = GraphedSequence(sequence, ins)
    for _ in range(3): g(*ins)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps): out_g = g(*ins)
    e1.record(); torch.cuda.synchronize()
    ms_graph = e0.elapsed_time(e1) / a.steps
    same = bool(torch.equal(out_g[2], keep))
    print(json.dumps({'rois': n, 'cuda_graph_ms': ms_graph, 'eager_ms': float(acc.sum()), 'rois_per_s_graph': n / ms_graph * 1e3,
                      'keep_mask_equal_to_eager': same}))
except Exception as exc:
    print(json.dumps({'cuda_graph': 'failed', 'error': f'{type(exc).__name__}: {exc}'[:300]}))
