"""Score stage: the oracle against a torch transcription of the reference lines (CPU), the CUDA
epilogue kernels against the oracle and MLPScoreHead.forward_rows against the reference's torch sequence (GPU)."""
import numpy as np
import pytest
import torch

import monorun_b200
from monorun_b200 import heads
from oracle import score_oracle as so


def _random_rows(n, seed=0):
    rng = np.random.default_rng(seed)
    rows = np.zeros((n, 24), np.float32)
    rows[:, 0] = rng.uniform(-np.pi, np.pi, n)
    rows[:, 1:4] = np.stack([rng.uniform(-20, 20, n), rng.normal(1.6, 0.2, n), rng.uniform(5, 60, n)], 1)
    a = rng.normal(size=(n, 4, 4))
    rows[:, 4:20] = (a @ a.transpose(0, 2, 1) * 1e-2 + 1e-3 * np.eye(4)).reshape(n, 16)
    rows[:, 20] = rng.random(n) > 0.1
    dims = np.abs(rng.normal([3.9, 1.5, 1.6], 0.3, (n, 3))).astype(np.float32)
    return rows, dims


def _score_head(seed=1):
    torch.manual_seed(seed)
    sh = monorun_b200.build_head(dict(type='MLPScoreHead', loss_score=dict(type='CrossEntropyLoss', use_sigmoid=True)))
    sh.init_weights()
    with torch.no_grad():
        sh.pose_norm.running_mean.normal_(0, 1.0)
        sh.pose_norm.running_var.uniform_(0.5, 2.0)
        sh.pose_norm.weight.normal_(1.0, 0.1)
        sh.pose_norm.bias.normal_(0, 0.1)
        sh.fc_out.weight.normal_(0, 0.05)
    return sh.eval()


def test_score_head_mirrors_reference_interface():
    sh = _score_head()
    keys = set(sh.state_dict())
    for k in ('pose_norm.running_mean', 'pose_norm.running_var', 'pose_norm.weight', 'pose_norm.num_batches_tracked',
              'pose_fcs.0.weight', 'fused_fcs.0.bias', 'fc_out.weight'):
        assert k in keys, k
    assert sh.pose_fcs[0].weight.shape == (1024, 17) and sh.fused_fcs[0].weight.shape == (256, 1024)
    assert sh.fc_out.weight.shape == (1, 256) and sh.pre_sigmoid
    with pytest.raises(AssertionError):
        monorun_b200.build_head(dict(type='MLPScoreHead', fusion_type='add', pose_fc_out_channels=512))


def test_oracle_matches_torch_transcription_of_the_reference_lines():
    """uncert_prop_pnp_optimizer.py:96-97, distance_invar_proj_error_coder.py:62-63, mlp_score_head.py:99-106 and
    monorun_roi_head.py:544-551 written with the reference's own torch calls."""
    rows, dims = _random_rows(257)
    sh = _score_head()
    logscale = torch.tensor([0.1, -0.2, 0.3, 0.05], dtype=torch.float64)
    r = torch.from_numpy(rows).double()
    yaw, t_vec, cov = r[:, :1], r[:, 1:4], r[:, 4:20].reshape(-1, 4, 4)
    s = torch.exp(logscale)
    cal = (s * s[:, None]) * cov
    sd = 1.6 * 722 * 0.15
    cal = cal * (sd / torch.norm(t_vec, p=2, dim=1)).square().view(-1, 1, 1)
    xi, yi = torch.tril_indices(4, 4)
    x = torch.cat([yaw, t_vec, cal[:, xi, yi], torch.from_numpy(dims).double()], dim=1)
    pn = sh.pose_norm.double()
    x = pn(x)
    norm = tuple(v.detach().double().numpy() for v in (pn.running_mean, pn.running_var, pn.weight, pn.bias)) + (pn.eps,)
    feat, cal_o = so.pose_features_ref(rows, dims, logscale.numpy(), sd, False, True, norm)
    np.testing.assert_allclose(feat, x.detach().numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(cal_o, cal.numpy(), rtol=1e-12)
    logits = torch.linspace(-4, 4, rows.shape[0], dtype=torch.float64)
    det = torch.rand(rows.shape[0], dtype=torch.float64)
    sc = logits.sigmoid()
    sc[~(r[:, 20] > 0.5)] = 0
    sc = det * sc
    s_o, bbox = so.finish_scores_ref(logits.numpy(), rows, dims, det.numpy(), True)
    np.testing.assert_allclose(s_o, sc.numpy(), rtol=1e-12)
    np.testing.assert_allclose(bbox, torch.cat((torch.from_numpy(dims).double(), t_vec, yaw, sc.unsqueeze(1)), dim=1).numpy(),
                               rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize('use_calib,z_depth,sd', [(True, False, 173.28), (False, False, 173.28), (True, True, 173.28),
                                                  (True, False, 0.0)])
def test_pose_feature_kernel_matches_oracle(cuda_lib, use_calib, z_depth, sd):
    from monorun_b200 import pnp
    rows, dims = _random_rows(1000, seed=3)
    sh = _score_head().cuda()
    logscale = torch.tensor([0.1, -0.2, 0.3, 0.05], device='cuda')
    pn = sh.pose_norm
    norm = tuple(v.detach().double().cpu().numpy() for v in (pn.running_mean, pn.running_var, pn.weight, pn.bias)) + (pn.eps,)
    feat, cal = pnp.pose_features(torch.from_numpy(rows).cuda(), torch.from_numpy(dims).cuda(), logscale, sd, z_depth,
                                  use_calib, pn)
    f_ref, c_ref = so.pose_features_ref(rows, dims, logscale.cpu().numpy(), sd, z_depth, use_calib, norm)
    np.testing.assert_allclose(feat.cpu().numpy(), f_ref, rtol=2e-5, atol=2e-5)     # fp32 arithmetic on the device
    np.testing.assert_allclose(cal.cpu().numpy().reshape(-1, 4, 4), c_ref, rtol=1e-5, atol=1e-9)
    # no pose_norm, no calibration
    feat2, _ = pnp.pose_features(torch.from_numpy(rows).cuda(), torch.from_numpy(dims).cuda())
    f2, _ = so.pose_features_ref(rows, dims)
    np.testing.assert_allclose(feat2.cpu().numpy(), f2, rtol=1e-6, atol=1e-7)
    logits = torch.linspace(-6, 6, rows.shape[0], device='cuda')
    det = torch.rand(rows.shape[0], device='cuda')
    sc, bbox = pnp.finish_scores(logits, torch.from_numpy(rows).cuda(), torch.from_numpy(dims).cuda(), det, True)
    s_ref, b_ref = so.finish_scores_ref(logits.cpu().numpy(), rows, dims, det.cpu().numpy(), True)
    np.testing.assert_allclose(sc.cpu().numpy(), s_ref, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(bbox.cpu().numpy(), b_ref, rtol=1e-5, atol=1e-7)
    assert (sc.cpu().numpy()[rows[:, 20] < 0.5] == 0).all()
    e = torch.zeros((0, 24), device='cuda')
    assert pnp.pose_features(e, e[:, :3])[0].shape == (0, 17) and pnp.finish_scores(e[:, 0], e, e[:, :3])[1].shape == (0, 8)


@pytest.mark.gpu
def test_score_head_rows_path_matches_reference_sequence(cuda_lib):
    """MLPScoreHead.forward_rows (ONE launch: mrpnp_score_stage; or 2 launches + library GEMMs for other network
    shapes) == the reference's torch sequence (monorun_roi_head.py:530-551) on the same result rows; and through
    MonoRUnRoIHead.forward_scores."""
    from tests.test_host import _roi_head_cfg
    from monorun_b200 import pnp
    torch.manual_seed(0)
    head = monorun_b200.build_head(_roi_head_cfg()).cuda().eval()
    head.init_weights()
    with torch.no_grad():
        head.pose_head.cov_calib_logscale.copy_(torch.tensor([0.1, -0.2, 0.3, 0.05]))
        head.score_head.pose_norm.running_mean.normal_(0, 1.0)
        head.score_head.pose_norm.running_var.uniform_(0.5, 2.0)
        head.score_head.fc_out.weight.normal_(0, 0.05)
    n = 353   # 44 tiles of 8 objects + 1; above MLPScoreHead.native_mlp_max_objects the GEMM path is taken
    rows_np, dims_np = _random_rows(n, seed=5)
    rows, dims = torch.from_numpy(rows_np).cuda(), torch.from_numpy(dims_np).cuda()
    reg = torch.randn(n, 1024, device='cuda')
    det = torch.rand(n, device='cuda')
    before = pnp.launch_count()
    with torch.no_grad():
        scores, bbox, cal = head.forward_scores(rows, reg, dims, det_scores=det, cov_correction=True, calib_scoring=True)
        assert pnp.launch_count() == before + 1
        sh, ph_ = head.score_head, head.projection_head
        s2, b2, c2 = sh.forward_rows(reg, rows, dims, cov_calib_logscale=head.pose_head.cov_calib_logscale.detach(),
                                     cov_correction_sd=ph_.proj_error_coder.scaling_denomitor,
                                     distance_z_depth=ph_.distance_mode == 'z-depth', calib_scoring=True, det_scores=det,
                                     native_mlp=False)
        assert pnp.launch_count() == before + 3
        assert torch.allclose(scores, s2, rtol=1e-3, atol=1e-5) and torch.equal(cal.reshape(-1, 16), c2.reshape(-1, 16))
        # a parameter update (optimizer step, load_state_dict) invalidates the transposed-weight cache
        w_before = sh.fused_fcs[0].weight.clone()
        sh.fused_fcs[0].weight.mul_(0.5)
        s3 = head.forward_scores(rows, reg, dims, det_scores=det, cov_correction=True, calib_scoring=True)[0]
        assert not torch.allclose(s3, scores, rtol=1e-3, atol=1e-5)
        sh.fused_fcs[0].weight.copy_(w_before)
        s4 = head.forward_scores(rows, reg, dims, det_scores=det, cov_correction=True, calib_scoring=True)[0]
        assert torch.equal(s4, scores)
        # the reference's sequence with torch ops
        yaw, t_vec, cov = rows[:, :1], rows[:, 1:4], rows[:, 4:20].reshape(-1, 4, 4)
        s = torch.exp(head.pose_head.cov_calib_logscale)
        cal_ref = (s * s[:, None]) * cov
        cal_ref = head.projection_head.proj_error_coder.cov_correction(cal_ref, head.projection_head.get_distance(t_vec))
        ref = head.score_head(reg, yaw, t_vec, cal_ref, dims).sigmoid()
        ref[~(rows[:, 20] > 0.5)] = 0
        ref = det * ref
    assert torch.allclose(cal, cal_ref, rtol=1e-5, atol=1e-9)
    assert torch.allclose(scores, ref, rtol=1e-3, atol=1e-5)
    assert torch.equal(bbox[:, :3], dims) and torch.equal(bbox[:, 3:6], t_vec) and torch.equal(bbox[:, 6], rows[:, 0])
    assert torch.equal(bbox[:, 7], scores)


def _random_boxes(n, n_groups, seed):
    rng = np.random.default_rng(seed)
    b = np.zeros((n, 8), np.float32)
    b[:, 0:3] = np.abs(rng.normal([3.9, 1.5, 1.6], [0.4, 0.1, 0.1], (n, 3)))
    b[:, 3] = rng.uniform(-12, 12, n); b[:, 4] = 1.6; b[:, 5] = rng.uniform(5, 30, n)   # crowded: many overlaps
    b[:, 6] = rng.uniform(-np.pi, np.pi, n)
    b[:, 7] = rng.random(n)
    b[::17, 7] = b[1::17, 7][:len(b[::17])]                                             # some exactly equal scores
    labels = rng.integers(0, 3, n)
    cuts = np.sort(rng.choice(np.arange(1, n), n_groups - 1, replace=False)) if n_groups > 1 else np.array([], int)
    offsets = [0] + cuts.tolist() + [n]
    return b, labels, offsets


def test_bev_iou_oracle_known_answers():
    a = np.array([4.0, 1.5, 2.0, 0.0, 0, 10.0, 0.0, 1.0])
    assert so.bev_iou_ref(a, a) == pytest.approx(1.0)
    b = a.copy(); b[3] = 2.0                     # shifted by half the length along x
    assert so.bev_iou_ref(a, b) == pytest.approx(4.0 / 12.0)
    c = a.copy(); c[6] = np.pi / 2               # same centre, rotated by 90 degrees: 2 x 2 square in common
    assert so.bev_iou_ref(a, c) == pytest.approx(4.0 / 12.0)
    d = a.copy(); d[6] = np.pi                   # a rectangle is symmetric under 180 degrees
    assert so.bev_iou_ref(a, d) == pytest.approx(1.0)
    e = a.copy(); e[3] = 10.0
    assert so.bev_iou_ref(a, e) == 0.0
    f = np.array([2.0, 1.5, 2.0, 0.0, 0, 10.0, np.pi / 4, 1.0])    # square rotated by 45 degrees vs the same square
    g = f.copy(); g[6] = 0.0
    assert so.bev_iou_ref(f, g) == pytest.approx((8 * (np.sqrt(2) - 1)) / (8 - 8 * (np.sqrt(2) - 1)), rel=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize('n,n_groups,thr', [(400, 1, 0.25), (1500, 7, 0.25), (300, 40, 0.1), (64, 64, 0.25), (700, 1, 0.5)])
def test_bev_nms_kernel_matches_oracle(cuda_lib, n, n_groups, thr):
    from monorun_b200 import pnp
    b, labels, offsets = _random_boxes(n, n_groups, seed=n + n_groups)
    keep_ref, margin = so.nms_bev_ref(b, labels, offsets, thr)
    keep = pnp.nms_bev(torch.from_numpy(b).cuda(), torch.from_numpy(labels).cuda(), offsets, thr).cpu().numpy()
    assert margin > 1e-5, 'regenerate: an IoU sits on the threshold'     # fp32 on the device vs fp64 in the oracle
    assert np.array_equal(keep, keep_ref), (keep != keep_ref).sum()
    assert 0 < keep.sum() < n or n_groups == n
    # one class, one image, defaults
    k1 = pnp.nms_bev(torch.from_numpy(b).cuda()).cpu().numpy()
    r1, m1 = so.nms_bev_ref(b)
    assert m1 > 1e-5 and np.array_equal(k1, r1)
    assert pnp.nms_bev(torch.zeros((0, 8), device='cuda')).shape == (0,)


@pytest.mark.gpu
def test_cuda_graph_replay_of_the_native_sequence(cuda_lib):
    """RoI features -> dense head (tcgen05) -> fused PnP -> score stage -> 3-D NMS captured into one CUDA graph
    (monorun_b200.graph.GraphedSequence): no entry point synchronises or allocates outside torch's allocator, and the
    replay returns bitwise what the eager sequence returns -- also on fresh inputs copied into the static buffers."""
    from tests.test_host import _roi_head_cfg
    from monorun_b200 import synth
    from monorun_b200.graph import GraphedSequence
    torch.manual_seed(0)
    head = monorun_b200.build_head(_roi_head_cfg()).cuda().eval()
    head.init_weights()
    n = 48
    nh, ph = head.noc_head, head.pose_head
    C = nh.num_classes
    cam = None

    def inputs(seed):
        b = synth.make_batch(n, config=3, mode='S1', rng=np.random.default_rng(seed))
        raw = synth.to_head_raw(b, rng=np.random.default_rng(seed + 1))
        d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
        lab = d(b['labels']).long()
        all_pred = torch.zeros(n, 5 * C, 28, 28, device='cuda')
        idx = torch.arange(n, device='cuda')
        for c in range(3):
            all_pred[idx, 3 * lab + c] = d(raw['noc_pred'])[:, c]
        for c in range(2):
            all_pred[idx, 3 * C + 2 * lab + c] = d(raw['proj_logstd'])[:, c]
        g = torch.Generator(device='cuda').manual_seed(seed)
        return (torch.randn(n, 256, 14, 14, device='cuda', generator=g), torch.randn(n, 16, device='cuda', generator=g),
                all_pred, d(raw['rois']), d(raw['dims']), d(raw['dims_var']), torch.randn(n, 1024, device='cuda', generator=g),
                torch.rand(n, device='cuda', generator=g), lab), d(b['cam_mat'][None])

    ins, cam = inputs(3)
    img_shapes = cam.new_tensor((375, 1242))[None]
    off = torch.tensor([0, 20, n], dtype=torch.int32, device='cuda')

    def sequence(feats, latent, all_pred, rois, dims, dims_var, reg, det, lab):
        head_out = nh.forward_all(feats, latent, False, native=True)
        ret_val, yaw, t_vec, cov, _ = ph.forward_fused(all_pred, None, rois, dims, dims_var, cam, img_shapes, nh.coord_coder,
                                                       head.projection_head.proj_error_coder, labels=lab, num_classes=C)
        rows = torch.cat([yaw, t_vec, cov.reshape(n, 16), ret_val.float()[:, None], torch.zeros(n, 3, device='cuda')], 1)
        scores, bbox, _ = head.forward_scores(rows, reg, dims, det_scores=det, cov_correction=True, calib_scoring=True)
        keep = head.nms_3d(bbox, lab, off, max_group=28)
        return head_out, bbox, keep

    graphed = GraphedSequence(sequence, ins)
    for seed in (3, 11):
        ins, _ = inputs(seed)
        with torch.no_grad():
            eager = [t.clone() for t in sequence(*ins)]
        out = graphed(*ins)
        torch.cuda.synchronize()
        for a, b in zip(out, eager):
            assert torch.equal(a, b)
        assert (eager[1][:, 7] >= 0).all() and eager[2].any()
