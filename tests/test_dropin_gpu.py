"""Drop-in boundary on the GPU (-m gpu): the reference's own native signature, the module-level numpy export, and
``MonoRUnRoIHead.simple_test`` built from the reference's real config block."""
import numpy as np
import pytest
import torch

from monorun_b200 import synth

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_reference_native_signature_is_gpu_backed(cuda_lib, oracle):
    """``void pnp_uncert(double*...)`` of ext.h:1-13, bound the way pnp_uncert_cpu.py:70-106 binds it (cffi, numpy fp64
    buffers, one object per call) -- against the oracle's function of the same signature."""
    from monorun_b200 import _native
    b = synth.make_batch(6, config=2, weights='diag', mode='S0')
    op = synth.to_op_level(b)
    ffi = _native.ffi
    dp = lambda a: ffi.cast('double*', a.ctypes.data)
    for i in range(6):
        m = oracle.istd_inlier_masks(op['coords_2d_istd'][i:i + 1], 0.6)[0]
        p2, p3, w = (np.ascontiguousarray(op[k][i][m], np.float64) for k in ('coords_2d', 'coords_3d', 'coords_2d_istd'))
        K = np.ascontiguousarray(op['cam_mats'][0], np.float64)
        init = np.ascontiguousarray(b['init_pose'][i], np.float64)
        clips = np.array([0.5, op['u_range'][0, 0], op['u_range'][0, 1], op['v_range'][0, 0], op['v_range'][0, 1]], np.float64)
        val, pose, cov, tr = np.zeros(1, np.int32), np.zeros(4), np.eye(4), np.zeros(1)
        cuda_lib.pnp_uncert(dp(p2), dp(p3), dp(w), dp(K), dp(init), ffi.cast('int*', val.ctypes.data), dp(pose), dp(cov), dp(tr),
                            p2.shape[0], dp(clips))
        rv, rpose, rcov, rtr = oracle.lm_single(p2, p3, w, K, init, clips, with_pose_cov=True)
        assert val[0] == 1 and rv
        np.testing.assert_allclose(pose, rpose, rtol=1e-8, atol=1e-9)
        assert np.linalg.norm(cov - rcov) / np.linalg.norm(rcov) < 1e-3
        assert tr[0] == pytest.approx(rtr, rel=1e-6)
    # pn out of range: no crash, result_val stays 0 and the pose is the start (pnp_uncert_cpu.cpp:259)
    val[:] = 7
    cuda_lib.pnp_uncert(dp(p2), dp(p3), dp(w), dp(K), dp(init), ffi.cast('int*', val.ctypes.data), dp(pose), ffi.NULL, dp(tr), 3, dp(clips))
    assert val[0] == 0 and np.array_equal(pose, init)


def test_module_level_u2d_pnp_cpu_export(cuda_lib, oracle):
    """``monorun.ops.u2d_pnp_cpu`` (pnp_uncert_cpu.py:128-209): numpy in, numpy 6-tuple out; same basin as the restated
    reference driver with its OpenCV EPnP start."""
    import monorun_b200
    n = 48
    b = synth.make_batch(n, config=2, weights='diag', mode='S1')
    op = synth.to_op_level(b)
    args = (op['coords_2d'], op['coords_2d_istd'], op['coords_3d'], op['cam_mats'], op['u_range'], op['v_range'])
    out = monorun_b200.u2d_pnp_cpu(*args, z_min=0.5, epnp_istd_thres=0.6, inlier_opt_only=True)
    ref = oracle.u2d_pnp_cpu(*args, z_min=0.5, epnp_istd_thres=0.6, inlier_opt_only=True)
    assert [type(o) for o in out] == [np.ndarray] * 6
    assert out[0].dtype == bool and out[0].shape == (n,) and out[1].shape == (n, 1) and out[2].shape == (n, 3)
    assert out[3].shape == (n, 4, 4) and out[4].shape == (n, 1) and out[5].shape == (n, 784) and out[5].dtype == bool
    assert np.array_equal(out[5], ref[5]) and out[0].all() and ref[0].all()
    t_err = np.linalg.norm(out[2] - ref[2], axis=1) / np.linalg.norm(ref[2], axis=1)
    assert np.median(t_err) < 1e-4 and t_err.max() < 2e-3, (np.median(t_err), t_err.max())
    empty = monorun_b200.u2d_pnp_cpu(*(a[:0] if a.shape[0] == n else a for a in args))
    assert empty[0].shape == (0,) and empty[5].shape == (0, 784)


def _frame(n, seed=11):
    """One synthetic frame: n detections of the seeded generator, the head's output that is consistent with them
    (teacher forcing, as tools/e2e_config4.py) and five FPN levels of features."""
    b = synth.make_batch(n, config=3, mode='S1', rng=np.random.default_rng(seed))
    raw = synth.to_head_raw(b)
    labels = dev(b['labels']).long()
    C = 3
    forced = torch.zeros(n, 2, 5 * C, 28, 28, device='cuda')   # both flip halves (fcn_noc_decoder.py:225-235)
    idx = torch.arange(n, device='cuda')
    for c in range(3):
        forced[idx, 0, 3 * labels + c] = dev(raw['noc_pred'])[:, c]
    for c in range(2):
        forced[idx, 0, 3 * C + 2 * labels + c] = dev(raw['proj_logstd'])[:, c]
    feats = [torch.randn(1, 256, 192 >> i, 624 >> i, device='cuda') for i in range(5)]   # strides 2..32 of a 384 x 1248 image
    det = torch.cat([dev(b['boxes']).float(), torch.rand(n, 1, device='cuda') * 0.5 + 0.5], 1)
    return b, raw, labels, forced, feats, det


@pytest.mark.parametrize('cfg_name', ['kitti_multiclass.py', 'kitti_multiclass_lidar_supv.py'])
def test_simple_test_from_the_reference_config_block(cuda_lib, cfg_name):
    """MonoRUnRoIHead.simple_test (monorun_roi_head.py:442-605) on the head built from the reference's REAL roi_head /
    test_cfg.rcnn blocks (committed fixture): the reference's result structure, 2-D and 3-D lists filtered alike by the
    3-D NMS, and -- with the head's output teacher-forced to consistent correspondences -- the generating poses back."""
    from monorun_b200.config import build_roi_head_from_fixture
    from monorun_b200 import pnp
    torch.manual_seed(0)
    head = build_roi_head_from_fixture(cfg_name).cuda().eval()
    head.init_weights()
    assert head.noc_roi_extractor.featmap_strides == [2, 4, 8, 16, 32] and head.noc_roi_extractor.finest_scale == 28
    assert head.test_cfg.nms_3d_thr == 0.01 and head.pose_head.epnp_ransac_thres_ratio == 0.2
    n = 40
    b, raw, labels, forced, feats, det = _frame(n)
    head.set_bbox_stage(lambda x, proposals, metas, rescale, cfg: (det, labels))
    metas = [dict(img_shape=(384, 1248, 3), scale_factor=1.0, flip=False)]
    cam = [[dev(b['cam_mat']).float()]]

    # (A) the network as initialised: structure only
    res = head.simple_test(feats, [det], metas, cam_intrinsic=cam, coord_2d=None, rescale=False)
    assert isinstance(res, list) and len(res) == 1 and set(res[0]) == {'bbox_results', 'bbox_3d_results'}
    r2, r3 = res[0]['bbox_results'], res[0]['bbox_3d_results']
    assert len(r2) == len(r3) == 3
    for a2, a3 in zip(r2, r3):
        assert isinstance(a2, np.ndarray) and isinstance(a3, np.ndarray)
        assert a2.shape[1] == 5 and a3.shape[1] == 8 and a2.shape[0] == a3.shape[0]
        assert np.isfinite(a3).all() and ((a3[:, 7] >= 0) & (a3[:, 7] <= 1)).all()

    # (B) teacher-forced head output and dimensions: the poses of the generator come back through the whole method
    head.noc_head.forward_all = lambda x, latent, flip=False, native=False: head.noc_head._select_flip_half(
        forced.view(n, 2 * 15, 28, 28), flip)
    dims = dev(raw['dims'])
    reg_forward = head.reg_forward
    coder = head.global_head.dim_coder
    def forced_reg(reg_feats, det_labels, decode_dims=True):
        out = reg_forward(reg_feats, det_labels, decode_dims=decode_dims)
        means, stds = dims.new_tensor(coder.target_means)[det_labels], dims.new_tensor(coder.target_stds)[det_labels]
        out['dim_pred'], out['dim_var'] = (dims - means) / stds, dev(raw['dims_var']) / stds.square()   # encoded
        if decode_dims:
            out['dimensions_pred'], out['dimensions_var'] = dims, dev(raw['dims_var'])
        return out
    head.reg_forward = forced_reg
    launches = pnp.launch_count()
    res = head.simple_test(feats, [det], metas, cam_intrinsic=cam, coord_2d=None, rescale=False)
    assert pnp.launch_count() - launches == 3    # fused decode + PnP, score stage, 3-D NMS
    gt = b['gt_pose']
    kept = 0
    for c in range(3):
        a2, a3 = res[0]['bbox_results'][c], res[0]['bbox_3d_results'][c]
        assert a2.shape[0] == a3.shape[0]
        cls = np.nonzero(b['labels'] == c)[0]
        assert (np.diff(a3[:, 7]) <= 1e-7).all()                       # descending score, like mmdet3d's nms_gpu
        for row2, row3 in zip(a2, a3):
            j = cls[np.argmin(np.abs(b['boxes'][cls] - row2[None, :4]).sum(1))]   # which detection this row is
            assert np.abs(b['boxes'][j] - row2[:4]).max() < 1e-3
            assert np.linalg.norm(row3[3:6] - gt[j, 1:]) / np.linalg.norm(gt[j, 1:]) < 6e-2   # the generator adds pixel noise of several px
            assert np.allclose(row3[:3], raw['dims'][j], rtol=1e-5)   # encode -> in-kernel decode round trip
        kept += a3.shape[0]
    assert 0 < kept <= n
