"""CPU tests of the host-side logic: registries and config loading, coders, analytic coords_2d, mask packing,
the C-ABI library's symbols, the dense head in torch, and the world_size-2 gather on gloo."""
import glob
import os
import re

import numpy as np
import pytest
import torch

import monorun_b200
from monorun_b200 import _native, coders, heads, pnp, registry, synth
from monorun_b200.config import ConfigDict, build_roi_head, load_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CONFIGS = sorted(glob.glob('/root/reference/configs/kitti_*.py'))


# ------------------------------------------------------------------ registries / plugin surface
def test_registry_contract():
    reg = registry.Registry('x')

    @reg.register_module()
    class A:
        def __init__(self, a=1, b=2):
            self.a, self.b = a, b

    obj = registry.build_from_cfg(dict(type='A', a=5), reg, default_args=dict(b=7))
    assert (obj.a, obj.b) == (5, 7)
    with pytest.raises(KeyError):
        registry.build_from_cfg(dict(type='Nope'), reg)
    with pytest.raises(KeyError):
        reg.register_module()(A)
    with pytest.raises(KeyError):
        registry.build_from_cfg(dict(a=1), reg)


def test_drop_in_names_are_registered():
    assert 'PnPUncert' in registry.PNP
    for name in ('UncertPropPnPOptimizer', 'FCNNOCDecoder', 'MonoRUnRoIHead', 'UncertProjectionHead'):
        assert name in registry.HEADS
    assert 'NOCCoder' in registry.COORD_CODERS and 'DistanceInvarProjErrorCoder' in registry.PROJ_ERROR_CODERS
    assert 'MultiClassNormDimCoder' in registry.DIM_CODERS and 'Vec2DRotationCoder' in registry.ROTATION_CODERS


def test_pnp_uncert_constructor_matches_reference_signature():
    # configs/kitti_multiclass.py:124-129
    m = monorun_b200.build_pnp(dict(type='PnPUncert', z_min=0.5, epnp_istd_thres=0.6, inlier_opt_only=True,
                                    forward_exact_hessian=False))
    assert (m.z_min, m.epnp_istd_thres, m.inlier_opt_only, m.coord_istd_normalize, m.use_6dof, m.eps) == \
        (0.5, 0.6, True, False, False, 1e-6)
    with pytest.raises(TypeError):  # the reference's PnPUncert rejects this key too (pnp_uncert.py:93-99)
        monorun_b200.build_pnp(dict(type='PnPUncert', backward_exact_hessian=True))


def _roi_head_cfg(num_classes=3):
    """Hand-written block with the keys of configs/kitti_multiclass.py:36-144 (used when /root/reference is absent)."""
    return dict(
        type='MonoRUnRoIHead',
        bbox_roi_extractor=dict(type='SingleRoIExtractor'), bbox_head=dict(type='Shared2FCBBoxHead'),
        global_head=dict(type='FCExtractorMonteCarlo'), noc_roi_extractor=dict(type='SingleRoIExtractor'),
        noc_head=dict(type='FCNNOCDecoder', num_convs=3, roi_feat_size=14, in_channels=256, conv_kernel_size=3,
                      conv_out_channels=256, num_classes=num_classes, class_agnostic=False,
                      upsample_cfg=dict(type='carafe', scale_factor=2), num_convs_upsampled=1, loss_noc=None,
                      noc_channels=3, uncert_channels=2, dropout2d_rate=0.2, flip_correction=True,
                      coord_coder=dict(type='NOCCoder', target_means=(-0.1, -0.5, 0.0),
                                       target_stds=(0.35, 0.23, 0.34), eps=1e-5), latent_channels=16),
        projection_head=dict(type='UncertProjectionHead', loss_proj=dict(type='RobustKLLoss'),
                             proj_error_coder=dict(type='DistanceInvarProjErrorCoder', ref_length=1.6,
                                                   ref_focal_y=722, target_std=0.15)),
        pose_head=dict(type='UncertPropPnPOptimizer',
                       pnp=dict(type='PnPUncert', z_min=0.5, epnp_istd_thres=0.6, inlier_opt_only=True,
                                forward_exact_hessian=False),
                       rotation_coder=dict(type='Vec2DRotationCoder'), allowed_border=200,
                       epnp_ransac_thres_ratio=0.2),
        score_head=dict(type='MLPScoreHead'))


def test_roi_head_builds_from_config_block():
    h = monorun_b200.build_head(_roi_head_cfg())
    assert isinstance(h.pose_head, heads.UncertPropPnPOptimizer) and isinstance(h.pose_head.pnp, pnp.PnPUncert)
    assert h.noc_head.conv_final.out_channels == 30 and h.pose_head.cov_calib_logscale.shape == (4,)
    assert h.projection_head.proj_error_coder.scaling_denomitor == pytest.approx(1.6 * 722 * 0.15)
    keys = set(h.state_dict())
    for k in ('noc_head.convs.0.conv.weight', 'noc_head.latent_decoder.weight', 'noc_head.convs_upsampled.0.conv.bias',
              'noc_head.upsample.channel_compressor.weight', 'noc_head.upsample.content_encoder.weight',
              'noc_head.conv_final.weight', 'pose_head.cov_calib_logscale'):
        assert k in keys, k
    with pytest.raises(TypeError):
        monorun_b200.build_head(dict(_roi_head_cfg(), bogus_argument=1))


@pytest.mark.skipif(not REF_CONFIGS, reason='/root/reference is only present in the authoring container')
@pytest.mark.parametrize('path', REF_CONFIGS)
def test_reference_configs_still_load(path):
    cfg = load_config(path)
    h = build_roi_head(cfg)
    ncls = cfg['model']['roi_head']['noc_head']['num_classes']
    assert h.noc_head.conv_final.out_channels == 5 * ncls * 2
    assert h.pose_head.pnp.z_min == 0.5 and h.pose_head.pnp.epnp_istd_thres == 0.6
    assert h.test_cfg.cov_correction is True and h.pose_head.allowed_border == 200


def test_product_path_refuses_cpu_tensors():
    a = torch.zeros(2, 16, 2)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        monorun_b200.pnp_uncert(a, a, torch.zeros(2, 16, 3), torch.eye(3)[None], torch.zeros(1, 2), torch.zeros(1, 2))
    e = monorun_b200.pnp_uncert(a[:0], a[:0], torch.zeros(0, 16, 3), torch.eye(3)[None], torch.zeros(1, 2),
                                torch.zeros(1, 2))  # empty batch never touches the device (pnp_uncert.py:60-61)
    assert e[0].shape == (0,) and e[3].shape == (0, 4, 4) and e[4].shape == (0, 16)


def test_product_code_never_imports_the_oracle():
    for f in glob.glob(os.path.join(ROOT, 'monorun_b200', '**', '*.py'), recursive=True) + \
            glob.glob(os.path.join(ROOT, 'monorun_b200', 'csrc', '*')):
        text = open(f).read()
        assert not re.search(r'^\s*(from|import)\s+oracle', text, re.M), f
        assert 'pnp_oracle' not in text and 'pnp_driver' not in text, f


# ------------------------------------------------------------------ C ABI
def test_cabi_library_exports_every_declared_symbol():
    _native.build()
    lib = _native.lib()
    header = open(_native.HEADER).read()
    declared = set(re.findall(r'\b(mrpnp_[a-z_0-9]+)\s*\(', header)) | {'pnp_uncert'}   # + the reference's own entry (ext.h:1-13)
    assert re.search(r'^void pnp_uncert\(double\* pts2d, double\* pts3d, double\* wgt2d, double\* K, double\* init_pose, int\* result_val,',
                     header, re.M)
    assert declared == set(_native.EXPORTED), declared ^ set(_native.EXPORTED)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mrpnp_version() == _native.CONST['MRPNP_VERSION']
    p = _native.ffi.new('mrpnp_params*')
    lib.mrpnp_default_params(p, 7, 784)
    assert (p.n_obj, p.n_pts, p.inlier_opt_only, p.max_iterations) == (7, 784, 1, 50)
    assert (p.z_min, p.std_scale) == (0.5, 10.0) and p.istd_thres == pytest.approx(0.6)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the error path on a box without a GPU')
def test_cabi_create_fails_loudly_without_gpu():
    lib = _native.lib()
    out = _native.ffi.new('mrpnp_ctx**')
    rc = lib.mrpnp_create(out, 0)
    assert rc == _native.CONST['MRPNP_ERR_CUDA'] and out[0] == _native.ffi.NULL
    assert len(_native.last_error()) > 0


# ------------------------------------------------------------------ coders / analytic coords_2d / masks
def test_mask_pack_roundtrip():
    for p in (784, 196, 49, 1024, 33):
        m = torch.rand(5, p) > 0.4
        m[:, 31 % p] = True
        w = pnp.pack_mask(m)
        assert w.shape == (5, (p + 31) // 32) and w.dtype == torch.int32
        assert torch.equal(pnp.unpack_mask(w, p), m)


def test_coders_follow_reference_formulas():
    torch.manual_seed(0)
    noc, dims, dvar = torch.randn(4, 3, 28, 28), torch.rand(4, 3) + 1, torch.rand(4, 3) * 0.01
    c = coders.NOCCoder()
    c3, c3v = c.decode(noc, None, dims, dvar, False)
    pn = noc * torch.tensor([0.35, 0.23, 0.34])[:, None, None] + torch.tensor([-0.1, -0.5, 0.0])[:, None, None]
    assert torch.allclose(c3, pn * dims[..., None, None]) and torch.allclose(c3v, dvar[..., None, None] * pn ** 2)
    pc = coders.DistanceInvarProjErrorCoder(ref_length=1.6, ref_focal_y=722, target_std=0.15)
    sd = 1.6 * 722 * 0.15
    ls = torch.randn(4, 2, 28, 28) * 0.3
    out = pc.decode_logstd(ls, c3v, None)
    v2 = torch.stack([0.5 * (c3v[:, 0] + c3v[:, 2]), c3v[:, 1]], 1)
    assert torch.allclose(out, 0.5 * torch.log((v2 * 722 ** 2 + torch.exp(2 * ls) * sd ** 2) / sd ** 2), atol=1e-6)
    assert torch.allclose(pc.decode_logstd(ls, None, None), ls, atol=1e-6)
    cov = torch.eye(4)[None].repeat(3, 1, 1)
    assert torch.allclose(pc.cov_correction(cov, torch.tensor([sd, 2 * sd, 0.5 * sd]))[:, 0, 0], torch.tensor([1, .25, 4.]))
    dc = coders.MultiClassNormDimCoder()
    d, dv = dc.decode(torch.zeros(3, 3), torch.ones(3, 3), torch.tensor([0, 1, 2]))
    assert torch.allclose(d, torch.tensor(dc.target_means)) and torch.allclose(dv, torch.tensor(dc.target_stds) ** 2)


def test_analytic_coords_2d_equals_roi_align_of_pixel_grid():
    """monorun_roi_head.py:521-523: roi_align(coord_2d, rois, (28,28), 1.0, 0, 'avg', True)."""
    tv = pytest.importorskip('torchvision')
    h, w = 384, 1248
    v, u = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing='ij')
    grid = torch.stack([u, v])[None]  # (1, 2, H, W): u = column, v = row (loading.py:67-78)
    rois = torch.tensor([[0, 100.3, 50.7, 180.9, 120.2], [0, 600.0, 170.0, 640.0, 200.0],
                         [0, 3.0, 2.5, 300.0, 370.0], [0, 1000.5, 10.0, 1240.0, 375.0]])
    ref = tv.ops.roi_align(grid, rois, (28, 28), 1.0, 0, True)
    out = coders.coords_2d_from_rois(rois, 28)
    assert out.shape == (4, 2, 28, 28)
    assert (out - ref).abs().max() < 2e-3


# ------------------------------------------------------------------ dense head (torch layers)
def test_carafe_matches_naive_definition():
    torch.manual_seed(1)
    m = heads.CARAFEPack(channels=8, scale_factor=2, up_kernel=5, compressed_channels=4).double()
    x = torch.randn(2, 8, 5, 6, dtype=torch.float64)
    out = m(x)
    mask = torch.softmax(torch.nn.functional.pixel_shuffle(m.content_encoder(m.channel_compressor(x)), 2), dim=1)
    ref = torch.zeros(2, 8, 10, 12, dtype=torch.float64)
    xp = torch.nn.functional.pad(x, (2, 2, 2, 2))
    for yy in range(10):
        for xx in range(12):
            for a in range(5):
                for b in range(5):
                    ref[:, :, yy, xx] += mask[:, a * 5 + b, yy, xx][:, None] * xp[:, :, yy // 2 + a, xx // 2 + b]
    assert torch.allclose(out, ref, atol=1e-12)


def test_fcn_noc_decoder_forward_and_slicing():
    torch.manual_seed(2)
    d = monorun_b200.build_head(_roi_head_cfg()['noc_head']).eval()
    d.init_weights()
    x, lat, labels = torch.randn(5, 256, 14, 14), torch.randn(5, 16), torch.tensor([0, 1, 2, 1, 0])
    noc, noc_var, logstd, reg = d(x, lat, None, labels, flip=False)
    assert noc.shape == (5, 3, 28, 28) and logstd.shape == (5, 2, 28, 28) and noc_var is None and reg is None
    noc_f, _, logstd_f, _ = d(x, lat, None, labels, flip=[False, True, False, True, True])
    assert torch.equal(noc_f[0], noc[0]) and not torch.equal(noc_f[1], noc[1])
    e = d(x[:0], lat[:0], None, labels[:0])
    assert e[0].shape == (0, 3, 28, 28) and e[2].shape == (0, 2, 28, 28)


# ------------------------------------------------------------------ multi-GPU plumbing on gloo (world_size 2)
def _gloo_worker(rank, world, port, n_total, ret):
    import torch.distributed as dist
    from monorun_b200 import dist as mdist
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    try:
        def solve(start, stop):  # stand-in for the kernel: row i carries its global object index
            rows = torch.zeros(stop - start, pnp.RESULT_STRIDE)
            rows[:, 0] = torch.arange(start, stop)
            rows[:, 20] = rank
            return rows
        out = mdist.solve_sharded(solve, n_total)
        ok = out.shape == (n_total, pnp.RESULT_STRIDE) and torch.equal(out[:, 0], torch.arange(n_total).float())
        counts = mdist.shard_counts(n_total, world)
        ok = ok and torch.equal(out[:, 20], torch.repeat_interleave(torch.arange(world), torch.tensor(counts)).float())
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_total', [64, 37])
def test_sharded_gather_world_size_2_gloo(n_total):
    import torch.multiprocessing as mp
    from monorun_b200 import dist as mdist
    assert [mdist.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sum(mdist.shard_counts(65536, 8)) == 65536 and set(mdist.shard_counts(65536, 8)) == {8192}
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = 29500 + (os.getpid() + n_total) % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n_total, ret)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs) and ret.get(0) and ret.get(1)


def test_head_raw_generator_inverts_the_decoders():
    """synth.to_head_raw: NOCCoder.decode(noc_pred) gives back coords_3d, the RoI grid gives back coords_2d."""
    from monorun_b200 import coders
    b = synth.make_batch(16, config=3, weights='diag', mode='S1')
    raw = synth.to_head_raw(b, rng=np.random.default_rng(5))
    cc = coders.NOCCoder(synth.NOC_MEANS, synth.NOC_STDS)
    c3, c3v = cc.decode(torch.from_numpy(raw['noc_pred']), None, torch.from_numpy(raw['dims']),
                        torch.from_numpy(raw['dims_var']), False)
    assert np.abs(c3.numpy() - b['coords_3d']).max() < 1e-5 and (c3v >= 0).all()
    c2 = coders.coords_2d_from_rois(torch.from_numpy(raw['rois']))
    assert np.abs(c2.numpy() - b['coords_2d']).max() < 2e-4
    with pytest.raises(ValueError):
        synth.to_head_raw(synth.make_batch(2, config=2, mode='S0'))


def test_global_extractor_mirrors_reference_interface():
    """FCExtractorMonteCarlo (fc_extractor.py:12-156, fc_extractor_monte_carlo.py:21-82): state-dict keys, the
    5-tuple, sample mean / variance over num_samples dropout passes, per-class slicing and dimension decode."""
    g = monorun_b200.build_head(dict(type='FCExtractorMonteCarlo', num_samples=8, num_classes=3, latent_channels=16,
                                     in_channels=4, fc_out_channels=32, roi_feat_size=7,
                                     loss_dim=dict(type='SmoothL1LossMod', loss_weight=1.0, beta=1.0),
                                     dim_coder=dict(type='MultiClassNormDimCoder'))).eval()
    g.init_weights()
    assert set(g.state_dict()) == {'fcs.0.weight', 'fcs.0.bias', 'fcs.1.weight', 'fcs.1.bias', 'fc_reg.weight',
                                   'fc_reg.bias'}
    assert g.fcs[0].in_features == 4 * 49 and g.fc_reg.out_features == (3 + 16) * 3
    x, labels = torch.randn(5, 4, 7, 7), torch.tensor([0, 2, 1, 1, 0])
    torch.manual_seed(3)
    mean, var, dist, dist_logstd, feat = g(x)
    assert mean.shape == (5, 57) and var.shape == (5, 57) and feat.shape == (5, 32) and dist is None and dist_logstd is None
    torch.manual_seed(3)   # same dropout masks: the reference's repeat -> forward -> view(num_samples, n, -1) -> var_mean
    pred, f = g._fc_forward(x.repeat(8, 1, 1, 1))
    v_ref, m_ref = torch.var_mean(pred.view(8, 5, -1), dim=0)
    assert torch.equal(mean, m_ref) and torch.equal(var, v_ref) and torch.equal(feat, f.view(8, 5, -1).mean(0))
    assert (var > 0).any()   # dropout is active at test time
    dim_pred, dim_var, latent_pred, latent_var = g.slice_pred(mean, var, labels)
    assert torch.equal(dim_pred[1], mean[1, 2 * 19:2 * 19 + 3]) and torch.equal(latent_var[2], var[2, 19 + 3:2 * 19])
    dims, dims_var = g.dim_coder.decode(dim_pred, dim_var, labels)
    assert torch.allclose(dims[1], dim_pred[1] * torch.tensor([0.15, 0.10, 0.14]) + torch.tensor([1.77, 1.72, 0.57]))
    assert torch.allclose(dims_var[1], dim_var[1] * torch.tensor([0.15, 0.10, 0.14]) ** 2)
    # the deterministic parent: no variance, dropout off in eval
    d = monorun_b200.build_head(dict(type='FCExtractor', in_channels=4, fc_out_channels=32)).eval()
    a, b = d(x), d(x)
    assert torch.equal(a[0], b[0]) and a[1] is None
    h = monorun_b200.build_head(dict(_roi_head_cfg(), global_head=dict(type='FCExtractorMonteCarlo', in_channels=4,
                                                                        fc_out_channels=32)))
    out = h.eval().reg_forward(x, labels)
    assert out['dimensions_pred'].shape == (5, 3) and out['latent_pred'].shape == (5, 16) and out['reg_fc_out'].shape == (5, 32)
