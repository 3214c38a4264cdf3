"""-m gpu tests of libmonorun_head.so (tcgen05 dense correspondence head) against plain PyTorch fp32 references.

Tolerances: operands are bf16 (8 mantissa bits) with fp32 accumulation.  Layer tests feed the torch reference the
SAME bf16-rounded operands, so only the accumulation order and the bf16 rounding of the stored output differ:
|err| <= 2^-8 |ref| + 1e-3 max|ref|.  The whole-head test compares against the fp32 torch modules end to end
(six bf16 layers deep): relative RMS error < 2e-2.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dh(cuda_lib):
    from monorun_b200 import _native, dense_head
    _native.head_lib()
    return dense_head


def _bf(t):
    return t.to(torch.bfloat16).float()


def _close(out, ref, what, abs_frac=1e-3):
    err = (out - ref).abs()
    bound = ref.abs() * 2.0 ** -8 + abs_frac * ref.abs().max()
    bad = (err > bound)
    assert not bad.any(), f'{what}: {int(bad.sum())} of {bad.numel()} outside tolerance, max err {err.max().item():.4g} ' \
                          f'(ref max {ref.abs().max().item():.4g}) first at {bad.nonzero()[0].tolist()}'


def _conv_case(dh, n, h, cin, cout, k, relu, seed, row_bias=False, out_mode='bf16_rows'):
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.randn(n, cin, h, h, device='cuda', generator=g)
    conv = torch.nn.Conv2d(cin, cout, k, padding=(k - 1) // 2).cuda()
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, device='cuda', generator=g) / (cin * k * k) ** 0.5)
        conv.bias.copy_(torch.randn(cout, device='cuda', generator=g))
    rb = torch.randn(n, cout, device='cuda', generator=g) if row_bias else None
    layer = dh.Layer(conv, relu)
    act = dh.pack_input(x)
    out = dh.conv(layer, act, n, h, h, row_bias=rb, out_mode=out_mode)
    ref = F.conv2d(_bf(x).double(), _bf(conv.weight).double(), conv.bias.double(), padding=(k - 1) // 2)
    if relu:
        ref = ref.relu()
    if rb is not None:
        ref = ref + rb.double()[:, :, None, None]
    return out, ref.float(), layer


def test_pack_input_layout(dh):
    x = torch.randn(3, 256, 14, 14, device='cuda')
    act = dh.pack_input(x)
    assert act.shape == (3, 256, 256) and act.dtype == torch.bfloat16
    full = act.view(3, 16, 16, 256).float()
    assert torch.equal(full[:, 1:15, 1:15].permute(0, 3, 1, 2), _bf(x))
    halo = full.clone()
    halo[:, 1:15, 1:15] = 0
    assert halo.abs().max().item() == 0.0


@pytest.mark.parametrize('n,h,cin,cout,k,relu', [
    (4, 14, 256, 64, 1, False),      # CARAFE channel compressor
    (3, 14, 64, 64, 1, True),        # a single K chunk
    (5, 14, 256, 256, 3, True),      # trunk convolution, rows not a multiple of the 256-row tile
    (2, 28, 256, 256, 3, True),      # upsampled convolution (wp = 30: tiles straddle RoIs)
    (700, 14, 256, 256, 3, True),    # more tiles than SMs: the persistent loop and the TMEM hand-over
])
def test_conv_bf16_rows(dh, n, h, cin, cout, k, relu):
    out, ref, _ = _conv_case(dh, n, h, cin, cout, k, relu, seed=n * 7 + k, row_bias=(k == 3))
    full = out.view(n, h + 2, h + 2, cout).float()
    inner = full[:, 1:h + 1, 1:h + 1].permute(0, 3, 1, 2)
    _close(inner, ref, f'conv{k}x{k} {cin}->{cout} @{h}')
    halo = full.clone()
    halo[:, 1:h + 1, 1:h + 1] = 0
    assert halo.abs().max().item() == 0.0, 'halo rows must be written as zeros'


def test_conv_f32_rows_and_planar(dh):
    out, ref, layer = _conv_case(dh, 3, 14, 64, 100, 3, False, seed=5, out_mode='f32_rows')   # CARAFE content encoder
    assert out.shape == (3, 256, 112)
    inner = out.view(3, 16, 16, 112)[:, 1:15, 1:15, :100].permute(0, 3, 1, 2)
    assert (inner - ref).abs().max().item() < 1e-3 * ref.abs().max().item() + 1e-4
    out, ref, _ = _conv_case(dh, 3, 28, 256, 30, 1, False, seed=6, out_mode='f32_planar')     # conv_final
    assert out.shape == (3, 30, 28, 28)
    assert (out - ref).abs().max().item() < 1e-3 * ref.abs().max().item() + 1e-4


@pytest.mark.parametrize('n,h', [(3, 14), (1, 14), (80, 14), (2, 10)])
def test_carafe_matches_torch(dh, n, h):
    """14 x 14 maps go through the tensor-core kernel (banded GEMM; n = 80: more tiles than SMs, so CTAs loop over
    tiles and both accumulator sets and barrier phases wrap), other sizes through the fp32 kernel."""
    from monorun_b200 import heads
    g = torch.Generator(device='cuda').manual_seed(11)
    x = _bf(torch.randn(n, 256, h, h, device='cuda', generator=g))
    logits = torch.randn(n, 100, h, h, device='cuda', generator=g) * 2
    feat = dh.pack_input(x)
    lg = torch.zeros(n, h + 2, h + 2, 112, device='cuda')
    lg[:, 1:h + 1, 1:h + 1, :100] = logits.permute(0, 2, 3, 1)
    out = dh.carafe(feat, lg.view(n, -1, 112).contiguous(), n, h, h)
    # torch statement of pixel shuffle + softmax + reassembly (heads.CARAFEPack.forward after the two convs)
    mask = F.softmax(F.pixel_shuffle(logits, 2).view(n, 1, 25, 2 * h, 2 * h), dim=2)
    patches = F.unfold(x, 5, padding=2).view(n, 1, 256, 25, h, h)
    patches = patches.repeat_interleave(2, dim=4).repeat_interleave(2, dim=5)
    ref = (patches * mask.unsqueeze(2)).sum(dim=3).view(n, 256, 2 * h, 2 * h)
    full = out.view(n, 2 * h + 2, 2 * h + 2, 256).float()
    # the tensor-core kernel rounds the 25 softmax weights to bf16 (2^-9 relative each): a few 1e-3 of the largest value
    _close(full[:, 1:-1, 1:-1].permute(0, 3, 1, 2), ref, 'carafe', abs_frac=5e-3)
    halo = full.clone()
    halo[:, 1:-1, 1:-1] = 0
    assert halo.abs().max().item() == 0.0


def test_whole_head_matches_fp32_torch_modules(dh):
    import monorun_b200
    from tests.test_host import _roi_head_cfg
    torch.manual_seed(3)
    head = monorun_b200.build_head(_roi_head_cfg()).cuda().eval()
    dec = head.noc_head
    for p in dec.parameters():          # random weights on every layer (init_weights zeroes the latent decoder)
        torch.nn.init.normal_(p, std=0.05) if p.dim() > 1 else torch.nn.init.normal_(p, std=0.1)
    n = 37
    x = torch.randn(n, 256, 14, 14, device='cuda').relu()
    latent = torch.randn(n, 16, device='cuda')
    with torch.no_grad():
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        ref = dec.forward_all(x, latent, flip=False)          # [n, 15, 28, 28]: first flip half
        ref1 = dec.forward_all(x, latent, flip=True)
        torch.backends.cudnn.allow_tf32 = prev
    runner = dh.DenseHeadB200(dec)
    before = dh.launch_count()
    all_pred = runner.forward(x, latent)
    assert dh.launch_count() - before == 10    # latent, pack, 3 convs, compressor, encoder, carafe, conv@28, final
    assert all_pred.shape == (n, 30, 28, 28)
    for half, r in ((0, ref), (1, ref1)):
        got = all_pred.view(n, 2, 15, 28, 28)[:, half]
        rel_rms = ((got - r).pow(2).mean().sqrt() / r.pow(2).mean().sqrt()).item()
        assert rel_rms < 2e-2, rel_rms
        assert (got - r).abs().max().item() < 0.15 * r.abs().max().item()


def test_head_to_pose_pipeline_native(dh):
    """FCNNOCDecoder(native) -> fused PnP entry: RoI features to poses in 11 launches of this repo's kernels."""
    import monorun_b200
    from monorun_b200 import pnp, synth
    from tests.test_host import _roi_head_cfg
    torch.manual_seed(0)
    head = monorun_b200.build_head(_roi_head_cfg()).cuda().eval()
    head.init_weights()
    n = 16
    b = synth.make_batch(n, config=3, mode='S1')
    rois = torch.cat([torch.zeros(n, 1), torch.from_numpy(b['boxes'])], 1).cuda()
    args = (torch.randn(n, 256, 14, 14, device='cuda'), rois, torch.from_numpy(b['labels']).cuda(),
            torch.randn(n, 16, device='cuda'), torch.from_numpy(b['dims']).cuda(), torch.full((n, 3), 1e-3, device='cuda'),
            torch.from_numpy(b['cam_mat'][None]).cuda(), (375, 1242))
    h0, p0 = dh.launch_count(), pnp.launch_count()
    with torch.no_grad():
        out = head.forward_3d(*args, fused=True, native_head=True)
    # PnP: one launch (objects the fp32 path hands back are solved by the exact routine inside it)
    assert dh.launch_count() - h0 == 10 and pnp.launch_count() - p0 == 1
    assert out['t_vec_pred'].shape == (n, 3) and torch.isfinite(out['t_vec_pred']).all()
