"""Host side of libmonorun_head.so: the dense correspondence head (FCNNOCDecoder.forward up to slice_pred,
monorun/models/roi_heads/bbox_3d_heads/dense_decoders/fcn_noc_decoder.py:189-235) on the tcgen05 tensor cores.

PyTorch is used for device memory only: weights are re-packed once (``pack_conv_weight``), the scratch buffer is a
``torch.empty`` and every layer runs in the library's own kernels.  There is no fallback inside these calls -- a
missing library or a CPU tensor raises.  The fp32 torch modules in ``heads.py`` remain the numerical reference of
the tests.

Precision: bf16 operands and inter-layer activations, fp32 accumulation (tensor-core TMEM) -- the reference runs
these convolutions in fp32/TF32 under cuDNN.  tests/test_head_gpu.py states the tolerance against fp32 torch.
"""
import torch

from . import _native

HC = _native.HEAD_CONST
_ctx_cache = {}


def _ctx(device):
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError(f'the tcgen05 dense head runs on CUDA tensors only; got {device}')
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _ctx_cache:
        out = _native.head_ffi.new('mrhead_ctx**')
        _native.head_check(_native.head_lib().mrhead_create(out, idx))
        _ctx_cache[idx] = out[0]
    return _ctx_cache[idx]


def launch_count(device='cuda'):
    return int(_native.head_lib().mrhead_launch_count(_ctx(device)))


def _p(t, ctype='void*'):
    return _native.head_ffi.cast(ctype, t.data_ptr()) if t is not None else _native.head_ffi.NULL


def _stream(dev):
    return _native.head_ffi.cast('void*', torch.cuda.current_stream(dev).cuda_stream)


def pad16(c):
    return (c + 15) // 16 * 16


def pack_conv_weight(weight):
    """nn.Conv2d weight [cout, cin, kh, kw] -> bf16 [kh*kw, cout_pad, cin] (tap = ky*kw + kx, zero rows up to
    cout_pad = ceil16(cout)): each tap is one K-major B tile of the implicit GEMM."""
    cout, cin, kh, kw = weight.shape
    cp = pad16(cout)
    w = torch.zeros((kh * kw, cp, cin), dtype=torch.bfloat16, device=weight.device)
    w[:, :cout] = weight.detach().permute(2, 3, 0, 1).reshape(kh * kw, cout, cin).to(torch.bfloat16)
    return w.contiguous()


class Layer:
    """One packed convolution (keeps the tensors alive behind the C struct)."""

    def __init__(self, conv, relu):
        kh, kw = conv.kernel_size
        assert kh == kw and kh in (1, 3) and conv.stride == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1
        assert conv.padding == ((kh - 1) // 2,) * 2
        self.weight = pack_conv_weight(conv.weight)
        self.bias = conv.bias.detach().float().contiguous() if conv.bias is not None else None
        self.cin, self.cout, self.cout_pad, self.taps, self.relu = conv.in_channels, conv.out_channels, \
            pad16(conv.out_channels), kh * kw, int(relu)

    def fill(self, c):
        c.weight, c.bias = _p(self.weight), _p(self.bias, 'float*')
        c.cin, c.cout, c.cout_pad, c.taps, c.relu = self.cin, self.cout, self.cout_pad, self.taps, self.relu

    def cstruct(self):
        c = _native.head_ffi.new('mrhead_layer*')
        self.fill(c)
        return c


def pack_input(x):
    """fp32 [n, c, h, w] -> bf16 padded-flat [n, (h+2)(w+2), c] (zero halo)."""
    n, c, h, w = x.shape
    x = x.float().contiguous()
    out = torch.empty((n, (h + 2) * (w + 2), c), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _native.head_check(_native.head_lib().mrhead_pack_input(_ctx(x.device), _p(x, 'float*'), n, c, h, w, _p(out), _stream(x.device)))
    return out


def conv(layer, act, n, h, w, row_bias=None, out_mode='bf16_rows'):
    """One implicit-GEMM convolution on a padded-flat activation [n, (h+2)(w+2), cin] bf16."""
    dev = act.device
    rows = n * (h + 2) * (w + 2)
    assert act.dtype == torch.bfloat16 and act.is_contiguous() and act.numel() == rows * layer.cin
    if out_mode == 'bf16_rows':
        out = torch.empty((n, (h + 2) * (w + 2), layer.cout), dtype=torch.bfloat16, device=dev)
        mode = HC['MRHEAD_OUT_BF16_ROWS']
    elif out_mode == 'f32_rows':
        out = torch.empty((n, (h + 2) * (w + 2), layer.cout_pad), dtype=torch.float32, device=dev)
        mode = HC['MRHEAD_OUT_F32_ROWS']
    else:
        out = torch.empty((n, layer.cout, h, w), dtype=torch.float32, device=dev)
        mode = HC['MRHEAD_OUT_F32_PLANAR']
    rb = row_bias.float().contiguous() if row_bias is not None else None
    with torch.cuda.device(dev):
        _native.head_check(_native.head_lib().mrhead_conv(_ctx(dev), layer.cstruct(), _p(act), n, h, w, _p(rb, 'float*'), mode,
                                                          _p(out), _stream(dev)))
    return out


def carafe(feat, logits, n, h, w):
    """feat bf16 padded-flat [n,(h+2)(w+2),256], logits fp32 padded-flat [n,(h+2)(w+2),ld] -> bf16 padded-flat
    [n,(2h+2)(2w+2),256]."""
    dev = feat.device
    out = torch.empty((n, (2 * h + 2) * (2 * w + 2), 256), dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        _native.head_check(_native.head_lib().mrhead_carafe(_ctx(dev), _p(feat), _p(logits, 'float*'), logits.shape[-1], n, h, w,
                                                            _p(out), _stream(dev)))
    return out


def unpad(act, n, h, w):
    """padded-flat [n, (h+2)(w+2), c] -> NCHW float32 [n, c, h, w] (test helper)."""
    c = act.shape[-1]
    return act.view(n, h + 2, w + 2, c)[:, 1:h + 1, 1:w + 1].permute(0, 3, 1, 2).float().contiguous()


class DenseHeadB200:
    """Packed weights + scratch of one FCNNOCDecoder; ``forward`` returns the unsliced ``all_pred``
    [n, conv_final.out_channels, 2h, 2w] fp32 (both flip halves; pick one with ``.view(n, 2, -1, H, W)[:, half]``)."""

    def __init__(self, decoder):
        up = decoder.upsample
        if not (up.scale_factor == 2 and up.up_kernel == 5 and up.up_group == 1 and up.channels == 256):
            raise NotImplementedError('the CARAFE kernel is built for 256 channels, k_up=5, x2, group 1 (every reference config)')
        if decoder.use_dropout2d and decoder.training:
            raise RuntimeError('DenseHeadB200 is inference only (Dropout2d must be in eval mode)')
        self.convs = [Layer(m.conv, True) for m in decoder.convs]
        self.convs_up = [Layer(m.conv, True) for m in decoder.convs_upsampled]
        self.compressor = Layer(up.channel_compressor, False)
        self.encoder = Layer(up.content_encoder, False)
        self.final = Layer(decoder.conv_final, False)
        self.latent_w = self.latent_b = None
        self.latent_activation = 0
        if decoder.use_latent_vec:
            self.latent_w = decoder.latent_decoder.weight.detach().float().contiguous()
            self.latent_b = decoder.latent_decoder.bias.detach().float().contiguous()
            act = decoder.latent_activation
            self.latent_activation = 0 if act is None else 1 if isinstance(act, torch.nn.ReLU) else 2
        w = _native.head_ffi.new('mrhead_weights*')
        for i, l in enumerate(self.convs):
            l.fill(w.convs[i])
        w.num_convs = len(self.convs)
        for i, l in enumerate(self.convs_up):
            l.fill(w.convs_up[i])
        w.num_convs_up = len(self.convs_up)
        self.compressor.fill(w.compressor)
        self.encoder.fill(w.encoder)
        self.final.fill(w.final)
        w.latent_w, w.latent_b = _p(self.latent_w, 'float*'), _p(self.latent_b, 'float*')
        w.latent_channels = self.latent_w.shape[1] if self.latent_w is not None else 0
        w.latent_activation = self.latent_activation
        self.cweights = w
        self.device = decoder.conv_final.weight.device
        self._workspace = None

    def forward(self, x, latent_pred=None):
        n, c, h, w = x.shape
        dev = x.device
        x = x.float().contiguous()
        out = torch.empty((n, self.final.cout, 2 * h, 2 * w), dtype=torch.float32, device=dev)
        if n == 0:
            return out
        lib = _native.head_lib()
        need = int(lib.mrhead_workspace_bytes(self.cweights, n, h, w))
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != dev:
            self._workspace = torch.empty(need + 1024, dtype=torch.uint8, device=dev)
        base = self._workspace.data_ptr()
        off = (-base) % 1024
        lat = latent_pred.float().contiguous() if (latent_pred is not None and self.latent_w is not None) else None
        with torch.cuda.device(dev):
            _native.head_check(lib.mrhead_forward(
                _ctx(dev), self.cweights, _p(x, 'float*'), _p(lat, 'float*'), n, h, w,
                _native.head_ffi.cast('void*', base + off), need, _p(out, 'float*'), _stream(dev)))
        return out
