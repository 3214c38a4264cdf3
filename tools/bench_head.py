"""Dense correspondence head (FCNNOCDecoder up to slice_pred): libmonorun_head (tcgen05 implicit GEMM + CARAFE kernel)
against the fp32 torch modules (cuDNN, TF32 off and on) on the same random weights.  Prints one JSON line.

    python tools/bench_head.py [--rois 1024] [--steps 20]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import monorun_b200
from monorun_b200 import dense_head as dh
from tests.test_host import _roi_head_cfg

ap = argparse.ArgumentParser()
ap.add_argument('--rois', type=int, default=1024)
ap.add_argument('--steps', type=int, default=20)
a = ap.parse_args()
torch.manual_seed(0)
head = monorun_b200.build_head(_roi_head_cfg()).cuda().eval()
dec = head.noc_head
for p in dec.parameters():
    torch.nn.init.normal_(p, std=0.05) if p.dim() > 1 else torch.nn.init.normal_(p, std=0.1)
n = a.rois
x = torch.randn(n, 256, 14, 14, device='cuda').relu()
latent = torch.randn(n, 16, device='cuda')
runner = dh.DenseHeadB200(dec)
# algorithmic flops per RoI: 3 conv3x3 256->256 @14x14, compressor 1x1 256->64 @14x14, encoder 3x3 64->100 @14x14,
# CARAFE reassembly 25 MAC x 256 ch @28x28, conv3x3 256->256 @28x28, final 1x1 256->30 @28x28
flops = 2 * (3 * 9 * 256 * 256 * 196 + 256 * 64 * 196 + 9 * 64 * 100 * 196 + 25 * 256 * 784 + 9 * 256 * 256 * 784 + 256 * 30 * 784)

def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.steps, out

with torch.no_grad():
    ms_native, out = timed(lambda: runner.forward(x, latent))
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    ms_fp32, ref = timed(lambda: dec.forward_all(x, latent, flip=False))
    torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
    ms_tf32, _ = timed(lambda: dec.forward_all(x, latent, flip=False))
    # cuDNN bf16 channels-last, CONVOLUTIONS ONLY (the bar SURVEY 2b set): the head's seven convolutions with their bias
    # + ReLU as torch ops on bf16 NHWC tensors -- no latent bias, no CARAFE softmax / reassembly, no fp32 output conversion,
    # i.e. strictly less work than the native head's 10 launches
    import torch.nn.functional as F
    bf = lambda t: t.detach().to(torch.bfloat16)
    convs = [(bf(m.conv.weight).contiguous(memory_format=torch.channels_last), bf(m.conv.bias)) for m in dec.convs]
    up = dec.upsample
    wc, bc = bf(up.channel_compressor.weight).contiguous(memory_format=torch.channels_last), bf(up.channel_compressor.bias)
    we, be = bf(up.content_encoder.weight).contiguous(memory_format=torch.channels_last), bf(up.content_encoder.bias)
    xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x28 = torch.randn(n, 256, 28, 28, device='cuda', dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w28 = bf(dec.convs_upsampled[0].conv.weight).contiguous(memory_format=torch.channels_last)
    b28 = bf(dec.convs_upsampled[0].conv.bias)
    wf = bf(dec.conv_final.weight).contiguous(memory_format=torch.channels_last)
    bfin = bf(dec.conv_final.bias)

    def cudnn_convs():
        h = xb
        for w_, b_ in convs:
            h = F.relu(F.conv2d(h, w_, b_, padding=1))
        c = F.conv2d(h, wc, bc)
        e = F.conv2d(c, we, be, padding=1)
        h2 = F.relu(F.conv2d(x28, w28, b28, padding=1))
        o = F.conv2d(h2, wf, bfin)
        return e, o
    torch.backends.cudnn.benchmark = True
    ms_cudnn, _ = timed(cudnn_convs)
got = out.view(n, 2, -1, 28, 28)[:, 0]
rel_rms = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
peak = None
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['bf16_tflops']
except Exception:
    pass
line = {'rois': n, 'gflop_per_roi': flops / 1e9,
        'native': {'ms': ms_native, 'rois_per_s': n / ms_native * 1e3, 'tflops': flops * n / ms_native / 1e9,
                   'frac_of_measured_bf16_peak': (flops * n / ms_native / 1e9) / peak if peak else None,
                   'launches_per_forward': 10, 'dtype': 'bf16 operands, fp32 accumulate (TMEM)'},
        'torch_fp32_cudnn': {'ms': ms_fp32, 'rois_per_s': n / ms_fp32 * 1e3},
        'torch_tf32_cudnn': {'ms': ms_tf32, 'rois_per_s': n / ms_tf32 * 1e3},
        'cudnn_bf16_channels_last_convs_only': {
            'ms': ms_cudnn, 'note': 'F.conv2d bf16 NHWC (cudnn.benchmark) + bias + ReLU for the 7 convolutions only: no latent bias, '
                                    'CARAFE softmax / reassembly or output conversion', 'native_whole_head_over_this': ms_native / ms_cudnn},
        'speedup_vs_fp32': ms_fp32 / ms_native, 'speedup_vs_tf32': ms_tf32 / ms_native,
        'rel_rms_vs_fp32': rel_rms, 'steps': a.steps}
print(json.dumps(line))
