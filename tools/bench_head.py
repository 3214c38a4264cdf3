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
        'speedup_vs_fp32': ms_fp32 / ms_native, 'speedup_vs_tf32': ms_tf32 / ms_native,
        'rel_rms_vs_fp32': rel_rms, 'steps': a.steps}
print(json.dumps(line))
