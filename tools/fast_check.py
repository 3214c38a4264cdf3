"""Scratch: parity + timing of MRPNP_PREC_FAST against MIXED and the oracle (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tools.first_gpu_check import parity, timing

if __name__ == '__main__':
    print(torch.cuda.get_device_name(0))
    precs = sys.argv[1].split(',') if len(sys.argv) > 1 else ['mixed', 'fast']
    if 'noparity' not in sys.argv:
        for prec in precs:
            parity(2048, 2, 'diag', 'S0', prec)
            parity(2048, 2, 'diag', 'S1', prec)
            parity(2048, 3, 'full', 'S0', prec)
            parity(2048, 3, 'full', 'S1', prec)
    for prec in precs:
        timing(8192, 'diag', 'S1', prec)
        timing(32768, 'diag', 'S1', prec)
        timing(8192, 'full', 'S1', prec, config=3)
