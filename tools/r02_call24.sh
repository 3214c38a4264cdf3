#!/bin/bash
# call 24: which band term buys what: mismatches against the fp64 kernel and hand-backs over 12 seeds x 2 workloads
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1400 python tools/band_sweep.py 8192 "8e-6:4e-3:2e-6,8e-6:4e-3:1e-6,8e-6:4e-3:5e-7,8e-6:4e-3:0,8e-6:2e-3:2e-6,8e-6:1e-3:2e-6,8e-6:4e-3:2e-6:2e-5:1e-4,8e-6:4e-3:5e-7:2e-5:1e-4,4e-6:4e-3:2e-6" 0,1,2,3,4,5,6,7,16,17,18,19 > gpurun_out/r02_c24_band_sweep.txt 2>&1
grep TOTAL gpurun_out/r02_c24_band_sweep.txt
