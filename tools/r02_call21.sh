#!/bin/bash
# call 21: decision-band variants (adaptive relative band, narrower mix band): hand-backs, mismatches, kernel time
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python tools/band_sweep.py 8192 "0:0:0,8e-6:4e-3:2e-6,8e-6:4e-3:2e-6:2e-5:1e-4,8e-6:4e-3:1e-6:2e-5:1e-4,8e-6:4e-3:5e-7:1e-5:5e-5,8e-6:4e-3:2e-6:5e-5:3e-4" 0,1,2,3,16,17 > gpurun_out/r02_c21_band_sweep.txt 2>&1
grep TOTAL gpurun_out/r02_c21_band_sweep.txt
