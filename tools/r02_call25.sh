#!/bin/bash
# call 25: adaptive band for slowly converging objects only (from 6 / 8 / 10 evaluations)
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1400 python tools/band_sweep.py 8192 "8e-6:4e-3:2e-6,8e-6:4e-3:2e-6:2e-5:1e-4:6,8e-6:4e-3:2e-6:2e-5:1e-4:8,8e-6:4e-3:2e-6:2e-5:1e-4:10,8e-6:4e-3:2e-6:5e-5:2e-4:8,8e-6:4e-3:2e-6:1e-5:5e-5:8" 0,1,2,3,4,5,6,7,16,17,18,19 > gpurun_out/r02_c25_band_sweep.txt 2>&1
grep TOTAL gpurun_out/r02_c25_band_sweep.txt
