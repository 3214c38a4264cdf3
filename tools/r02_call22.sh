#!/bin/bash
# call 22: handed-back objects solved by whole CTAs in a redo phase (team evaluation) vs inline by the appending warp
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pnp_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/band_sweep.py 8192 "0:0:0,8e-6:4e-3:2e-6,8e-6:4e-3:2e-6:2e-5:1e-4" 0,1,2,3,16,17 > gpurun_out/r02_c22_band_sweep_team.txt 2>&1
grep TOTAL gpurun_out/r02_c22_band_sweep_team.txt
MRPNP_LIB=tools/ab/inline_redo.so timeout 600 python tools/band_sweep.py 8192 "8e-6:4e-3:2e-6" 0,1,2,3,16,17 > gpurun_out/r02_c22_band_sweep_inline.txt 2>&1
grep TOTAL gpurun_out/r02_c22_band_sweep_inline.txt
