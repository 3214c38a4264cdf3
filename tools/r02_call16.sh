#!/bin/bash
# call 16: A/B of the shared-memory warp reduction, split first passes, delta unroll 2
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python tools/ab_time.py base_shfl,smem,smem_split,smem_u2 2 > gpurun_out/r02_c16_ab.txt 2>&1
cat gpurun_out/r02_c16_ab.txt
MRPNP_LIB=tools/ab/smem.so timeout 600 python -m pytest tests/test_pnp_gpu.py -m gpu -x -q 2>&1 | tail -5
