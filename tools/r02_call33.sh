#!/bin/bash
# call 33: 6-DoF kernel with the inlier index list: parity tests + timing
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_6dof_gpu.py tests/test_noc_gpu.py -m gpu -q 2>&1 | tail -4
timeout 600 python tools/bench_noc.py > gpurun_out/r02_c33_noc_bench.txt 2>&1
tail -6 gpurun_out/r02_c33_noc_bench.txt | cut -c1-600
