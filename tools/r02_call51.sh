#!/bin/bash
# call 51 (2 GPUs): final check of the multi-GPU path: dist test + bench line
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dist_gpu.py -m gpu -q 2>&1 | tail -3
bash tools/r02_scale.sh 2 "fused"
