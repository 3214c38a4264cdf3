#!/bin/bash
# call 28: CARAFE two pixels per iteration (head tests + timing), strong-scaling point at N = 1
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_head_gpu.py tests/test_score.py tests/test_e2e_gpu.py -m gpu -q 2>&1 | tail -4
timeout 300 python tools/bench_head.py --rois 1024 > gpurun_out/r02_c28_head1024.json 2> gpurun_out/r02_c28_head.err
cut -c1-300 gpurun_out/r02_c28_head1024.json; tail -2 gpurun_out/r02_c28_head.err
bash tools/r02_scale_strong.sh 1
