#!/bin/bash
# call 39: CARAFE reassembly as a banded GEMM on the tensor cores: parity (short timeouts), then timing
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_head_gpu.py -m gpu -q -x -k carafe 2>&1 | tail -15
rc=$?
timeout 300 python -m pytest tests/test_head_gpu.py -m gpu -q 2>&1 | tail -4
timeout 200 python tools/bench_head.py --rois 1024 > gpurun_out/r02_c39_head_tc.json 2> gpurun_out/r02_c39_head_tc.err
cut -c1-300 gpurun_out/r02_c39_head_tc.json; tail -2 gpurun_out/r02_c39_head_tc.err
MRHEAD_CARAFE=fma timeout 200 python tools/bench_head.py --rois 1024 > gpurun_out/r02_c39_head_fma.json 2>/dev/null
cut -c1-300 gpurun_out/r02_c39_head_fma.json
