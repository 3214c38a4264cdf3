#!/bin/bash
# call 19: 3x3 convolution with one activation load per K chunk (shifted operand views): parity in both descriptor modes, timing
set -x
cd /root/repo
mkdir -p gpurun_out
for mode in reuse128 reuse256 pertap; do
  echo "== MRHEAD_CONV=$mode"
  MRHEAD_CONV=$mode timeout 600 python -m pytest tests/test_head_gpu.py -m gpu -x -q 2>&1 | tail -4
  MRHEAD_CONV=$mode timeout 600 python tools/bench_head.py --rois 1024 > gpurun_out/r02_c19_head_$mode.json 2> gpurun_out/r02_c19_head_$mode.err
  cut -c1-400 gpurun_out/r02_c19_head_$mode.json; tail -2 gpurun_out/r02_c19_head_$mode.err
done
