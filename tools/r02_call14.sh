#!/bin/bash
# call 14: full GPU suite after the score-stage kernel, the in-prologue dimension decode and the packed default
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c14_pytest.txt 2>&1
tail -15 gpurun_out/r02_c14_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c14_bench.txt 2> gpurun_out/r02_c14_bench.err
cat gpurun_out/r02_c14_bench.txt; tail -3 gpurun_out/r02_c14_bench.err
