#!/bin/bash
# call 15: GPU suite, bench with the longer warm-up, phase trace of the fast kernel
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_c15_pytest.txt 2>&1
tail -15 gpurun_out/r02_c15_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c15_bench.txt 2> gpurun_out/r02_c15_bench.err
cut -c1-600 gpurun_out/r02_c15_bench.txt; tail -3 gpurun_out/r02_c15_bench.err
timeout 300 python tools/trace_run.py 8192 full fast > gpurun_out/r02_c15_trace_full.txt 2>&1
timeout 300 python tools/trace_run.py 8192 diag fast > gpurun_out/r02_c15_trace_diag.txt 2>&1
cat gpurun_out/r02_c15_trace_full.txt gpurun_out/r02_c15_trace_diag.txt
