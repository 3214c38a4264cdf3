#!/bin/bash
# call 20: CARAFE without branches, full GPU suite, N=1 bench
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/bench_head.py --rois 1024 > gpurun_out/r02_c20_head.json 2> gpurun_out/r02_c20_head.err
cut -c1-300 gpurun_out/r02_c20_head.json; tail -2 gpurun_out/r02_c20_head.err
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_c20_pytest.txt 2>&1
tail -6 gpurun_out/r02_c20_pytest.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c20_bench.json 2> gpurun_out/r02_c20_bench.err
cut -c1-300 gpurun_out/r02_c20_bench.json; tail -2 gpurun_out/r02_c20_bench.err
