#!/bin/bash
# call 27: early CTAs leave the redo phase; tests + bench
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pnp_gpu.py tests/test_head_gpu.py tests/test_score.py tests/test_dropin_gpu.py -m gpu -q 2>&1 | tail -4
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c27_bench_full.json 2> gpurun_out/r02_c27_bench_full.err
cut -c1-260 gpurun_out/r02_c27_bench_full.json; tail -2 gpurun_out/r02_c27_bench_full.err
timeout 400 python bench.py --steps 20 --warmup 5 --streams 1 --no-cpu-baseline > gpurun_out/r02_c27_bench_full_s1.json 2> gpurun_out/r02_c27_bench_full_s1.err
cut -c1-260 gpurun_out/r02_c27_bench_full_s1.json
timeout 400 python bench.py --steps 20 --warmup 5 --streams 3 --no-cpu-baseline > gpurun_out/r02_c27_bench_full_s3.json 2> gpurun_out/r02_c27_bench_full_s3.err
cut -c1-260 gpurun_out/r02_c27_bench_full_s3.json
