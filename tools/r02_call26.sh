#!/bin/bash
# call 26: full GPU suite, bench lines, ncu launch list + full capture of the solver, phase trace, config-4 end to end, head
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_c26_pytest.txt 2>&1
tail -8 gpurun_out/r02_c26_pytest.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c26_bench_full.json 2> gpurun_out/r02_c26_bench_full.err
cut -c1-260 gpurun_out/r02_c26_bench_full.json; tail -2 gpurun_out/r02_c26_bench_full.err
timeout 400 python bench.py --steps 20 --warmup 5 --workload diag --no-cpu-baseline > gpurun_out/r02_c26_bench_diag.json 2> gpurun_out/r02_c26_bench_diag.err
cut -c1-260 gpurun_out/r02_c26_bench_diag.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c26_bench_ref.json 2> gpurun_out/r02_c26_bench_ref.err
cut -c1-400 gpurun_out/r02_c26_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c26_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sustain-seconds 0 --large-batch 0 > gpurun_out/r02_c26_b_ncu.log 2>&1
tail -2 gpurun_out/r02_c26_b_ncu.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pnp_lm_fast_kernel -s 2 -c 1 -f -o gpurun_out/r02_c26_fast_full python tools/prof_run.py 8192 fast full S1 4 > gpurun_out/r02_c26_ncu.log 2>&1
tail -2 gpurun_out/r02_c26_ncu.log
timeout 300 python tools/trace_run.py 8192 full fast > gpurun_out/r02_c26_trace_full.txt 2>&1
timeout 300 python tools/trace_run.py 8192 diag fast > gpurun_out/r02_c26_trace_diag.txt 2>&1
head -12 gpurun_out/r02_c26_trace_full.txt
timeout 600 python tools/e2e_config4.py --out gpurun_out/r02_c26_e2e_config4.json > gpurun_out/r02_c26_e2e.log 2>&1
tail -30 gpurun_out/r02_c26_e2e.log
timeout 300 python tools/bench_head.py --rois 1024 > gpurun_out/r02_c26_head1024.json 2> gpurun_out/r02_c26_head.err
timeout 300 python tools/bench_head.py --rois 128 > gpurun_out/r02_c26_head128.json 2>> gpurun_out/r02_c26_head.err
cut -c1-300 gpurun_out/r02_c26_head1024.json
timeout 300 python tools/bench_pipeline.py > gpurun_out/r02_c26_pipeline.json 2> gpurun_out/r02_c26_pipeline.err
cut -c1-600 gpurun_out/r02_c26_pipeline.json; tail -3 gpurun_out/r02_c26_pipeline.err
