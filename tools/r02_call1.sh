set -x
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench4 tools/microbench4.cu && ./tools/microbench4 > gpurun_out/r02_microbench4.txt 2>&1
python tools/decision_stats.py 8192 mixed,fast > gpurun_out/r02_decision_stats_base.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_base.json 2> gpurun_out/r02_bench_base.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload diag > gpurun_out/r02_bench_base_diag.json 2>> gpurun_out/r02_bench_base.err
tail -5 gpurun_out/r02_microbench4.txt gpurun_out/r02_decision_stats_base.txt
