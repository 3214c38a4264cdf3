set -x
cd /root/repo
timeout 600 python -m pytest tests/test_pnp_gpu.py -x -q > gpurun_out/r02_c2_pytest_pnp.log 2>&1; tail -15 gpurun_out/r02_c2_pytest_pnp.log
timeout 300 python tools/decision_stats.py 8192 fast default,0:0:0 > gpurun_out/r02_c2_decision.txt 2>&1; tail -3 gpurun_out/r02_c2_decision.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c2_bench_full.json 2> gpurun_out/r02_c2_bench.err; tail -3 gpurun_out/r02_c2_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload diag > gpurun_out/r02_c2_bench_diag.json 2>> gpurun_out/r02_c2_bench.err
cat gpurun_out/r02_c2_bench_full.json | cut -c1-600
