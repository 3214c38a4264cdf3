#!/bin/bash
# call 30: predicated compaction stores + mbarrier team barrier: tests, sanitizer, bench
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pnp_gpu.py -m gpu -q -x 2>&1 | tail -4
for tool in racecheck synccheck; do
  timeout 700 compute-sanitizer --tool $tool --print-limit 40 python tools/sanitize_run.py > gpurun_out/r02_c30_sanitize_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|valid|redo phase|consensus" gpurun_out/r02_c30_sanitize_$tool.txt | head -12
  grep -E "Race reported|Barrier error|at mrpnp" gpurun_out/r02_c30_sanitize_$tool.txt | cut -c1-220 | sort | uniq -c | sort -rn | head -12
done
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c30_bench_full.json 2> gpurun_out/r02_c30_bench_full.err
cut -c1-260 gpurun_out/r02_c30_bench_full.json; tail -2 gpurun_out/r02_c30_bench_full.err
