#!/bin/bash
# call 50: final state -- full GPU suite, sanitizer (three tools), bench lines
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_c50_pytest.txt 2>&1
tail -5 gpurun_out/r02_c50_pytest.txt
for tool in memcheck racecheck synccheck; do
  timeout 700 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/r02_c50_sanitize_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02_c50_sanitize_$tool.txt | head -3
done
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c50_bench_full.json 2> gpurun_out/r02_c50_bench_full.err
cut -c1-260 gpurun_out/r02_c50_bench_full.json; tail -2 gpurun_out/r02_c50_bench_full.err
timeout 400 python bench.py --steps 20 --warmup 5 --workload diag --no-cpu-baseline > gpurun_out/r02_c50_bench_diag.json 2> gpurun_out/r02_c50_bench_diag.err
cut -c1-260 gpurun_out/r02_c50_bench_diag.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
