#!/bin/bash
# call 65: final state with the 6-DoF mixed kernel -- full GPU suite, sanitizer (6-DoF legs, three tools), bench lines, smoke,
# side-kernel bench, 6-DoF report and one ncu capture of the mixed kernel
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_c65_pytest.txt 2>&1
tail -5 gpurun_out/r02_c65_pytest.txt
for tool in memcheck racecheck synccheck; do
  MRSAN_ONLY=6dof timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/r02_c65_sanitize_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02_c65_sanitize_$tool.txt | head -3
done
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c65_bench_full.json 2> gpurun_out/r02_c65_bench_full.err
cut -c1-260 gpurun_out/r02_c65_bench_full.json; tail -2 gpurun_out/r02_c65_bench_full.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python tools/bench_noc.py > gpurun_out/r02_c65_noc_bench.txt 2>&1; tail -3 gpurun_out/r02_c65_noc_bench.txt | cut -c1-900
timeout 200 python tools/sixdof_report.py 512 8192 > gpurun_out/r02_c65_6dof.txt 2>&1; grep -v '"parity"' gpurun_out/r02_c65_6dof.txt | tail -8
timeout 200 ncu --set full --import-source on --clock-control none -k regex:pnp_6dof_mixed -s 16 -c 1 -o gpurun_out/r02_c65_6dof_mixed -f python tools/sixdof_report.py 0 8192 mixedonly > gpurun_out/r02_c65_ncu.log 2>&1; tail -2 gpurun_out/r02_c65_ncu.log
