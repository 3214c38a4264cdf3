#!/bin/bash
# call 61: 6-DoF mixed kernel with the single-lane controller: tests, racecheck, timing and parity report
cd "$(dirname "$0")/.."
timeout 400 python -m pytest tests/test_6dof_gpu.py -m gpu -x -q 2>&1 | tail -15
for tool in racecheck memcheck; do
  MRSAN_ONLY=6dof timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/r02_c61_sanitize_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|6dof|hazard|Invalid|Barrier error" gpurun_out/r02_c61_sanitize_$tool.txt | head -12
done
timeout 200 python tools/sixdof_report.py 512 8192 > gpurun_out/r02_c61_6dof.txt 2>&1; grep -v '"parity"' gpurun_out/r02_c61_6dof.txt | tail -8
