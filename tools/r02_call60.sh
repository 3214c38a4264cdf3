#!/bin/bash
# call 60: the 6-DoF mixed kernel: its tests, compute-sanitizer over both 6-DoF kernels
cd "$(dirname "$0")/.."
timeout 400 python -m pytest tests/test_6dof_gpu.py -m gpu -x -q 2>&1 | tail -15
for tool in memcheck racecheck synccheck; do
  MRSAN_ONLY=6dof timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/r02_c60_sanitize_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|6dof|hazard|Invalid|Barrier error" gpurun_out/r02_c60_sanitize_$tool.txt | head -12
done
