#!/bin/bash
# call 29: compute-sanitizer over every kernel variant incl. the redo phase, consensus prune, score stage
set -x
cd /root/repo
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 700 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/r02_c29_sanitize_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|valid|redo phase|consensus|hazard|Invalid|Barrier error" gpurun_out/r02_c29_sanitize_$tool.txt | head -20
done
