#!/bin/bash
# call 31: split delta passes of long objects (FastTeam): tests, sanitizer, A/B against the build without it
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pnp_gpu.py -m gpu -q -x 2>&1 | tail -4
for tool in racecheck synccheck; do
  timeout 700 compute-sanitizer --tool $tool --print-limit 40 python tools/sanitize_run.py > gpurun_out/r02_c31_sanitize_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02_c31_sanitize_$tool.txt | head -4
  grep -E "Race reported|Barrier error" gpurun_out/r02_c31_sanitize_$tool.txt | cut -c1-200 | sort | uniq -c | sort -rn | head -8
done
timeout 600 python tools/ab_time.py noteam,team 3 8e-6:4e-3:2e-6:2e-5:1e-4:10 > gpurun_out/r02_c31_ab.txt 2>&1
cat gpurun_out/r02_c31_ab.txt
timeout 600 python tools/band_sweep.py 8192 "8e-6:4e-3:2e-6:2e-5:1e-4:10" 0,1,2,3,16,17 > gpurun_out/r02_c31_band_sweep.txt 2>&1
grep TOTAL gpurun_out/r02_c31_band_sweep.txt
