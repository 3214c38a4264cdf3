#!/bin/bash
# call 49: head with the tensor-core CARAFE: tests, bench (1024 / 128 RoIs), launch list, pipeline
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_head_gpu.py tests/test_e2e_gpu.py tests/test_score.py tests/test_dropin_gpu.py -m gpu -q 2>&1 | tail -3
timeout 200 python tools/bench_head.py --rois 1024 > gpurun_out/r02_c49_head1024.json 2> gpurun_out/r02_c49_head.err
cut -c1-300 gpurun_out/r02_c49_head1024.json
timeout 200 python tools/bench_head.py --rois 128 > gpurun_out/r02_c49_head128.json 2>/dev/null
cut -c1-200 gpurun_out/r02_c49_head128.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv|carafe|pack|latent" -s 20 -c 10 --csv --log-file gpurun_out/r02_c49_head_launches.csv python tools/bench_head.py --rois 1024 --steps 2 > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_c49_head_launches.csv | awk -F'","' '{print $1, $NF}' | tail -10 | cut -c1-120
timeout 300 python tools/bench_pipeline.py > gpurun_out/r02_c49_pipeline.json 2>/dev/null
cut -c1-500 gpurun_out/r02_c49_pipeline.json
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_head_gpu.py -m gpu -q -k "carafe or whole" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | head -3
