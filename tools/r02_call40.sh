#!/bin/bash
# call 48: tensor-core CARAFE: head tests, per-kernel time, ncu capture of the kernel
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_head_gpu.py tests/test_e2e_gpu.py -m gpu -q 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"carafe" -s 2 -c 2 --csv --log-file gpurun_out/r02_c48_carafe_launches.csv python tools/bench_head.py --rois 1024 --steps 2 > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_c48_carafe_launches.csv | cut -d, -f5,13- | tail -3 | cut -c1-160
