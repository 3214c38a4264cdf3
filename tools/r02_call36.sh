#!/bin/bash
# call 36: 1x1 layers through the double-buffered kernel, packed input re-pack: parity + timing
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_head_gpu.py tests/test_e2e_gpu.py tests/test_score.py -m gpu -q 2>&1 | tail -4
timeout 300 python tools/bench_head.py --rois 1024 > gpurun_out/r02_c36_head1024.json 2> gpurun_out/r02_c36_head.err
cut -c1-300 gpurun_out/r02_c36_head1024.json; tail -2 gpurun_out/r02_c36_head.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv|carafe|pack|latent" -s 20 -c 10 --csv --log-file gpurun_out/r02_c36_head_launches.csv python tools/bench_head.py --rois 1024 --steps 2 > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_c36_head_launches.csv | cut -d, -f5,13- | tail -11 | cut -c1-160
