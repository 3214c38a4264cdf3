#!/bin/bash
# call 35: ncu capture of the head's convolution kernels (new 3x3 kernel with two TMEM accumulator sets) + launch list of the head
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv -s 14 -c 7 -f -o gpurun_out/r02_head_conv python tools/bench_head.py --rois 1024 --steps 2 > gpurun_out/r02_c35_ncu.log 2>&1
tail -3 gpurun_out/r02_c35_ncu.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv|carafe|pack|latent" -s 20 -c 10 --csv --log-file gpurun_out/r02_c35_head_launches.csv python tools/bench_head.py --rois 1024 --steps 2 > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_c35_head_launches.csv | cut -d, -f5,13- | tail -12 | cut -c1-200
