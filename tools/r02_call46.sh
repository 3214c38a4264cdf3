#!/bin/bash
# call 46: where does the tensor-core CARAFE spend its time: variants without MMAs / TMA loads / global stores (timing only)
set -x
cd /root/repo
mkdir -p gpurun_out
for v in nomma notma nostore; do
  MRHEAD_LIB=tools/ab/head_$v.so timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"carafe" -s 2 -c 1 --csv --log-file gpurun_out/r02_c46_$v.csv python tools/bench_head.py --rois 1024 --steps 2 > /dev/null 2>&1
  echo "$v $(grep -v '^==' gpurun_out/r02_c46_$v.csv | tail -1 | cut -d, -f5,16)"
done
