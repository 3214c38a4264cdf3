#!/bin/bash
# multi-GPU bench lines: tools/r02_scale.sh N "gather modes" [extra bench args]; every run under its own short timeout
N=$1; MODES=$2; shift 2
set -x
cd /root/repo
mkdir -p gpurun_out
for g in $MODES; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 40 --warmup 5 --gather $g --sustain-seconds 0.5 "$@" > gpurun_out/r02_scale_n${N}_$g.json 2> gpurun_out/r02_scale_n${N}_$g.err
  echo "rc=$?"; tail -2 gpurun_out/r02_scale_n${N}_$g.err; cut -c1-260 gpurun_out/r02_scale_n${N}_$g.json
done
