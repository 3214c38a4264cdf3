#!/bin/bash
# strong scaling of BASELINE configs[4]: 65,536 objects in total over N GPUs
N=$1
set -x
cd /root/repo
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --total-objects 65536 --sustain-seconds 0 --no-cpu-baseline > gpurun_out/r02_strong_n1.json 2> gpurun_out/r02_strong_n1.err
else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 --total-objects 65536 --sustain-seconds 0 > gpurun_out/r02_strong_n$N.json 2> gpurun_out/r02_strong_n$N.err
fi
echo "rc=$?"; tail -2 gpurun_out/r02_strong_n$N.err; cut -c1-260 gpurun_out/r02_strong_n$N.json
