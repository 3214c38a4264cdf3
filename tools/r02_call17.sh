#!/bin/bash
# call 17 (2 GPUs): fused gather with completion flags -- dist test, bench N=2 flags / barrier / nccl, N=1
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_pnp_gpu.py -m gpu -x -q 2>&1 | tail -5
for g in fused fused-barrier nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 40 --warmup 5 --gather $g --sustain-seconds 0 > gpurun_out/r02_c17_n2_$g.json 2> gpurun_out/r02_c17_n2_$g.err
  tail -2 gpurun_out/r02_c17_n2_$g.err; cut -c1-220 gpurun_out/r02_c17_n2_$g.json
done
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --sustain-seconds 0 > gpurun_out/r02_c17_n1.json 2> gpurun_out/r02_c17_n1.err
cut -c1-220 gpurun_out/r02_c17_n1.json
