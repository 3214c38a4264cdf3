#!/bin/bash
# Build A/B variants of libmonorun_pnp.so into tools/ab/<name>.so: tools/ab_build.sh name "-DFLAG ..." [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/ab
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -ccbin /usr/bin/g++ -I include $flags monorun_b200/csrc/pnp_capi.cu -o tools/ab/$name.so &
done
wait
ls -la tools/ab/
