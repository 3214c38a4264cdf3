#!/bin/bash
# call 23: why are objects handed back, and how long are their exact solves
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/hand_back_report.py 8192 0,1,2,3 > gpurun_out/r02_c23_hand_back.txt 2>&1
cat gpurun_out/r02_c23_hand_back.txt
