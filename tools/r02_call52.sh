#!/bin/bash
# call 52: CARAFE tests over tile counts and the fallback size
set -x
cd /root/repo
timeout 300 python -m pytest tests/test_head_gpu.py -m gpu -q 2>&1 | tail -4
