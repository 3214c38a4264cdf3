set -x
cd /root/repo
./tools/microbench5 > gpurun_out/r02_c9_microbench5.txt 2>&1; cat gpurun_out/r02_c9_microbench5.txt
timeout 600 python -m pytest tests/test_pnp_gpu.py -q -x > gpurun_out/r02_c9_pytest.log 2>&1; tail -30 gpurun_out/r02_c9_pytest.log
