set -x
cd /root/repo
timeout 600 python -m pytest tests/test_pnp_gpu.py -q -x > gpurun_out/r02_c10_pytest.log 2>&1; tail -15 gpurun_out/r02_c10_pytest.log
timeout 900 python tools/ab_time.py round1,head,scal,scal_u1 2 0:0:0 > gpurun_out/r02_c10_ab.txt 2>&1
cat gpurun_out/r02_c10_ab.txt
