set -x
cd /root/repo
timeout 300 python -m pytest tests/test_pnp_gpu.py -q > gpurun_out/r02_c7_pytest.log 2>&1; tail -5 gpurun_out/r02_c7_pytest.log
timeout 600 python tools/band_timing.py 8192 0:0:0,8e-6:4e-3:2e-6 > gpurun_out/r02_c7_band_timing.txt 2>&1
cat gpurun_out/r02_c7_band_timing.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pnp_lm_fast_kernel -s 2 -c 1 -o gpurun_out/r02_c7_fast_full python tools/prof_run.py 8192 fast full S1 4 0:0:0 > gpurun_out/r02_c7_ncu.log 2>&1
tail -3 gpurun_out/r02_c7_ncu.log
