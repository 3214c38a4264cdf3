set -x
cd /root/repo
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r02_c12_pytest_all.log 2>&1; tail -15 gpurun_out/r02_c12_pytest_all.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pnp_lm_fast_kernel -s 2 -c 1 -o gpurun_out/r02_c12_fast_full python tools/prof_run.py 8192 fast full S1 4 0:0:0 > gpurun_out/r02_c12_ncu.log 2>&1
tail -3 gpurun_out/r02_c12_ncu.log
timeout 600 python tools/band_timing.py 8192 0:0:0,8e-6:4e-3:2e-6 fastonly > gpurun_out/r02_c12_band_timing.txt 2>&1
cat gpurun_out/r02_c12_band_timing.txt
