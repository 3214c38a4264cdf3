set -x
cd /root/repo
timeout 600 python tools/band_timing.py 8192 0:0:0,8e-6:4e-3:2e-6,8e-6:1e-2:2e-6,8e-6:1e-3:5e-7 > gpurun_out/r02_c4_band_timing.txt 2>&1
cat gpurun_out/r02_c4_band_timing.txt
timeout 300 python tools/decision_stats.py 8192 fast default,8e-6:1e-2:2e-6 > gpurun_out/r02_c4_decision.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pnp_lm_fast_kernel -s 2 -c 1 -o gpurun_out/r02_c4_fast_full python tools/prof_run.py 8192 fast full S1 4 > gpurun_out/r02_c4_ncu.log 2>&1
tail -3 gpurun_out/r02_c4_ncu.log
