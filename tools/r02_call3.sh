set -x
cd /root/repo
timeout 600 python tools/band_timing.py 8192 0:0:0,8e-6:4e-3:2e-6,8e-6:1e-3:5e-7,0:1e9:0 > gpurun_out/r02_c3_band_timing.txt 2>&1
cat gpurun_out/r02_c3_band_timing.txt
