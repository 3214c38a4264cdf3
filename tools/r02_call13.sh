set -x
cd /root/repo
timeout 900 python tools/ab_time.py round1,scal_w8u1,c_w8,c_w10,c_w8p3,c_w10p3 2 0:0:0 > gpurun_out/r02_c13_ab.txt 2>&1
cat gpurun_out/r02_c13_ab.txt
