set -x
cd /root/repo
timeout 900 python tools/ab_time.py round1,head,cur,nol2pf,noearly,unroll1 3 0:0:0 > gpurun_out/r02_c8_ab.txt 2>&1
cat gpurun_out/r02_c8_ab.txt
