set -x
timeout 300 python tools/timeline4.py 7,160,160,64,64,1,1 7,160,160,32,32,3,1 7,80,80,64,64,3,1 1,160,160,64,64,1,1 > gpurun_out/r2_timeline_epi.txt 2>&1
tail -3 gpurun_out/r2_timeline_epi.txt
