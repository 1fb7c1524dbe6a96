export RMR_CONV_V2=1
timeout 300 python tools/timeline3.py 1,20,20,256,256,3,1 7,80,80,128,256,3,2 1,80,80,256,128,1,1 2>&1 | tee gpurun_out/r2_timeline3.txt
