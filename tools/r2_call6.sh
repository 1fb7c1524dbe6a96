set -x
mkdir -p gpurun_out
export RMR_CONV_V2=1
timeout 300 python tools/timeline2.py 7,160,160,32,32,3,1 7,80,80,64,64,3,1 7,80,80,128,128,3,1 1,20,20,256,256,3,1 7,80,80,128,256,3,2 > gpurun_out/r2_timeline_v2b.txt 2>&1
grep -E "^==|median|tile [12]:" gpurun_out/r2_timeline_v2b.txt | head -60
timeout 600 python tools/conv_check.py 7 20 > gpurun_out/r2_conv_check_v2c.txt 2>&1
tail -3 gpurun_out/r2_conv_check_v2c.txt
