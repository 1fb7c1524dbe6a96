set -x
mkdir -p gpurun_out
export RMR_CONV_V2=1
timeout 600 python tools/conv_check.py 7 20 > gpurun_out/r2_conv_check_v2d.txt 2>&1
grep -c ok gpurun_out/r2_conv_check_v2d.txt; grep -v " ok " gpurun_out/r2_conv_check_v2d.txt | tail -12
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -8
timeout 300 python tools/timeline2.py 7,160,160,32,32,3,1 7,80,80,64,64,3,1 7,80,80,128,128,3,1 7,160,160,64,64,1,1 > gpurun_out/r2_timeline_v2c.txt 2>&1
grep -E "^==|median|tile [12]:" gpurun_out/r2_timeline_v2c.txt | head -40
RMR_TMA_EPI=0 timeout 600 python tools/conv_check.py 7 20 > gpurun_out/r2_conv_check_v2d_noepi.txt 2>&1
tail -2 gpurun_out/r2_conv_check_v2d_noepi.txt
