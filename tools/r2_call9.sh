set -x
export RMR_CONV_V2=1
timeout 300 python tools/timeline2.py 1,20,20,256,256,3,1 7,80,80,64,64,3,1 7,80,80,128,256,3,2 > gpurun_out/r2_timeline_v2f.txt 2>&1
grep -E "^==|median|tile [01]:" gpurun_out/r2_timeline_v2f.txt | head -30
timeout 600 python tools/conv_check.py 7 20 > gpurun_out/r2_conv_check_v2f.txt 2>&1
grep -c " ok " gpurun_out/r2_conv_check_v2f.txt; tail -2 gpurun_out/r2_conv_check_v2f.txt
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_detect.py -x -q 2>&1 | tail -4
