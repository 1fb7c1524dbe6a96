set -x
export RMR_CONV_V2=1
timeout 600 python tools/conv_check.py 7 20 > gpurun_out/r2_conv_check_v2g.txt 2>&1
grep -c " ok " gpurun_out/r2_conv_check_v2g.txt; grep -v " ok " gpurun_out/r2_conv_check_v2g.txt | tail -5
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_detect.py -x -q 2>&1 | tail -4
timeout 300 python tools/profile_layers.py 7 > gpurun_out/r2_layers_v2g.txt 2>&1
grep "^==" gpurun_out/r2_layers_v2g.txt
timeout 300 python tools/timeline2.py 1,20,20,256,256,3,1 7,80,80,64,64,3,1 7,80,80,128,256,3,2 7,40,40,128,128,3,1 1,80,80,256,128,1,1 > gpurun_out/r2_timeline_v2g.txt 2>&1
grep -E "^==|median|tile [01]:" gpurun_out/r2_timeline_v2g.txt | head -40
