set -x
timeout 900 python -m pytest tests/test_gpu_detect.py tests/test_gpu_ref_kernels.py tests/test_gpu_conv.py -q 2>&1 | tail -6
timeout 300 python tools/profile_layers.py 7 > gpurun_out/r2_layers_v2i.txt 2>&1
grep "^==" gpurun_out/r2_layers_v2i.txt
grep -E "^ (40|46|47|35|41) " gpurun_out/r2_layers_v2i.txt | tail -6
