# round 2, GPU call 3: multi-issuer / lean-ingest microbenchmark + first parity/timing pass of conv2
set -x
mkdir -p gpurun_out
timeout 300 tools/umma_bench q > gpurun_out/r2_umma_bench2.txt 2>&1
cat gpurun_out/r2_umma_bench2.txt
RMR_CONV_V2=1 timeout 600 python tools/conv_check.py 7 20 > gpurun_out/r2_conv_check_v2.txt 2>&1
tail -75 gpurun_out/r2_conv_check_v2.txt
RMR_CONV_V2=1 RMR_HALO=0 timeout 600 python tools/conv_check.py 7 20 > gpurun_out/r2_conv_check_v2_nohalo.txt 2>&1
tail -75 gpurun_out/r2_conv_check_v2_nohalo.txt
RMR_CONV_V2=0 timeout 600 python tools/conv_check.py 7 20 > gpurun_out/r2_conv_check_v1.txt 2>&1
tail -75 gpurun_out/r2_conv_check_v1.txt
RMR_CONV_V2=1 timeout 900 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -15 > gpurun_out/r2_test_conv.log
cat gpurun_out/r2_test_conv.log
