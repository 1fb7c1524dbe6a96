set -x
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_gpu_tests_v2k.log
cat gpurun_out/r2_gpu_tests_v2k.log
timeout 600 python tools/library_baseline.py 7 20 > gpurun_out/r2_library_baseline.json 2> gpurun_out/r2_library_baseline.err
cat gpurun_out/r2_library_baseline.json; tail -3 gpurun_out/r2_library_baseline.err
