# round 2, GPU call 4: conv2 without planned split-K: per-shape check, whole-network tests, per-layer profile, bench
set -x
mkdir -p gpurun_out
export RMR_CONV_V2=1
timeout 600 python tools/conv_check.py 7 20 > gpurun_out/r2_conv_check_v2b.txt 2>&1
tail -3 gpurun_out/r2_conv_check_v2b.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_gpu_tests_v2.log
cat gpurun_out/r2_gpu_tests_v2.log
timeout 300 python tools/profile_layers.py 7 > gpurun_out/r2_layers_v2.txt 2>&1
grep "^==" gpurun_out/r2_layers_v2.txt
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_v2.json 2> gpurun_out/r2_bench_v2.err
cut -c1-1500 gpurun_out/r2_bench_v2.json
RMR_CONV_V2=0 timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_v1.json 2> gpurun_out/r2_bench_v1.err
cut -c1-600 gpurun_out/r2_bench_v1.json
