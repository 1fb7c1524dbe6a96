set -x
timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^  \|array(\[" | cut -c1-300 | tail -25 > gpurun_out/r2_gpu_tests_v2l.log
cat gpurun_out/r2_gpu_tests_v2l.log
timeout 900 python bench.py --config 4 --steps 5 > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err
cut -c1-1200 gpurun_out/r2_bench_c4_n1.json; tail -3 gpurun_out/r2_bench_c4_n1.err
