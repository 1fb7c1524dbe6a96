# round-end validation + evidence on one B200 (run through gpurun from the repo root)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
timeout 600 python bench.py --config 4 --steps 5 > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err
tail -3 gpurun_out/r2_gpu_tests.log; cut -c1-600 gpurun_out/r2_bench_final.json; tail -2 gpurun_out/r2_bench_final.err; cut -c1-300 gpurun_out/r2_bench_reference.json; tail -2 gpurun_out/r2_smoke.log; cut -c1-400 gpurun_out/r2_bench_c4_n1.json
