# round-end validation + evidence on one B200 (run through gpurun from the repo root)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r1_gpu_tests.log
python bench.py > gpurun_out/r1_bench_final.json 2> gpurun_out/r1_bench_final.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
python tools/jpeg_bench.py 200 > gpurun_out/r1_jpeg_bench.json 2> gpurun_out/r1_jpeg_bench.err
RMR_JPEG_SIMPLE=1 python tools/jpeg_bench.py 200 > gpurun_out/r1_jpeg_bench_simple.json 2>> gpurun_out/r1_jpeg_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jpeg -s 35 -c 14 --csv --log-file gpurun_out/r1_jpeg_launches.csv python tools/jpeg_bench.py 10 > gpurun_out/ncu_j.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jpeg_entropy_kernel -s 5 -c 1 -o /tmp/jpeg_entropy python tools/jpeg_bench.py 10 > gpurun_out/ncu_jf.log 2>&1
ncu -i /tmp/jpeg_entropy.ncu-rep --page raw --csv > gpurun_out/r1_jpeg_entropy_raw.csv 2> /dev/null
ncu -i /tmp/jpeg_entropy.ncu-rep --page source --csv > gpurun_out/r1_jpeg_entropy_source.csv 2> /dev/null
tail -3 gpurun_out/r1_gpu_tests.log; tail -1 gpurun_out/r1_bench_final.json | cut -c1-300; cat gpurun_out/r1_smoke.log | tail -2; tail -1 gpurun_out/r1_jpeg_bench.json | cut -c1-400
