set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r1_gpu_tests.log
python bench.py > gpurun_out/r1_bench_final.json 2> gpurun_out/r1_bench_final.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 400 --csv --log-file gpurun_out/r1_launches_step.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r1_launches_step.csv gpurun_out/r1_launches_step.md > /dev/null
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 161 -c 78 -o /tmp/r1_conv_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
ncu -i /tmp/r1_conv_full.ncu-rep --page raw --csv > gpurun_out/r1_conv_full_raw.csv 2> gpurun_out/ncu_raw.err
python tools/ncu_summary.py full gpurun_out/r1_conv_full_raw.csv gpurun_out/r1_conv_full.json gpurun_out/r1_conv_full.md > /dev/null
ls -la gpurun_out /tmp/r1_conv_full.ncu-rep
tail -3 gpurun_out/r1_gpu_tests.log; cat gpurun_out/r1_bench_final.json | tail -1 | cut -c1-400
