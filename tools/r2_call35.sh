set -x
for i in 1 2 3 4; do
timeout 600 python bench.py --steps 200 --warmup 10 --no-library-baseline --no-cpu-baseline > gpurun_out/r2_ab_new$i.json 2> gpurun_out/r2_ab_new$i.err; echo rc=$?; tail -3 gpurun_out/r2_ab_new$i.err
done
