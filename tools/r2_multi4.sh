N=$1
set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo rc=$?
tail -3 gpurun_out/r2_bench_n$N.err; cut -c1-300 gpurun_out/r2_bench_n$N.json
