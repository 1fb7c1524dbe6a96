# usage: bash tools/r2_multi.sh N
N=$1
set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -3 gpurun_out/r2_bench_n$N.err; cut -c1-400 gpurun_out/r2_bench_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config 4 --steps 5 > gpurun_out/r2_bench_c4_n$N.json 2> gpurun_out/r2_bench_c4_n$N.err
tail -3 gpurun_out/r2_bench_c4_n$N.err; cut -c1-700 gpurun_out/r2_bench_c4_n$N.json
