set -x
mkdir -p gpurun_out
export RMR_CONV_V2=1
timeout 600 python tools/conv_check.py 7 20 > gpurun_out/r2_conv_check_v2e.txt 2>&1
grep -c " ok " gpurun_out/r2_conv_check_v2e.txt; tail -2 gpurun_out/r2_conv_check_v2e.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_gpu_tests_v2e.log
cat gpurun_out/r2_gpu_tests_v2e.log
timeout 300 python tools/profile_layers.py 7 > gpurun_out/r2_layers_v2e.txt 2>&1
grep "^==" gpurun_out/r2_layers_v2e.txt
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_v2e.json 2> gpurun_out/r2_bench_v2e.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v2e.json')); r=d['roofline']
print(d['value'], d['e2e']['value'], r['frac'], r['car_net_ms'], r['armor_net_ms'], r['replayed_alone_ms'])"
timeout 300 python tools/timeline2.py 1,20,20,256,256,3,1 1,40,40,256,512,3,2 1,80,80,128,128,1,1 7,10,10,512,128,3,1 > gpurun_out/r2_timeline_v2e.txt 2>&1
grep -E "^==|median|tile 0:" gpurun_out/r2_timeline_v2e.txt | head -30
