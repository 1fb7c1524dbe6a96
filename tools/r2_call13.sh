set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_gpu_tests_v2h.log
cat gpurun_out/r2_gpu_tests_v2h.log
timeout 300 python tools/profile_layers.py 20 > gpurun_out/r2_layers_b20_v2.txt 2>&1
grep "^==" gpurun_out/r2_layers_b20_v2.txt
RMR_CONV_V2=0 timeout 300 python tools/profile_layers.py 20 > gpurun_out/r2_layers_b20_v1.txt 2>&1
grep "^==" gpurun_out/r2_layers_b20_v1.txt
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_v2h.json 2> gpurun_out/r2_bench_v2h.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v2h.json')); r=d['roofline']
print(d['value'], d['e2e']['value'], r['frac'], r['car_net_ms'], r['armor_net_ms'], r['replayed_alone_ms'])"
