set -x
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_gpu_tests_v2j.log
cat gpurun_out/r2_gpu_tests_v2j.log
timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_v2j.json 2> gpurun_out/r2_bench_v2j.err
tail -5 gpurun_out/r2_bench_v2j.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_v2j.json')); r=d['roofline']
print(d['value'], d['e2e']['value'], r['frac'], r['car_net_ms'], r['armor_net_ms'])
print(d.get('latency')); print(d.get('throughput')); print(d.get('library_baseline'))"
