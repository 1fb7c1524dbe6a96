set -x
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_conv_modes.py tests/test_gpu_detect.py -m gpu -q --tb=short -x 2>&1 | grep -v "^  \|array(\[" | cut -c1-300 | tail -8
timeout 300 python tools/timeline4.py 7,160,160,64,64,1,1 7,160,160,32,32,3,1 > gpurun_out/r2_timeline_epi2.txt 2>&1
grep -A8 "cta 74" gpurun_out/r2_timeline_epi2.txt | cut -c1-200
timeout 600 python bench.py --steps 200 --warmup 10 --no-library-baseline --no-cpu-baseline > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_c.json')); r=d['roofline']
print(d['value'], d['e2e']['value'], d['ms_per_step'], r['frac'], r.get('car_net_ms'), r.get('armor_net_ms'), d.get('throughput',{}).get('value'), d.get('throughput',{}).get('roofline',{}).get('frac'))"
