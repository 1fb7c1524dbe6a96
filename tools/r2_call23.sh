set -x
timeout 900 python -m pytest tests/test_gpu_detect.py tests/test_gpu_all_assets.py tests/test_gpu_conv.py -q --tb=short 2>&1 | grep -v "^  \|array(\[" | cut -c1-300 | tail -12
timeout 300 python tools/profile_layers.py 7 > gpurun_out/r2_layers_cap110.txt 2>&1; grep "^==" gpurun_out/r2_layers_cap110.txt; grep -E "^  0 " gpurun_out/r2_layers_cap110.txt
RMR_SMEM_CAP_KB=0 timeout 300 python tools/profile_layers.py 7 > gpurun_out/r2_layers_cap0.txt 2>&1; grep "^==" gpurun_out/r2_layers_cap0.txt
RMR_SMEM_CAP_KB=0 timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-throughput --no-library-baseline > gpurun_out/r2_bench_cap0.json 2>/dev/null
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-throughput --no-library-baseline > gpurun_out/r2_bench_cap110.json 2>/dev/null
python -c "
import json
for f in ('cap0','cap110'):
    d=json.load(open('gpurun_out/r2_bench_%s.json'%f)); r=d['roofline']
    print(f, d['value'], d['e2e']['value'], r['frac'], r['car_net_ms'], r['armor_net_ms'])"
