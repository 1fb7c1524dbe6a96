set -x
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_conv_modes.py tests/test_gpu_detect.py tests/test_gpu_all_assets.py -m gpu -q --tb=short -x 2>&1 | tail -3
for i in 1 2; do
RMR_EXIT_WAIT_ALL=1 timeout 600 python bench.py --steps 200 --warmup 10 --no-library-baseline --no-cpu-baseline --no-throughput > gpurun_out/r2_ab_wait$i.json 2>/dev/null
timeout 600 python bench.py --steps 200 --warmup 10 --no-library-baseline --no-cpu-baseline --no-throughput > gpurun_out/r2_ab_new$i.json 2>/dev/null
done
python -c "
import json
for f in ('wait1','new1','wait2','new2'):
    d=json.load(open('gpurun_out/r2_ab_%s.json'%f)); r=d['roofline']
    print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],4), r.get('car_net_ms'), r.get('armor_net_ms'), r['frac'])"
timeout 300 python tools/timeline2.py 1,80,80,64,64,3,1 1,40,40,256,256,1,1 2>&1 | grep -E "^==|median" | cut -c1-200
