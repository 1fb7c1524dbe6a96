set -x
for i in 1 2; do
RMR_LIB_PATH=$PWD/tools/ab/base.so timeout 600 python bench.py --steps 200 --warmup 10 --no-library-baseline --no-cpu-baseline --no-throughput > gpurun_out/r2_ab_base$i.json 2>/dev/null
timeout 600 python bench.py --steps 200 --warmup 10 --no-library-baseline --no-cpu-baseline --no-throughput > gpurun_out/r2_ab_new$i.json 2>/dev/null
done
python -c "
import json
for f in ('base1','new1','base2','new2'):
    d=json.load(open('gpurun_out/r2_ab_%s.json'%f)); r=d['roofline']
    print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],4), r.get('car_net_ms'), r.get('armor_net_ms'))"
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_conv_modes.py -m gpu -q --tb=short -x 2>&1 | tail -3
