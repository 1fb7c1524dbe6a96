set -x
RMR_TRACE=2 timeout 300 python tools/step_once.py 8 2> gpurun_out/r2_trace_a.txt; grep collect gpurun_out/r2_trace_a.txt | tail -4
RMR_POST_TWO_PASS=1 RMR_TRACE=2 timeout 300 python tools/step_once.py 8 2> gpurun_out/r2_trace_b.txt; grep collect gpurun_out/r2_trace_b.txt | tail -4
for i in 1 2; do
RMR_POST_TWO_PASS=1 timeout 600 python bench.py --steps 200 --warmup 10 --no-library-baseline --no-cpu-baseline --no-throughput > gpurun_out/r2_ab_two$i.json 2>/dev/null
timeout 600 python bench.py --steps 200 --warmup 10 --no-library-baseline --no-cpu-baseline --no-throughput > gpurun_out/r2_ab_new$i.json 2>/dev/null
done
python -c "
import json
for f in ('two1','new1','two2','new2'):
    d=json.load(open('gpurun_out/r2_ab_%s.json'%f)); r=d['roofline']
    print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],4), r.get('car_net_ms'), r.get('armor_net_ms'))"
