set -x
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | grep -v "^  \|array(\[" | cut -c1-300 | tail -6
RMR_TRACE=1 timeout 300 python tools/step_once.py 6 2> gpurun_out/r2_trace4.txt; grep rmr_run_once gpurun_out/r2_trace4.txt | tail -2
for i in 1 2; do
RMR_SYNC_WAIT=1 timeout 600 python bench.py --steps 200 --warmup 10 --no-library-baseline --no-cpu-baseline --no-throughput > gpurun_out/r2_ab_sync$i.json 2>/dev/null
timeout 600 python bench.py --steps 200 --warmup 10 --no-library-baseline --no-cpu-baseline --no-throughput > gpurun_out/r2_ab_new$i.json 2>/dev/null
done
python -c "
import json
for f in ('sync1','new1','sync2','new2'):
    d=json.load(open('gpurun_out/r2_ab_%s.json'%f)); r=d['roofline']
    print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],4), r.get('car_net_ms'), r.get('armor_net_ms'))"
