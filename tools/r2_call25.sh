set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | grep -v "^  \|array(\[" | cut -c1-300 | tail -15 | tee gpurun_out/r2_gpu_tests_b.log
RMR_TRACE=1 timeout 300 python tools/step_once.py 8 2> gpurun_out/r2_trace.txt; grep rmr_run_once gpurun_out/r2_trace.txt | tail -4
RMR_LETTERBOX_TWO_PASS=1 RMR_TRACE=1 timeout 300 python tools/step_once.py 8 2> gpurun_out/r2_trace_twopass.txt; grep rmr_run_once gpurun_out/r2_trace_twopass.txt | tail -2
timeout 600 python bench.py --steps 200 --warmup 10 --no-library-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_b.json')); r=d['roofline']
print(d['value'], d['e2e']['value'], d['ms_per_step'], r['frac'], r.get('car_net_ms'), r.get('armor_net_ms'), d.get('latency'), d.get('throughput',{}).get('value'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_all.csv python tools/step_once.py 2 > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
