set -x
timeout 300 python tools/stress_forward.py 300 3 7 2>&1 | tail -3
for i in 1 2 3 4; do
timeout 600 python bench.py --steps 200 --warmup 10 --no-library-baseline --no-cpu-baseline > gpurun_out/r2_ab_new$i.json 2> gpurun_out/r2_ab_new$i.err; echo rc=$?; tail -2 gpurun_out/r2_ab_new$i.err | cut -c1-200
done
python -c "
import json
for f in ('new1','new2','new3','new4'):
    d=json.load(open('gpurun_out/r2_ab_%s.json'%f)); r=d['roofline']
    print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],4), r.get('car_net_ms'), r.get('armor_net_ms'), r['frac'], d['throughput']['value'], d['throughput']['roofline']['frac'])"
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | grep -v "^  \|array(\[" | cut -c1-300 | tail -6
