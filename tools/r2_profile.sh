# round 2 evidence: launch list of one bench step + ncu --set full of every conv2 launch of one step
set -x
mkdir -p gpurun_out
# whole run of two steps; tools/ncu_summary.py keeps the last step (from the last letterbox_fused_kernel on)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_step.csv python tools/step_once.py 2 > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
# conv2 launches per step: count them in the list above, skip two steps' worth, capture the third step
N=$(python - <<'PY'
import csv
rows=[l for l in open('gpurun_out/r2_launches_step.csv') if l.startswith('"')]
r=list(csv.reader(rows)); k=r[0].index("Kernel Name")
print(sum('conv2_kernel' in x[k] for x in r[1:])//2)
PY
)
echo "conv2 launches per step: $N"
timeout 1500 ncu --set full --clock-control none -k regex:conv2_kernel -s $((2*N)) -c $N -o /tmp/r2_conv_full python tools/step_once.py 3 > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/ncu_c.log
ncu -i /tmp/r2_conv_full.ncu-rep --page raw --csv > gpurun_out/r2_conv_full_raw.csv 2>/dev/null
ls -la /tmp/r2_conv_full.ncu-rep gpurun_out/r2_conv_full_raw.csv gpurun_out/r2_launches_step.csv
