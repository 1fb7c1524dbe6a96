# round 2 evidence: launch list of one bench step + ncu --set full of conv2 launches of that step
set -x
mkdir -p gpurun_out
# step_once: setup + 3 steps; one step = ~200 launches.  Skip the first two steps' launches, take the third.
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 230 --csv --log-file gpurun_out/r2_launches_step.csv python tools/step_once.py 3 > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
# every conv2 launch of the car net (77) and the first of the armor net at batch 7 of the third step
timeout 900 ncu --set full --clock-control none -k regex:conv2_kernel -s 330 -c 100 -o /tmp/r2_conv_full python tools/step_once.py 3 > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/ncu_c.log
ncu -i /tmp/r2_conv_full.ncu-rep --page raw --csv > gpurun_out/r2_conv_full_raw.csv 2>/dev/null
ls -la /tmp/r2_conv_full.ncu-rep gpurun_out/r2_conv_full_raw.csv gpurun_out/r2_launches_step.csv
