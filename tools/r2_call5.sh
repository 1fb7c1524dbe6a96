# round 2, GPU call 5: where a conv2 CTA's time goes (timeline + ncu source view), reference-kernel tests
set -x
mkdir -p gpurun_out
export RMR_CONV_V2=1
timeout 300 python tools/timeline2.py > gpurun_out/r2_timeline_v2.txt 2>&1
cat gpurun_out/r2_timeline_v2.txt
timeout 600 python -m pytest tests/test_gpu_ref_kernels.py tests/test_gpu_jpeg.py -q 2>&1 | tail -15 > gpurun_out/r2_test_ref.log
cat gpurun_out/r2_test_ref.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2_kernel -s 3 -c 1 -o gpurun_out/r2_conv2_a python tools/timeline2.py 7,160,160,32,32,3,1 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2_kernel -s 3 -c 1 -o gpurun_out/r2_conv2_b python tools/timeline2.py 7,80,80,128,128,3,1 > gpurun_out/ncu_b.log 2>&1
for f in a b; do
  ncu -i gpurun_out/r2_conv2_$f.ncu-rep --page raw --csv > gpurun_out/r2_conv2_${f}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2_conv2_$f.ncu-rep --page source --csv > gpurun_out/r2_conv2_${f}_source.csv 2>/dev/null
done
ls -la gpurun_out | grep conv2_
