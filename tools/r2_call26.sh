set -x
mkdir -p gpurun_out
timeout 300 python tools/timeline2.py 7,320,320,32,64,3,2 7,160,160,64,64,1,1 7,160,160,32,32,3,1 7,160,160,96,64,1,1 7,160,160,64,128,3,2 7,80,80,128,128,3,1 7,80,80,64,64,3,1 > gpurun_out/r2_timeline_early.txt 2>&1
tail -5 gpurun_out/r2_timeline_early.txt
timeout 300 python tools/library_baseline.py 7 20 > gpurun_out/r2_library_baseline_b.json 2> gpurun_out/r2_library_baseline_b.err; cat gpurun_out/r2_library_baseline_b.json; tail -2 gpurun_out/r2_library_baseline_b.err
