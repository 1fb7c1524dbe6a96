RMR_TRACE=2 timeout 300 python tools/step_once.py 6 2> gpurun_out/r2_trace2.txt; tail -12 gpurun_out/r2_trace2.txt
