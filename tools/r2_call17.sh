timeout 900 python -m pytest tests/test_gpu_all_assets.py "tests/test_gpu_detect.py::test_nms_kernel_ties_labels_and_order" -q --tb=short 2>&1 | grep -v "^  \|array(\[" | cut -c1-400 | tail -60 > gpurun_out/r2_tests_dbg.log
cat gpurun_out/r2_tests_dbg.log
timeout 600 python tools/library_baseline.py 7 20 > gpurun_out/r2_library_baseline.json 2> gpurun_out/r2_library_baseline.err
cat gpurun_out/r2_library_baseline.json; tail -3 gpurun_out/r2_library_baseline.err
