# round 2, GPU call 1: hardware limits for the conv redesign + ncu --set full of the non-conv kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_c1_smi.txt
timeout 300 tools/umma_bench > gpurun_out/r2_umma_bench.txt 2>&1
tail -60 gpurun_out/r2_umma_bench.txt
# non-conv kernels: full capture of the 2nd step's launches (skip the first step = warm-up)
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'letterbox|decode_compact|nms_restore|project_kernel|resolve_diff|scan_blocks|compact_kernel|link_kernel|flatten|collect_roots|rank_roots|label_kernel|search_kernel|conv_stem|sppf|copy_channels|upsample2' \
  -s 40 -c 40 -o gpurun_out/r2_aux python tools/step_once.py 3 > gpurun_out/r2_ncu_aux.log 2>&1
tail -5 gpurun_out/r2_ncu_aux.log
ncu -i gpurun_out/r2_aux.ncu-rep --page raw --csv > gpurun_out/r2_aux_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
