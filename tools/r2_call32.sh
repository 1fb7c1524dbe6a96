set -x
RMR_LIB_PATH=$PWD/tools/ab/epi.so timeout 300 python tools/profile_layers.py 7 > gpurun_out/r2_layers_epi.txt 2>&1
timeout 300 python tools/profile_layers.py 7 > gpurun_out/r2_layers_new.txt 2>&1
grep "^==" gpurun_out/r2_layers_epi.txt gpurun_out/r2_layers_new.txt
