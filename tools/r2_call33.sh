set -x
for v in epi walk; do
RMR_LIB_PATH=$PWD/tools/ab/$v.so timeout 300 python tools/profile_layers.py 7 > gpurun_out/r2_layers_$v.txt 2>&1
done
timeout 300 python tools/profile_layers.py 7 > gpurun_out/r2_layers_new.txt 2>&1
grep "^==" gpurun_out/r2_layers_epi.txt gpurun_out/r2_layers_walk.txt gpurun_out/r2_layers_new.txt
