python tools/dbg_frame8.py 2>&1 | tail -30
echo ---- exp silu
RMR_SILU_EXP=1 python tools/dbg_frame8.py 2>&1 | tail -12
echo ---- v1 kernel
RMR_CONV_V2=0 python tools/dbg_frame8.py 2>&1 | tail -12
