#!/bin/bash
# probe build (DDIF_VAR_TS2): where does the transform group spend the ~1000 cycles between two stages?  xform stamps = [loop top, tables ready, landed, arrived]
mkdir -p gpurun_out
export DDIF_LIB=gpurun_var/lib_ts2.so
python tools/ts_probe.py 256 8 8 128 128 1 1 1 10 0 2>&1 | tee gpurun_out/r02s2_ts2_8x8_128.txt
python tools/ts_probe.py 256 64 64 32 32 1 0 1 10 20 2>&1 | tee gpurun_out/r02s2_ts2_64x64_32_nores.txt
python tools/ts_probe.py 256 32 32 64 64 1 0 1 8 4 2>&1 | tee gpurun_out/r02s2_ts2_32x32_64_nores.txt
