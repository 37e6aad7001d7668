#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
python tools/run_fwm_front.py 256 128 128 20
python tools/run_fwm_front.py 32 128 128 20
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fwm_front64 -s 3 -c 1 -f -o gpurun_out/prof_fwm_front \
  python tools/run_fwm_front.py 256 128 128 6 > gpurun_out/ncu_ff.log 2>&1; tail -1 gpurun_out/ncu_ff.log
