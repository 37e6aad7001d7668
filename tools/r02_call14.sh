#!/bin/bash
# round 2, session 2, call 2: halo conv ring depth (16 stages) / one TMA producer per ring, same-box A/B; steady-state timelines of the current kernel
mkdir -p gpurun_out
for v in cur st16 2prod 2prod_st16; do
  echo "== $v"
  if [ $v = cur ]; then unset DDIF_LIB; else export DDIF_LIB=gpurun_var/lib_$v.so; fi
  for a in "256 64 64 32 32 1 1 1" "256 64 64 32 32 1 0 1" "256 64 64 32 32 0 0 0" "256 32 32 64 64 1 1 1" "256 32 32 64 64 1 0 1" "256 64 64 32 64 0 0 0 1" "256 64 64 64 32 0 1 1" "256 32 32 128 64 0 1 1" "32 64 64 32 32 1 1 1" "32 32 32 64 64 1 1 1"; do
    python tools/layer_bench.py $a
  done
  python tools/profile_step.py --batch 256 | head -3
  python tools/profile_step.py --batch 32 | head -2
done 2>&1 | grep -v "Traceback\|File \|Broken\|print(\|main()" | tee gpurun_out/r02s2_ring_ab.txt
unset DDIF_LIB
python tools/ts_probe.py 256 64 64 32 32 1 1 1 16 24 2>&1 | tee gpurun_out/r02s2_timeline_32_steady.txt
python tools/ts_probe.py 256 32 32 64 64 1 1 1 16 4 2>&1 | tee gpurun_out/r02s2_timeline_64_steady.txt
