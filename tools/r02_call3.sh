#!/bin/bash
# round 2, GPU call 3: same-box A/B of the halo-conv changes (round-1 library vs relay variant vs current) + whole-scene per-op profile
mkdir -p gpurun_out
for v in r01 relay cur; do
  echo "== $v"
  if [ $v = cur ]; then unset DDIF_LIB; else export DDIF_LIB=gpurun_var/lib_$v.so; fi
  for a in "256 64 64 32 32 1 1 1" "256 64 64 32 32 1 0 1" "256 32 32 64 64 1 1 1" "256 32 32 64 64 1 0 1" "256 64 64 64 32 0 1 1" "256 32 32 128 64 0 1 1" "32 64 64 32 32 1 1 1" "256 64 64 32 32 1 1 1 0 1 24 1" "256 64 64 64 32 0 1 1 0 1"; do
    python tools/layer_bench.py $a
  done
  python tools/profile_step.py --batch 256 | head -2
  python tools/profile_step.py --batch 32 | head -2
done 2>&1 | tee gpurun_out/r02c3_ab.txt
unset DDIF_LIB
python tools/profile_step.py --batch 1 --size 512 --dataset gf2 --reps 3 --top 30 2>&1 | tee gpurun_out/r02c3_whole_scene_profile.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
