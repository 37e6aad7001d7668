#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_unet.py -m gpu -x -q 2>&1 | tail -3
for v in noresring cur; do
  echo "== $v"
  if [ $v = cur ]; then unset DDIF_LIB; else export DDIF_LIB=gpurun_var/lib_$v.so; fi
  for a in "256 32 32 64 64 1 1 1" "256 16 16 64 64 1 1 1" "256 64 64 64 64 1 1 1" "256 32 32 64 64 0 1 1" "32 32 32 64 64 1 1 1"; do
    python tools/layer_bench.py $a
  done
  python tools/profile_step.py --batch 256 | head -5
  python tools/profile_step.py --batch 32 | head -2
done 2>&1 | grep -v "Traceback\|File \|Broken\|print(\|main()" | tee gpurun_out/r02c12_resring_ab.txt
