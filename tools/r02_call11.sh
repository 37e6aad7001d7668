#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_unet.py -m gpu -x -q 2>&1 | tail -3
for v in nooptma cur; do
  echo "== $v"
  if [ $v = cur ]; then unset DDIF_LIB; else export DDIF_LIB=gpurun_var/lib_$v.so; fi
  for a in "256 64 64 32 32 0 0 1 0 1 24 1" "256 32 32 64 64 0 0 1 0 1 24 1" "256 16 16 64 64 0 0 1 0 1 24 1" "256 64 64 64 32 0 1 1 0 1" "256 32 32 128 64 0 1 1 0 1" "256 64 64 32 32 0 1 1 0 1"; do
    python tools/layer_bench.py $a
  done
  python tools/profile_step.py --batch 256 | head -12
  python tools/profile_step.py --batch 32 | head -2
done 2>&1 | grep -v "Traceback\|File \|Broken\|print(\|main()" | tee gpurun_out/r02c11_optma_ab.txt
