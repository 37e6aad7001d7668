#!/bin/bash
# column-softmax GEMM: grid-strided tile walk, 2 vs 4 transform pipelines (prev = contiguous walk, 2 pipelines: 75.1 / 123.3 / 43.7 us)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -2
for v in cs_str2 cur; do
  echo "== $v"
  if [ $v = cur ]; then unset DDIF_LIB; else export DDIF_LIB=gpurun_var/lib_$v.so; fi
  for s in "256 64 64 64 32" "256 64 64 128 32" "256 32 32 128 64"; do python tools/run_cs_gemm.py $s; done
  python tools/layer_bench.py 256 32 32 128 64 0 1 1
  python tools/layer_bench.py 256 64 64 32 64 0 0 0 1
  python tools/profile_step.py --batch 256 | head -9
  python tools/profile_step.py --batch 32 | head -2
done 2>&1 | grep -v "Traceback\|File \|Broken\|print(\|main()" | tee gpurun_out/r02s2_cs_walk_ab.txt
unset DDIF_LIB
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
