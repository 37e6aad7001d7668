#!/bin/bash
# round 2, session 2: (a) idle transform warps as two more epilogue groups in the layers without GroupNorm prologue, (b) column-softmax GEMM with four
# transform pipelines (864 threads, 72 registers): tests, then same-box A/B against the previous build
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in prev cur; do
  echo "== $v"
  if [ $v = cur ]; then unset DDIF_LIB; else export DDIF_LIB=gpurun_var/lib_$v.so; fi
  for a in "256 64 64 32 64 0 0 0 1" "256 64 64 64 32 0 1 1" "256 32 32 64 128 0 0 0 1" "256 32 32 128 64 0 1 1" "256 16 16 128 256 0 0 0 1" "256 16 16 256 128 0 1 1" "256 64 64 32 32 0 0 0" "256 64 64 64 64 0 0 0" "32 64 64 32 64 0 0 0 1"; do
    python tools/layer_bench.py $a
  done
  for s in "256 64 64 64 32" "256 64 64 128 32" "256 32 32 128 64"; do python tools/run_cs_gemm.py $s; done
  python tools/profile_step.py --batch 256 | head -14
  python tools/profile_step.py --batch 32 | head -2
done 2>&1 | grep -v "Traceback\|File \|Broken\|print(\|main()" | tee gpurun_out/r02s2_extra_epi_cs4_ab.txt
# which kernel reads uninitialised global memory (initcheck reported 12 4-byte reads in tests/test_gpu_gemm.py)?
timeout 480 /usr/local/cuda/bin/compute-sanitizer --tool initcheck --error-exitcode 0 --print-limit 4 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | grep -A14 "Uninitialized" | head -80 | cut -c1-260 | tee gpurun_out/r02s2_initcheck_detail.txt
