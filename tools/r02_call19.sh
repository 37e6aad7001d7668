#!/bin/bash
# round 2, session 2: fused column-softmax GEMM: kernel tests (short timeout first), full suite, A/B of the denoise step
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k "softmax_h" 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for f in "" "--no-cs-gemm"; do
  echo "== profile_step $f"
  timeout 300 python tools/profile_step.py --batch 256 $f --top 8 2>&1 | head -34
  timeout 300 python tools/profile_step.py --batch 32 $f 2>&1 | head -2
done 2>&1 | tee gpurun_out/r02s2_cs_gemm_ab.txt
