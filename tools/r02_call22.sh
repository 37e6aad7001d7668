#!/bin/bash
# round 2, session 2: compute-sanitizer over the tensor-core kernels: memcheck on the GEMM / halo conv / column-softmax GEMM tests and
# the UNet parity tests, initcheck on the GEMM tests, racecheck on the column-softmax GEMM and a halo conv case
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
{
echo "== memcheck tests/test_gpu_gemm.py"
timeout 1200 $S --tool memcheck --error-exitcode 0 --print-limit 20 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds|misaligned|Error" | tail -12
echo "== memcheck tests/test_gpu_unet.py"
timeout 1500 $S --tool memcheck --error-exitcode 0 --print-limit 20 python -m pytest tests/test_gpu_unet.py -m gpu -q -x 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds|misaligned|Error" | tail -12
echo "== initcheck tests/test_gpu_gemm.py"
timeout 1200 $S --tool initcheck --error-exitcode 0 --print-limit 20 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Uninitialized|Error" | tail -12
echo "== racecheck softmax_h GEMM + fused GN conv"
timeout 1200 $S --tool racecheck --error-exitcode 0 --print-limit 20 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x -k "softmax_h_fused or fused_gn_swish" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|Error" | tail -12
} 2>&1 | tee gpurun_out/r02s2_sanitizer.txt
