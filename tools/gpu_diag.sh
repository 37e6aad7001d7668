#!/bin/bash
# First-contact GPU diagnostic: every test group in its own process (a trapped kernel poisons the CUDA context),
# each under a timeout, logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/diag.log; timeout 300 "$@" >> gpurun_out/diag.log 2>&1; echo "--- exit $?" | tee -a gpurun_out/diag.log; }
: > gpurun_out/diag.log
run gemm_all python -m pytest tests/test_gpu_gemm.py -q -s -m gpu -p no:cacheprovider
if grep -q "failed" gpurun_out/diag.log; then
  for c in c32_sw64 c64_sw128 c16_sw32 c96_k32 c128_n384 c128_3x3_8 c256_n128 s2_64 s2_32_big n512 rect; do
    run "gemm_$c" python -m pytest "tests/test_gpu_gemm.py::test_conv_matches_torch[$c]" -q -s -m gpu -p no:cacheprovider
  done
  run gemm_epi python -m pytest tests/test_gpu_gemm.py::test_epilogue_all_options -q -s -m gpu -p no:cacheprovider
  run gemm_dual python -m pytest tests/test_gpu_gemm.py::test_dual_segment_per_sample_weights -q -s -m gpu -p no:cacheprovider
fi
run kernels python -m pytest tests/test_gpu_kernels.py -q -s -m gpu -p no:cacheprovider
run unet python -m pytest tests/test_gpu_unet.py -q -s -m gpu -p no:cacheprovider
run smoke python __graft_entry__.py smoke
tail -5 gpurun_out/diag.log
grep -E "^(===|--- exit|FAILED|ERROR|[0-9]+ (passed|failed))" gpurun_out/diag.log
