#!/bin/bash
# compute-sanitizer on the kernels added after the first sanitizer pass: fused FWM front (memcheck + racecheck: in-place q overwrite, cp.async weight
# ring), fused cond prep with the shared-memory coefficient tile (memcheck)
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
{
echo "== memcheck: fwm_front + make_cond"
timeout 400 $S --tool memcheck --error-exitcode 0 --print-limit 10 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_variants.py -m gpu -q -x -k "fwm_front or make_cond" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds|misaligned" | tail -6
echo "== racecheck: fwm_front"
timeout 400 $S --tool racecheck --error-exitcode 0 --print-limit 10 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fwm_front" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|Race reported" | cut -c1-200 | tail -8
} 2>&1 | tee gpurun_out/r02s2_sanitizer2.txt
