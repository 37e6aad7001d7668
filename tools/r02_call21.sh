#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02s2c_gpu_tests.txt
timeout 300 python tools/profile_step.py --batch 256 --top 10 > gpurun_out/r02s2c_step_profile_B256.txt 2>&1; head -32 gpurun_out/r02s2c_step_profile_B256.txt
cp gpurun_out/step_profile_B256_64.json gpurun_out/r02s2c_step_profile_B256.json
timeout 300 python tools/profile_step.py --batch 32 > gpurun_out/r02s2c_step_profile_B32.txt 2>&1; head -3 gpurun_out/r02s2c_step_profile_B32.txt
