#!/bin/bash
# round 2, session 2, call 1: re-establish the baseline on a fresh box: full parity suite, full bench line (all extras), layer suite, step profiles
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02s2_gpu_tests.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r02s2_bench_B256.json 2> gpurun_out/r02s2_bench.err; tail -c 3000 gpurun_out/r02s2_bench_B256.json; tail -3 gpurun_out/r02s2_bench.err
timeout 300 python tools/layer_bench.py suite > gpurun_out/r02s2_layer_suite.txt 2>&1; cat gpurun_out/r02s2_layer_suite.txt
timeout 300 python tools/profile_step.py --batch 256 > gpurun_out/r02s2_step_profile_B256.txt 2>&1; head -30 gpurun_out/r02s2_step_profile_B256.txt
timeout 300 python tools/profile_step.py --batch 32 > gpurun_out/r02s2_step_profile_B32.txt 2>&1; head -30 gpurun_out/r02s2_step_profile_B32.txt
