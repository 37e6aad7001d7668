#!/bin/bash
# round 2, GPU call 1: parity suite incl. the full-horizon fixtures, bench line with all extras, layer suite + per-op step profile
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "^\[|passed|failed|Error|error|assert" | tail -80 > gpurun_out/r02c1_gpu_tests.txt; tail -5 gpurun_out/r02c1_gpu_tests.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r02c1_bench.json 2> gpurun_out/r02c1_bench.err; tail -c 1500 gpurun_out/r02c1_bench.json; tail -3 gpurun_out/r02c1_bench.err
timeout 300 python tools/layer_bench.py suite > gpurun_out/r02c1_layer_suite.txt 2>&1; cat gpurun_out/r02c1_layer_suite.txt
timeout 300 python tools/profile_step.py --batch 256 > gpurun_out/r02c1_step_profile.txt 2>&1; head -24 gpurun_out/r02c1_step_profile.txt
