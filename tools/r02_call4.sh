#!/bin/bash
# round 2, GPU call 4: new tcgen05 attention + tall softmax kernels (tests under a short timeout first), then whole-scene profile + configs
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "tcgen05 or tall or softmax" 2>&1 | tail -25
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/profile_step.py --batch 1 --size 512 --dataset gf2 --reps 3 --top 14 2>&1 | tee gpurun_out/r02c4_whole_scene_profile.txt
python tools/layer_bench.py suite 2>&1 | tee gpurun_out/r02c4_layer_suite.txt
python tools/profile_step.py --batch 256 2>&1 | head -3
timeout 300 python tools/bench_configs.py config3 2>&1 | tee gpurun_out/r02c4_config3.jsonl
