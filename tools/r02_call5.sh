#!/bin/bash
# round 2, GPU call 5: training path (wgrad / colsum kernels, conv2d Function, gradient parity vs the reference's autograd), full suite, training step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -x -q -s 2>&1 | grep -v "^\[.*grad .*rel err 0.00\|^\[.*grad .*rel err 0.01[0-4]" | tail -70 | cut -c1-220 | tee gpurun_out/r02c5_training_tests.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python tools/train_bench.py --batch 32 --steps 5 --warmup 2 2>&1 | tail -2 | tee gpurun_out/r02c5_train_bench_1gpu.json
