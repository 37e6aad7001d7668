#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k fwm_context 2>&1 | tail -30
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -s -k "backward or autocast" 2>&1 | grep -v "^\[.*\] grad " | tail -60 | cut -c1-200 | tee gpurun_out/r02c6_training_tests.txt
