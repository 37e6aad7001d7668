#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "fwm_front" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_headline.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
python tools/profile_step.py --batch 256 --top 4 | sed -n 2,14p
python tools/profile_step.py --batch 32 | head -2 | tail -1
