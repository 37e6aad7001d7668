#!/bin/bash
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "fwm_front" 2>&1 | tail -4
python tools/run_fwm_front.py 256 128 128 20
python tools/run_fwm_front.py 256 128 64 20
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_headline.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
python tools/profile_step.py --batch 256 | head -2 | tail -1
python tools/profile_step.py --batch 32 | head -2 | tail -1
