#!/bin/bash
# round 2, session 2: fused FWM front at 8x8 (fwm_front64_kernel): unit test (short timeout), UNet parity, full suite, A/B of the step
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "fwm_front" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_headline.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
for f in "" "--no-fwm-front"; do
  echo "== $f"
  python tools/profile_step.py --batch 256 $f --top 6 | sed -n 2,25p
  python tools/profile_step.py --batch 32 $f | head -2 | tail -1
done 2>&1 | grep -v "Traceback\|File \|Broken\|print(\|main()" | tee gpurun_out/r02s2_fwm_front_ab.txt
