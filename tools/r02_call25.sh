#!/bin/bash
# mbarrier.try_wait with a suspend-time hint (20 us) against the default poll: same-box A/B (library otherwise identical)
mkdir -p gpurun_out
for v in cs_str2 hint cs_str2 hint; do
  echo "== $v"
  export DDIF_LIB=gpurun_var/lib_$v.so
  python tools/layer_bench.py 256 64 64 32 32 1 1 1
  python tools/layer_bench.py 256 32 32 64 64 1 1 1
  python tools/layer_bench.py 256 8 8 128 128 1 1 1
  python tools/run_cs_gemm.py 256 64 64 64 32
  python tools/profile_step.py --batch 256 | head -2 | tail -1
  python tools/profile_step.py --batch 32 | head -2 | tail -1
done 2>&1 | grep -v "Traceback\|File \|Broken\|print(\|main()" | tee gpurun_out/r02s2_wait_hint_ab.txt
