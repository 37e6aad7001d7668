#!/bin/bash
# round 2, session 2: time embedding as a side branch of the step graph (A/B), cond prep with 32 x 64 tiles, full suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02s2e_gpu_tests.txt
for b in 256 32 1; do
  for f in "" "--no-side-branch"; do
    echo "B=$b $f: $(python tools/profile_step.py --batch $b $f | head -2 | tail -1)"
  done
done 2>&1 | tee gpurun_out/r02s2e_side_branch_ab.txt
timeout 300 python tools/bench_configs.py kernels 2>&1 | tee gpurun_out/r02s2e_membound_kernels.jsonl | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: print(ln[:200]); continue
    for r in d.get('rows', []): print(f\"{r['kernel']:<50}{r['us']:>9.1f} us {r['achieved_gbs']:>8.0f} GB/s  {r['frac_of_hbm_peak']:.2f}\")
" | head -3
