#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_variants.py -m gpu -x -q -k "make_cond or scene or tile" 2>&1 | tail -3
timeout 300 python tools/bench_configs.py kernels 2>&1 | tee gpurun_out/r02s2d_membound_kernels.jsonl | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: print(ln[:200]); continue
    for r in d.get('rows', []): print(f\"{r['kernel']:<50}{r['us']:>9.1f} us {r['achieved_gbs']:>8.0f} GB/s  {r['frac_of_hbm_peak']:.2f}\")
" | head -3
