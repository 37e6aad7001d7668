#!/bin/bash
# round 2, session 2: final-state check: parity suite, smoke, bench line (with per-launch floors and the configs[0] CPU leg), memory-bound kernels vs HBM
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02s2d_gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r02s2d_bench_B256.json 2> gpurun_out/r02s2d_bench.err; tail -c 900 gpurun_out/r02s2d_bench_B256.json; tail -3 gpurun_out/r02s2d_bench.err
timeout 300 python tools/bench_configs.py kernels 2>&1 | tee gpurun_out/r02s2d_membound_kernels.jsonl | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: print(ln[:200]); continue
    for r in d.get('rows', []): print(f\"{r['kernel']:<50}{r['us']:>9.1f} us {r['achieved_gbs']:>8.0f} GB/s  {r['frac_of_hbm_peak']:.2f}\")
"
