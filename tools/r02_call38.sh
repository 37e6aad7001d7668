#!/bin/bash
# FINAL check of the committed tree: full parity suite, smoke, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02final_gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02final_smoke.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r02final_bench_B256.json 2> gpurun_out/r02final_bench.err; tail -c 200 gpurun_out/r02final_bench_B256.json; tail -2 gpurun_out/r02final_bench.err
timeout 200 python tools/profile_step.py --batch 256 --top 10 > gpurun_out/r02final_step_profile_B256.txt 2>&1; head -3 gpurun_out/r02final_step_profile_B256.txt
timeout 200 python tools/profile_step.py --batch 32 > gpurun_out/r02final_step_profile_B32.txt 2>&1; head -3 gpurun_out/r02final_step_profile_B32.txt
