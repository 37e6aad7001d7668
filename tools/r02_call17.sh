#!/bin/bash
# round 2, session 2: full parity suite + bench line after the ring-depth change
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02s2b_gpu_tests.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r02s2b_bench_B256.json 2> gpurun_out/r02s2b_bench.err; tail -c 1500 gpurun_out/r02s2b_bench_B256.json; tail -3 gpurun_out/r02s2b_bench.err
timeout 300 python tools/layer_bench.py suite > gpurun_out/r02s2b_layer_suite.txt 2>&1; cat gpurun_out/r02s2b_layer_suite.txt
