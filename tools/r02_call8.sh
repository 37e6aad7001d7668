#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/train_bench.py --batch 32 --steps 10 --warmup 3 --graph 2>&1 | tail -4 | cut -c1-1500 | tee gpurun_out/r02c8_train_bench_graph_1gpu.json
timeout 300 python tools/train_bench.py --batch 32 --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/r02c8_train_bench_eager_1gpu.json
