#!/bin/bash
# 2-GPU dry run of the multi-GPU measurements (weak + strong bench, training step with NCCL gradient all-reduce)
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 > gpurun_out/r02c7_bench_2gpu.json 2> gpurun_out/r02c7_bench_2gpu.err; tail -c 1200 gpurun_out/r02c7_bench_2gpu.json; tail -3 gpurun_out/r02c7_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/train_bench.py --batch 32 --steps 5 --warmup 2 2>&1 | tail -2 | tee gpurun_out/r02c7_train_bench_2gpu.json
