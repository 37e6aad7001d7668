#!/bin/bash
# round 2, session 2, 8-GPU box: weak + strong bench at N = 8 and 2, training step (CUDA-graph replay + eager) with NCCL gradient all-reduce at N = 8 and 1
mkdir -p gpurun_out
for n in 8 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 1 --warmup 3 > gpurun_out/r02s2_bench_${n}gpu.json 2> gpurun_out/r02s2_bench_${n}gpu.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02s2_bench_${n}gpu.json").read().strip().splitlines()[-1])
print("N=$n weak value", d["value"], "e2e", d["e2e"]["value"], "ms/denoise step", d["unet_ms_per_denoise_step"], "strong", d["strong"])
PY
  tail -2 gpurun_out/r02s2_bench_${n}gpu.err | cut -c1-300
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/train_bench.py --batch 32 --steps 10 --warmup 3 --graph 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/r02s2_train_bench_graph_8gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/train_bench.py --batch 32 --steps 5 --warmup 2 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/r02s2_train_bench_eager_8gpu.json
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/train_bench.py --batch 32 --steps 10 --warmup 3 --graph 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/r02s2_train_bench_graph_1gpu.json
