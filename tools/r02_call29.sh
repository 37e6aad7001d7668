#!/bin/bash
# round 2, session 2, 8-GPU box, FINAL build: weak + strong bench at N = 8, 4, 2 (the driver's SCALE run repeats this at round end)
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 1 --warmup 3 > gpurun_out/r02s2f_bench_${n}gpu.json 2> gpurun_out/r02s2f_bench_${n}gpu.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02s2f_bench_${n}gpu.json").read().strip().splitlines()[-1])
print("N=$n weak value", d["value"], "e2e", d["e2e"]["value"], "ms/denoise step", d["unet_ms_per_denoise_step"], "strong", d["strong"])
PY
done
