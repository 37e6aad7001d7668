#!/bin/bash
mkdir -p gpurun_out
python tools/ts_probe.py 256 8 8 128 128 1 1 1 8 0 2>&1 | tee gpurun_out/r02s2_timeline_8x8_128.txt
python tools/ts_probe.py 256 16 16 128 128 1 1 1 8 0 2>&1 | tee gpurun_out/r02s2_timeline_16x16_128.txt
python tools/ts_probe.py 32 8 8 128 128 1 1 1 8 0 2>&1 | tee gpurun_out/r02s2_timeline_8x8_128_B32.txt
python tools/layer_bench.py 256 8 8 128 128 1 1 1
python tools/layer_bench.py 256 8 8 128 128 0 0 0
python tools/layer_bench.py 256 8 8 128 64 0 0 0
python tools/layer_bench.py 256 8 8 64 64 0 0 0
python tools/layer_bench.py 256 8 8 32 32 0 0 0
python tools/layer_bench.py 32 8 8 128 128 1 1 1
python tools/layer_bench.py 32 8 8 32 32 0 0 0
python tools/layer_bench.py 8 8 8 32 32 0 0 0
