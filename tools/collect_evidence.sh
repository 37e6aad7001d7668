#!/bin/bash
# Round evidence on one B200: GPU tests, bench line, ncu launch list (+ DRAM bytes) of one denoise step, full ncu capture of
# the dominant kernel.  Everything lands in gpurun_out/; tools/summarise_ncu.py turns it into profiles/ files.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/gpu_tests.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_step_B256.csv python tools/profile_step.py --batch 256 --ncu > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 4 -c 1 -f -o gpurun_out/prof_halo32 \
  python tools/layer_bench.py 256 64 64 32 32 1 1 1 0 9 6 > gpurun_out/ncu_halo32.log 2>&1
tail -2 gpurun_out/ncu_halo32.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 4 -c 1 -f -o gpurun_out/prof_halo64 \
  python tools/layer_bench.py 256 32 32 64 64 1 1 1 0 9 6 > gpurun_out/ncu_halo64.log 2>&1
tail -2 gpurun_out/ncu_halo64.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_block64 -s 3 -c 1 -f -o gpurun_out/prof_attn_block \
  python tools/run_attn_block.py 256 3 > gpurun_out/ncu_attn_block.log 2>&1
tail -2 gpurun_out/ncu_attn_block.log
timeout 600 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; tail -2 gpurun_out/configs.err
python tools/profile_step.py --batch 256 > gpurun_out/step_profile.txt 2>&1; head -3 gpurun_out/step_profile.txt
