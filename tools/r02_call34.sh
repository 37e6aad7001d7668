#!/bin/bash
# round 2, session 2, FINAL evidence of the committed tree (after the fused FWM front): parity suite, smoke, bench line, ncu launch list of one denoise step,
# step profiles
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02final_gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02final_smoke.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r02final_bench_B256.json 2> gpurun_out/r02final_bench.err; tail -c 300 gpurun_out/r02final_bench_B256.json; tail -2 gpurun_out/r02final_bench.err
timeout 300 python tools/profile_step.py --batch 256 --top 10 > gpurun_out/r02final_step_profile_B256.txt 2>&1; head -3 gpurun_out/r02final_step_profile_B256.txt
timeout 300 python tools/profile_step.py --batch 32 > gpurun_out/r02final_step_profile_B32.txt 2>&1; head -3 gpurun_out/r02final_step_profile_B32.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_step_B256.csv python tools/profile_step.py --batch 256 --ncu > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
