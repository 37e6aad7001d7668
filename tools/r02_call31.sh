#!/bin/bash
# round 2, session 2, FINAL evidence of the committed tree: parity suite, smoke, bench line, ncu launch list of one denoise step, --set full of the halo conv
# (32 -> 32 @64^2 GN + residual + stats) and of the column-softmax GEMM, step profiles
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02final_gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02final_smoke.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r02final_bench_B256.json 2> gpurun_out/r02final_bench.err; tail -c 400 gpurun_out/r02final_bench_B256.json; tail -2 gpurun_out/r02final_bench.err
timeout 300 python tools/profile_step.py --batch 256 --top 10 > gpurun_out/r02final_step_profile_B256.txt 2>&1; head -3 gpurun_out/r02final_step_profile_B256.txt
timeout 300 python tools/profile_step.py --batch 32 > gpurun_out/r02final_step_profile_B32.txt 2>&1; head -3 gpurun_out/r02final_step_profile_B32.txt
timeout 300 python tools/layer_bench.py suite > gpurun_out/r02final_layer_suite.txt 2>&1; tail -13 gpurun_out/r02final_layer_suite.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_step_B256.csv python tools/profile_step.py --batch 256 --ncu > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 4 -c 1 -f -o gpurun_out/prof_halo32 \
  python tools/layer_bench.py 256 64 64 32 32 1 1 1 0 9 6 > gpurun_out/ncu_halo32.log 2>&1; tail -1 gpurun_out/ncu_halo32.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cs_gemm -s 3 -c 1 -f -o gpurun_out/prof_cs_gemm \
  python tools/run_cs_gemm.py 256 64 64 64 32 6 > gpurun_out/ncu_cs.log 2>&1; tail -1 gpurun_out/ncu_cs.log
