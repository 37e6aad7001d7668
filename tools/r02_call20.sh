#!/bin/bash
# round 2, session 2 evidence: bench line, ncu launch list of one denoise step (time + DRAM bytes), --set full captures of the halo conv and of the
# column-softmax GEMM, layer times of the column-softmax GEMM
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r02s2c_bench_B256.json 2> gpurun_out/r02s2c_bench.err; tail -c 600 gpurun_out/r02s2c_bench_B256.json; tail -3 gpurun_out/r02s2c_bench.err
for s in "256 64 64 64 32" "256 64 64 96 32" "256 32 32 128 64" "256 32 32 96 64" "256 16 16 128 64" "256 16 16 256 128" "32 64 64 64 32"; do python tools/run_cs_gemm.py $s; done 2>&1 | tee gpurun_out/r02s2c_cs_gemm_layers.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_step_B256.csv python tools/profile_step.py --batch 256 --ncu > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 4 -c 1 -f -o gpurun_out/prof_halo32 \
  python tools/layer_bench.py 256 64 64 32 32 1 1 1 0 9 6 > gpurun_out/ncu_halo32.log 2>&1; tail -1 gpurun_out/ncu_halo32.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cs_gemm -s 3 -c 1 -f -o gpurun_out/prof_cs_gemm \
  python tools/run_cs_gemm.py 256 64 64 64 32 6 > gpurun_out/ncu_cs.log 2>&1; tail -1 gpurun_out/ncu_cs.log
ls -la gpurun_out/*.ncu-rep
