#!/bin/bash
# round 2 evidence: ncu launch list of one denoise step (time + DRAM bytes), --set full captures of the halo conv, the CSM 1x1 GEMM and the
# tcgen05 attention kernel
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_step_B256.csv python tools/profile_step.py --batch 256 --ncu > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 4 -c 1 -f -o gpurun_out/prof_halo32 \
  python tools/layer_bench.py 256 64 64 32 32 1 1 1 0 9 6 > gpurun_out/ncu_halo32.log 2>&1; tail -1 gpurun_out/ncu_halo32.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 4 -c 1 -f -o gpurun_out/prof_igemm_xconv \
  python tools/layer_bench.py 256 64 64 32 32 0 0 1 0 1 6 1 > gpurun_out/ncu_igemm.log 2>&1; tail -1 gpurun_out/ncu_igemm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 3 -c 1 -f -o gpurun_out/prof_attn_tc \
  python tools/run_attn_tc.py 1 64 6 > gpurun_out/ncu_attn_tc.log 2>&1; tail -1 gpurun_out/ncu_attn_tc.log
python tools/run_attn_tc.py 1 64 20; python tools/run_attn_tc.py 1 32 20; python tools/run_attn_tc.py 8 32 20
ls -la gpurun_out/*.ncu-rep
