#!/bin/bash
# round 2, session 2, call 3: sweep of ring depth / K slab / L2 prefetch distance of the halo conv (tuning build with env knobs)
mkdir -p gpurun_out
export DDIF_LIB=gpurun_var/lib_tune.so
python tools/halo_sweep.py 2>&1 | tee gpurun_out/r02s2_halo_sweep.txt
for cfg in "8 64 0" "16 64 0" "16 32 0" "16 64 4" "16 32 4" "16 32 8"; do
  set -- $cfg
  echo "== stages $1 kslab $2 pf $3"
  DDIF_HALO_STAGES=$1 DDIF_HALO_KSLAB=$2 DDIF_HALO_PF=$3 python tools/profile_step.py --batch 256 | head -3 | tail -2
  DDIF_HALO_STAGES=$1 DDIF_HALO_KSLAB=$2 DDIF_HALO_PF=$3 python tools/profile_step.py --batch 32 | head -2 | tail -1
done 2>&1 | grep -v "Traceback\|File \|Broken\|print(\|main()" | tee gpurun_out/r02s2_halo_sweep_step.txt
