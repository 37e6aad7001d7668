for v in "" G_NO_EPI G_NO_EM G_NO_EMT; do
  echo "== ${v:-base}"
  for a in "256 64 64 32 32 0 0 1 0 1 24 1" "256 64 64 64 32 0 1 0 0 1 24 0" "256 32 32 64 64 0 0 1 0 1 24 1" "256 8 8 128 128 0 1 1 0 1 24 0"; do
    if [ -z "$v" ]; then python tools/layer_bench.py $a; else DDIF_LIB=gpurun_var/lib_$v.so python tools/layer_bench.py $a; fi
  done
done 2>&1 | tee gpurun_out/gemm1x1_variants.txt
