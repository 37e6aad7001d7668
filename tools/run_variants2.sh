# usage: run_variants2.sh V1 V2 ...  -- two representative ResBlock conv layers per variant library gpurun_var/lib_<V>.so
for v in "$@"; do
  echo "== $v"
  for a in "256 64 64 32 32 1 1 1" "256 32 32 64 64 1 1 1"; do
    if [ "$v" = base ]; then python tools/layer_bench.py $a; else DDIF_LIB=gpurun_var/lib_$v.so python tools/layer_bench.py $a; fi
  done
done 2>&1 | tee gpurun_out/layer_variants2.txt
