# usage: run_variants.sh V1 V2 ...   (variant libraries gpurun_var/lib_<V>.so built with csrc/build.py -D... -o...)
for v in "$@"; do
  echo "== $v"
  for a in "256 64 64 32 32 1 1 1" "256 64 64 32 32 0 0 0" "256 32 32 64 64 1 1 1" "256 32 32 64 128 0 0 0 1"; do
    DDIF_LIB=gpurun_var/lib_$v.so python tools/layer_bench.py $a
  done
done 2>&1 | tee gpurun_out/layer_variants.txt
