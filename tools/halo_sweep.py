"""Sweep of the halo conv's ring depth / K slab / L2-prefetch distance over the UNet's layer shapes in ONE process (tuning build:
csrc/build.py -DDDIF_VAR_HALO_TUNE_ENV -DDDIF_VAR_HALO_MAX_STAGES=16 -ogpurun_var/lib_tune.so; DDIF_LIB=gpurun_var/lib_tune.so)."""
import os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
from layer_bench import bench

SHAPES = [(256, 64, 64, 32, 32, 1, 1, 1, 0), (256, 64, 64, 32, 32, 1, 0, 1, 0), (256, 32, 32, 64, 64, 1, 1, 1, 0), (256, 32, 32, 64, 64, 1, 0, 1, 0),
          (256, 64, 64, 32, 64, 0, 0, 0, 1), (256, 64, 64, 64, 32, 0, 1, 1, 0), (256, 32, 32, 64, 128, 0, 0, 0, 1), (256, 32, 32, 128, 64, 0, 1, 1, 0),
          (256, 16, 16, 128, 128, 1, 1, 1, 0), (256, 8, 8, 128, 128, 1, 1, 1, 0), (256, 16, 16, 128, 256, 0, 0, 0, 1), (256, 16, 16, 256, 128, 0, 1, 1, 0),
          (256, 64, 64, 64, 64, 1, 0, 0, 0), (32, 64, 64, 32, 32, 1, 1, 1, 0)]
SETTINGS = [(8, 64, 0), (12, 64, 0), (16, 64, 0), (8, 32, 0), (16, 32, 0), (8, 64, 4), (16, 64, 4), (16, 32, 4), (16, 64, 8), (16, 32, 8)]
if len(sys.argv) > 1:
    SETTINGS = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
res = {}
for st, ks, pf in SETTINGS:
    os.environ.update(DDIF_HALO_STAGES=str(st), DDIF_HALO_KSLAB=str(ks), DDIF_HALO_PF=str(pf))
    for s in SHAPES:
        res[(st, ks, pf), s] = bench(*s[:5], gn=s[5], res=s[6], st=s[7], act=s[8], quiet=True)
print("us/launch; columns = (max stages, max K slab, L2 prefetch distance in tiles)")
print(f"{'shape (B H W Cin Cout gn res stats act)':<44}" + "".join(f"{str(k):>13}" for k in SETTINGS))
for s in SHAPES:
    print(f"{str(s):<44}" + "".join(f"{res[k, s]:>13.1f}" for k in SETTINGS), flush=True)
