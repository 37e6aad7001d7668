"""Run the tcgen05 flash-attention core (csrc/attn_tc.cu) a few times (CUDA-event time; for ncu captures).
usage: run_attn_tc.py [B=1] [hw=64] [reps=20]      (hw x hw tokens: 64 -> 4 096 tokens = the attention level of a 512 x 512 scene)"""
import math, sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from gpu_util import *
from dif_pan_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
C, heads = 128, 8
g = torch.Generator().manual_seed(3)
qkv = nhwc_bf16((torch.randn(B, 3 * C, hw, hw, generator=g) * 1.5).to(DEV))
out = torch.zeros(B, hw, hw, C, dtype=torch.bfloat16, device=DEV)
def run():
    _lib.launch("ddif_attn_t", stream(), qkv=qkv.data_ptr(), out=out.data_ptr(), batch=B, ntok=hw * hw, c=C, heads=heads, scale=1 / math.sqrt(C))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1000 / reps
fl = 4.0 * B * (hw * hw) ** 2 * C
print(f"attn_tc B={B} ntok={hw * hw}: {us:.1f} us per launch, {fl / us / 1e6:.1f} TFLOP/s (4 n^2 C), {B * hw * hw * heads * (hw * hw) / us / 1e3:.1f} G exp/s")
