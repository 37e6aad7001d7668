"""Time the column-softmax GEMM (cs_gemm_tc_kernel: FWM softmax over H fused into attn_out + bias + residual) for one shape; also usable under ncu.
usage: run_cs_gemm.py [B=256] [H=64] [W=64] [dim=64] [o=32] [reps=20]"""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from gpu_util import *
a = [int(v) for v in sys.argv[1:]] + [256, 64, 64, 64, 32, 20][len(sys.argv) - 1:]
B, H, W, dim, o, reps = a
g = torch.Generator().manual_seed(3)
qr = [(torch.randn(B, H, W, dim + o, generator=g) * 2.0).to(torch.bfloat16).to(DEV) for _ in range(2)]   # two buffers: > L2 at B = 256
weff = (torch.randn(B, (o + 15) // 16 * 16, dim, generator=g) * 0.2).to(torch.bfloat16).to(DEV)
bias = torch.randn(o, generator=g).to(DEV)
out = torch.zeros(B, H, W, o, dtype=torch.bfloat16, device=DEV)
def run(i):
    gemm([qr[i & 1]], [weff], o, taps=[1], bias=bias, per_sample=[1], w_s=[B], a_c=[dim], residual=qr[i & 1], residual_off=dim, softmax_h=True,
         reuse=(out, None), sync=False)
for i in range(3): run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps): run(i)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1000 / reps
by = B * H * W * 2 * (dim + 2 * o)
tiles = B * W // (128 // H)
print(f"cs_gemm B={B} {H}x{W} dim={dim} o={o}: {us:.1f} us per launch, {by / us / 1e3:.0f} GB/s, {us * 1965 / (tiles / 148):.0f} cyc/tile", flush=True)
