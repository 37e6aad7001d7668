"""Run the fused attention block at B=256 a few times (CUDA-event time; for ncu captures).  usage: run_attn_block.py [B=256] [reps=20]"""
import math, sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from gpu_util import *
from dif_pan_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
C, heads = 128, 8
g = torch.Generator().manual_seed(17)
x = torch.randn(B, C, 8, 8, generator=g)
xa = nhwc_bf16(x.to(DEV)); xf = to_nchw_f32(xa)
gamma, beta = (torch.rand(C, generator=g) + 0.5).to(DEV), (torch.randn(C, generator=g) * 0.2).to(DEV)
wqkv = (torch.randn(3 * C, C, 1, 1, generator=g) / math.sqrt(C)).to(DEV)
wout = (torch.randn(C, C, generator=g) / math.sqrt(C)).to(torch.bfloat16).to(DEV)
bout = torch.zeros(C, device=DEV)
stats_in = torch.stack([xf.double().sum(dim=(1, 2, 3)), (xf.double() ** 2).sum(dim=(1, 2, 3))], dim=1).contiguous()
wq_p = pack_w(wqkv)
out = torch.zeros(B, 8, 8, C, dtype=torch.bfloat16, device=DEV)
stats_out = torch.zeros(B, 2, dtype=torch.float64, device=DEV)
def run():
    _lib.launch("ddif_attn_block_t", stream(), x=xa.data_ptr(), stats_in=stats_in.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr(),
                wqkv=wq_p.data_ptr(), wout=wout.data_ptr(), bout=bout.data_ptr(), out=out.data_ptr(), stats_out=stats_out.data_ptr(), batch=B,
                ntok=64, c=C, heads=heads, scale=1.0 / math.sqrt(C), eps=1e-5)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
print(f"attn_block B={B}: {e0.elapsed_time(e1) * 1000 / reps:.1f} us per launch (back-to-back, includes launch gaps)")
