"""Run the fused FWM front at 8x8 (csrc/fwm_front.cu) a few times for one shape (CUDA-event time; also the target of ncu captures).
usage: run_fwm_front.py [B=256] [c1=128] [c2=128] [reps=20]"""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from gpu_util import DEV, stream
from dif_pan_b200 import _lib
a = [int(v) for v in sys.argv[1:]] + [256, 128, 128, 20][len(sys.argv) - 1:]
B, c1, c2, reps = a
dim, o, H = c1 + c2, 128, 8
g = torch.Generator().manual_seed(5)
x = (torch.randn(B, H, H, c1, generator=g) * 1.5).to(torch.bfloat16).to(DEV)
sk = (torch.randn(B, H, H, c2, generator=g) * 0.7).to(torch.bfloat16).to(DEV)
st = lambda v: torch.stack([v.double().sum(dim=(1, 2, 3)), (v.double() ** 2).sum(dim=(1, 2, 3))], dim=1).contiguous()
s1, s2 = st(x), st(sk)
gamma, beta = torch.ones(dim, device=DEV), torch.zeros(dim, device=DEV)
dw9 = (torch.randn(9, dim, generator=g) * 0.3).to(DEV)
w1 = (torch.randn(dim, dim, generator=g) / dim ** 0.5).to(torch.bfloat16).to(DEV)
b1 = torch.zeros(dim, device=DEV)
weff = (torch.randn(B, o, dim, generator=g) * 0.2).to(torch.bfloat16).to(DEV)
wres = (torch.randn(o, dim, generator=g) / dim ** 0.5).to(torch.bfloat16).to(DEV)
bias = torch.zeros(o, device=DEV)
out = torch.zeros(B, H, H, o, dtype=torch.bfloat16, device=DEV)
p = _lib.make("ddif_fwm_front_t", x=x.data_ptr(), skip=sk.data_ptr(), c1=c1, c2=c2, stats1=s1.data_ptr(), stats2=s2.data_ptr(), gamma=gamma.data_ptr(),
              beta=beta.data_ptr(), eps=1e-5, dw_w=dw9.data_ptr(), w1=w1.data_ptr(), w1_ld=dim, b1=b1.data_ptr(), weff=weff.data_ptr(), weff_ld=dim,
              weff_rows=o, wres=wres.data_ptr(), wres_ld=dim, bias=bias.data_ptr(), out=out.data_ptr(), out_ld=o, batch=B, h=H, w=H, o=o)
run = lambda: _lib.launch_kind("DDIF_OP_FWM_FRONT", p, stream())
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1000 / reps
fl = 2.0 * B * 64 * dim * (9 + dim + 2 * o)
print(f"fwm_front B={B} dim={dim}: {us:.1f} us per launch, {fl / us / 1e6:.1f} TFLOP/s", flush=True)
