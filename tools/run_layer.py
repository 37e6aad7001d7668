"""Run one 3x3 conv layer shape through DDIF_OP_GEMM a few times and print its CUDA-event time (for ncu captures).
usage: run_layer.py B H W Cin Cout [gn=1] [residual=0] [stats=1] [reps=5]"""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from gpu_util import *
a = [int(v) for v in sys.argv[1:]]
B, H, W, Cin, Cout = a[:5]
gnflag = a[5] if len(a) > 5 else 1
resflag = a[6] if len(a) > 6 else 0
stflag = a[7] if len(a) > 7 else 1
reps = a[8] if len(a) > 8 else 5
g = torch.Generator().manual_seed(1)
x = torch.randn(B, Cin, H, W, generator=g).to(DEV)
w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.03).to(DEV)
bias = torch.zeros(Cout, device=DEV)
xa, wp = nhwc_bf16(x), pack_w(w)
xf = to_nchw_f32(xa)
gamma = torch.ones(Cin, device=DEV); beta = torch.zeros(Cin, device=DEV)
stats = torch.stack([xf.double().sum(dim=(1, 2, 3)), (xf.double() ** 2).sum(dim=(1, 2, 3))], dim=1).contiguous()
del x, xf
kw = dict(gn=(stats, gamma, beta, 1)) if gnflag else {}
if resflag:
    kw['residual'] = nhwc_bf16(torch.randn(B, Cout, H, W, generator=g).to(DEV))
res = gemm([xa], [wp], Cout, taps=[9], bias=bias, want_stats=bool(stflag), **kw)
ts = []
for i in range(reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    gemm([xa], [wp], Cout, taps=[9], bias=bias, want_stats=bool(stflag), reuse=res, sync=False, **kw)
    e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e) * 1000)
tiles = B * ((H + 15) // 16) * ((W + 7) // 8)
cyc = min(ts) * 1965.0 / (tiles / 148.0)
print(f"B={B} {H}x{W} {Cin}->{Cout} gn={gnflag} res={resflag} stats={stflag}: us per launch {['%.1f' % t for t in ts]}  ~{cyc:.0f} cycles/tile @1965MHz, {2.0*B*H*W*Cin*9*Cout/min(ts)/1e6:.0f} TFLOP/s")
