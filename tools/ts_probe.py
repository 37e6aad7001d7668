"""Pipeline timestamps of CTA 0 of the 3x3 conv kernels (debug hook) for one layer shape.
usage: ts_probe.py B H W Cin Cout [gn=1] [residual=0]"""
import sys, ctypes, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from gpu_util import *
from dif_pan_b200 import _lib
B,H,W,Cin,Cout = [int(v) for v in sys.argv[1:6]]
gnflag = int(sys.argv[6]) if len(sys.argv) > 6 else 1
resflag = int(sys.argv[7]) if len(sys.argv) > 7 else 0
stflag = int(sys.argv[8]) if len(sys.argv) > 8 else 1
nrows = int(sys.argv[9]) if len(sys.argv) > 9 else 14
row0 = int(sys.argv[10]) if len(sys.argv) > 10 else 0
g=torch.Generator().manual_seed(1)
x=(torch.randn(B,Cin,H,W,generator=g)).to(DEV); w=(torch.randn(Cout,Cin,3,3,generator=g)*0.03).to(DEV)
bias=torch.zeros(Cout,device=DEV)
xa, wp = nhwc_bf16(x), pack_w(w); xf=to_nchw_f32(xa)
gamma=torch.ones(Cin,device=DEV); beta=torch.zeros(Cin,device=DEV)
stats=torch.stack([xf.double().sum(dim=(1,2,3)),(xf.double()**2).sum(dim=(1,2,3))],dim=1).contiguous()
lib=_lib.load()
ts=torch.zeros(4*64*4,dtype=torch.int64,device=DEV)
kw=dict(gn=(stats,gamma,beta,1)) if gnflag else {}
if resflag: kw['residual']=nhwc_bf16(torch.randn(B,Cout,H,W,generator=g).to(DEV))
gemm([xa],[wp],Cout,taps=[9],bias=bias,want_stats=bool(stflag),**kw)
lib.ddif_debug_set_timestamps(ctypes.c_void_p(ts.data_ptr()))
gemm([xa],[wp],Cout,taps=[9],bias=bias,want_stats=bool(stflag),**kw)
lib.ddif_debug_set_timestamps(None)
t=ts.cpu().view(4,64,4)
t0=int(t[t>0].min())
r=lambda v: (int(v)-t0) if v>0 else -1
print(f"B={B} {H}x{W} {Cin}->{Cout} gn={gnflag} res={resflag}")
print("tile | producer: pre-wait issue | xform: pre-wait landed stored arrived | mma: pre-empty got-empty got-a_full committed | epi: got-full tmem-loaded processed done")
for i in range(row0, row0 + nrows):
    print(i, '|', [r(v) for v in t[0,i][:2]], '|', [r(v) for v in t[3,i]], '|', [r(v) for v in t[1,i]], '|', [r(v) for v in t[2,i]])
n=min(40, int((t[1,:,3]>0).sum()))
d=[int(t[1,i+1,3]-t[1,i,3]) for i in range(4,n-1)]
print('cycles per tile (mma commit to commit), tiles 4..', n, ':', sum(d)/max(len(d),1))
