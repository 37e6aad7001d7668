import sys, math, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from gpu_util import *
import torch.nn.functional as F
which = sys.argv[1]
B,H,W,Cin,Cout = 1,32,32,96,64
g=torch.Generator().manual_seed(1)
x=(torch.randn(B,Cin,H,W,generator=g)).to(DEV); w=(torch.randn(Cout,Cin,3,3,generator=g)*0.03).to(DEV)
xa, wp = nhwc_bf16(x), pack_w(w)
xf=to_nchw_f32(xa)
ref=F.conv2d(xf, w.to(torch.bfloat16).float(), None, padding=1)
if which.startswith('gn'): pass
elif which=='plain': out,_=gemm([xa],[wp],Cout,taps=[9])
elif which=='tma': out,_=gemm([xa],[wp],Cout,taps=[9],force_tma=1)
if not which.startswith('gn'): print(which, 'rel', rel_err(to_nchw_f32(out), ref))
if which.startswith('gn'):
    Cin2 = int(which[2:] or 96)
    x=(torch.randn(B,Cin2,H,W,generator=g)*1.5+0.4).to(DEV); w=(torch.randn(Cout,Cin2,3,3,generator=g)*0.03).to(DEV)
    xa, wp = nhwc_bf16(x), pack_w(w); xf=to_nchw_f32(xa)
    gamma=(1+0.1*torch.randn(Cin2,generator=g)).to(DEV); beta=(0.1*torch.randn(Cin2,generator=g)).to(DEV)
    stats=torch.stack([xf.double().sum(dim=(1,2,3)),(xf.double()**2).sum(dim=(1,2,3))],dim=1).contiguous()
    out,_=gemm([xa],[wp],Cout,taps=[9],gn=(stats,gamma,beta,1))
    h=F.group_norm(xf,1,gamma,beta,eps=1e-5); h=(h*torch.sigmoid(h)).to(torch.bfloat16).float()
    ref=F.conv2d(h,w.to(torch.bfloat16).float(),None,padding=1)
    print(which,'rel',rel_err(to_nchw_f32(out),ref))
