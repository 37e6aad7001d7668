"""Helpers shared by the `-m gpu` tests: NHWC bf16 tensors, op launches through the C ABI, error metrics."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from dif_pan_b200 import _lib

DEV = "cuda:0"


def stream() -> int:
    return torch.cuda.current_stream(torch.device(DEV)).cuda_stream


def nhwc_bf16(x_nchw: torch.Tensor, c_pad: int = None) -> torch.Tensor:
    """fp32 NCHW -> contiguous NHWC bf16 (optionally zero-padded channels)."""
    x = x_nchw.permute(0, 2, 3, 1)
    if c_pad is not None and c_pad > x.shape[-1]:
        x = F.pad(x, (0, c_pad - x.shape[-1]))
    return x.contiguous().to(torch.bfloat16)


def to_nchw_f32(x_nhwc: torch.Tensor) -> torch.Tensor:
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


def pack_w(w_oihw: torch.Tensor, cin_pad: int = None) -> torch.Tensor:
    from dif_pan_b200.unet import _pack_conv
    return _pack_conv(w_oihw, cin_pad)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def gemm(srcs, weights, n_valid, *, taps, stride=1, bias=None, film=None, mod=None, residual=None, act=0,
         per_sample=None, want_nchw=False, want_stats=False, out_hw=None, gn=None, a_up=0, force_tma=0, reuse=None, sync=True, dw=None,
         w_s=None, softmax_h=False, a_c=None, residual_off=0):
    """Run DDIF_OP_GEMM on NHWC bf16 tensors; returns (out_nhwc_bf16 | out_nchw_f32, stats | None).
    gn = (stats[B,2] f64, gamma, beta, act) fuses GroupNorm(+Swish) of the source into the 3x3 kernel."""
    a0 = srcs[0]
    B, H, W, _ = a0.shape
    oh, ow = out_hw if out_hw else ((H << a_up) // stride, (W << a_up) // stride)
    nseg = len(srcs)
    n_pad = (n_valid + 15) // 16 * 16
    if reuse is not None:  # (out | out_nchw, stats) of an earlier call: no allocation, for timing loops
        out, out_nchw, stats = (None, reuse[0], reuse[1]) if want_nchw else (reuse[0], None, reuse[1])
    else:
        out = torch.zeros(B, oh, ow, n_valid if n_valid % 8 == 0 else n_pad, dtype=torch.bfloat16, device=DEV) if not want_nchw else None
        out_nchw = torch.zeros(B, n_valid, oh, ow, dtype=torch.float32, device=DEV) if want_nchw else None
        stats = torch.zeros(B, 2, dtype=torch.float64, device=DEV) if want_stats else None
    ps = list(per_sample) if per_sample else [0] * nseg
    pad2 = lambda lst, fill: list(lst) + [fill] * (2 - nseg)
    _lib.launch(
        "ddif_gemm_t", stream(),
        a=pad2([s.data_ptr() for s in srcs], None), a_ld=pad2([s.shape[3] for s in srcs], 0), a_c=pad2(list(a_c) if a_c else [s.shape[3] for s in srcs], 0),
        a_h=pad2([s.shape[1] for s in srcs], 0), a_w=pad2([s.shape[2] for s in srcs], 0), w=pad2([w.data_ptr() for w in weights], None),
        w_s=pad2(list(w_s) if w_s else [w.shape[0] for w in weights], 0), w_k=pad2([w.shape[2] for w in weights], 0), taps=pad2(taps, 0),
        w_per_sample=pad2(ps, 0), nseg=nseg, stride=stride, batch=B, out_h=oh, out_w=ow, n_pad=n_pad, n_valid=n_valid,
        bias=bias.data_ptr() if bias is not None else None, film=film.data_ptr() if film is not None else None,
        film_ld=film.shape[1] if film is not None else 0, mod=mod.data_ptr() if mod is not None else None,
        residual=residual.data_ptr() + 2 * residual_off if residual is not None else None, res_ld=residual.shape[3] if residual is not None else 0,
        act=act, out=out.data_ptr() if out is not None else None, out_ld=out.shape[3] if out is not None else 0,
        out_nchw=out_nchw.data_ptr() if out_nchw is not None else None, stats=stats.data_ptr() if stats is not None else None,
        gn_stats=gn[0].data_ptr() if gn else None, gn_gamma=gn[1].data_ptr() if gn else None, gn_beta=gn[2].data_ptr() if gn else None,
        gn_eps=1e-5, gn_act=gn[3] if gn else 0, a_up=a_up, force_tma=force_tma,
        gn_stats2=gn[4].data_ptr() if (gn and len(gn) > 4 and gn[4] is not None) else None,
        dw_w=dw[0].data_ptr() if dw else None, dw_n=dw[1] if dw else 0, a_softmax_h=1 if softmax_h else 0)
    if sync:
        torch.cuda.synchronize()
    return (out_nchw if want_nchw else out), stats
