"""Haar ("db1") DWT / IDWT and DDIF conditioning assembly on the GPU.

Replaces the CPU/float64 `pywt.wavedec2(x, "db1", level=1)` calls of the reference's datasets
(/root/reference/dataset/pan_dataset.py:75-80,97-102,127-142; dataset/hisr.py:50-59) and the `cond` packing of
/root/reference/diffusion_engine.py:221-228 with memory-bound fp32 CUDA kernels (csrc/sampler.cu):
8 bytes of HBM traffic per input element for DWT or IDWT.  No CPU fallback.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib


def _stream(t: torch.Tensor) -> int:
    return _lib.stream_of(t.device)


def _check(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("dif_pan_b200.wavelet runs on CUDA only (no CPU fallback)")
    if x.shape[-2] % 2 or x.shape[-1] % 4:
        raise ValueError(f"Haar DWT kernels need even height and width % 4 == 0, got {tuple(x.shape[-2:])}")
    return x.to(torch.float32).contiguous()


def haar_dwt2(x: torch.Tensor, divisor: float = 1.0) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
    """(cA, (cH, cV, cD)) over the last two axes, every coefficient divided by `divisor`
    (the dataset "division" applied after the DWT in pan_dataset.py:127-142)."""
    x = _check(x)
    lead, (h, w) = x.shape[:-2], x.shape[-2:]
    outs = [torch.empty(*lead, h // 2, w // 2, dtype=torch.float32, device=x.device) for _ in range(4)]
    planes = x.numel() // (h * w)
    if planes:
        _lib.launch("ddif_haar_t", _stream(x), kind="DDIF_OP_HAAR_DWT2", x=x.data_ptr(), ll=outs[0].data_ptr(), ch=outs[1].data_ptr(),
                    cv=outs[2].data_ptr(), cd=outs[3].data_ptr(), planes=planes, h=h, w=w, divisor=float(divisor))
    return outs[0], (outs[1], outs[2], outs[3])


def haar_idwt2(cA: torch.Tensor, coeffs: Tuple[torch.Tensor, torch.Tensor, torch.Tensor]) -> torch.Tensor:
    """Inverse of haar_dwt2 (divisor 1).  The reference has no IDWT call site; parity = round trip."""
    cH, cV, cD = coeffs
    bands = [b.to(torch.float32).contiguous() for b in (cA, cH, cV, cD)]
    if not bands[0].is_cuda:
        raise RuntimeError("dif_pan_b200.wavelet runs on CUDA only (no CPU fallback)")
    lead, (hh, wh) = cA.shape[:-2], cA.shape[-2:]
    if wh % 2:
        raise ValueError("Haar IDWT kernel needs an even coefficient width")
    out = torch.empty(*lead, hh * 2, wh * 2, dtype=torch.float32, device=cA.device)
    planes = cA.numel() // max(hh * wh, 1)
    if planes and hh and wh:
        _lib.launch("ddif_haar_t", _stream(cA), kind="DDIF_OP_HAAR_IDWT2", x=out.data_ptr(), ll=bands[0].data_ptr(), ch=bands[1].data_ptr(),
                    cv=bands[2].data_ptr(), cd=bands[3].data_ptr(), planes=planes, h=hh * 2, w=wh * 2, divisor=1.0)
    return out


def wavelet_channels(lms_dn: torch.Tensor, pan_dn: torch.Tensor, division: float, order: str = "pan") -> torch.Tensor:
    """order='pan':  [LL(lms), pan_cH, pan_cD, pan_cV] / division   (pan_dataset.py:139-142 — note h, d, v)
    order='hisr': [LL(hsi_up), rgb_cH, rgb_cV, rgb_cD]             (hisr.py:57-59)"""
    ll, _ = haar_dwt2(lms_dn, division)
    _, (ph, pv, pd) = haar_dwt2(pan_dn, division)
    parts = [ll, ph, pd, pv] if order == "pan" else [ll, ph, pv, pd]
    return torch.cat(parts, dim=1)


def assemble_cond(lms: torch.Tensor, pan: torch.Tensor, wavelets: torch.Tensor) -> torch.Tensor:
    """cond = cat([lms, pan, bilinear(wavelets -> H x W)], dim=1) (diffusion_engine.py:221-228), one kernel."""
    if not lms.is_cuda:
        raise RuntimeError("dif_pan_b200.wavelet runs on CUDA only (no CPU fallback)")
    lms, pan, wavelets = (t.to(torch.float32).contiguous() for t in (lms, pan, wavelets))
    B, C, H, W = lms.shape
    P, CW = pan.shape[1], wavelets.shape[1]
    cond = torch.empty(B, C + P + CW, H, W, dtype=torch.float32, device=lms.device)
    _lib.launch("ddif_cond_assemble_t", _stream(lms), lms=lms.data_ptr(), pan=pan.data_ptr(), wav=wavelets.data_ptr(), cond=cond.data_ptr(),
                batch=B, c=C, p=P, cw=CW, h=H, w=W, wh=wavelets.shape[2], ww=wavelets.shape[3])
    return cond


def make_cond(lms_dn: torch.Tensor, pan_dn: torch.Tensor, division: float = 1.0, order: str = "pan", return_wavelets: bool = False):
    """Raw `lms` [B,C,H,W] and `pan` [B,P,H,W] (digital numbers, or [0,1] data with division=1) -> `cond` [B, 2C+4P, H, W] in ONE
    kernel (`ddif_wavelet_cond_f32`): Haar DWT of both, /division, the dataset's channel order, bilinear x2 of the wavelet stack and
    the concat with lms/division and pan/division (pan_dataset.py:73-142, hisr.py:48-59, diffusion_engine.py:221-228).
    Same arithmetic as `assemble_cond(lms/div, pan/div, wavelet_channels(...))` up to 1.5 ulp (multiply by 1/division instead of dividing);
    4 B read + 4 B written per cond element instead of
    three kernels and two intermediate tensors.  With return_wavelets also returns the [B, C+3P, H/2, W/2] stack the datasets yield."""
    if not (lms_dn.is_cuda and pan_dn.is_cuda):
        raise RuntimeError("dif_pan_b200.wavelet runs on CUDA only (no CPU fallback)")
    if order not in ("pan", "hisr"):
        raise ValueError(f"order must be 'pan' or 'hisr', got {order!r}")
    lms_dn, pan_dn = lms_dn.to(torch.float32).contiguous(), pan_dn.to(torch.float32).contiguous()
    B, C, H, W = lms_dn.shape
    P = pan_dn.shape[1]
    if pan_dn.shape[0] != B or tuple(pan_dn.shape[2:]) != (H, W) or H % 2 or W % 4:
        raise ValueError(f"lms {tuple(lms_dn.shape)} and pan {tuple(pan_dn.shape)} must share batch and H x W with even H and W % 4 == 0")
    cond = torch.empty(B, 2 * C + 4 * P, H, W, dtype=torch.float32, device=lms_dn.device)
    wav = torch.empty(B, C + 3 * P, H // 2, W // 2, dtype=torch.float32, device=lms_dn.device) if return_wavelets else None
    _lib.launch("ddif_wavelet_cond_t", _stream(lms_dn), lms=lms_dn.data_ptr(), pan=pan_dn.data_ptr(), cond=cond.data_ptr(),
                wav=wav.data_ptr() if wav is not None else None, batch=B, c=C, p=P, h=H, w=W, order=0 if order == "pan" else 1,
                divisor=float(division))
    return (cond, wav) if return_wavelets else cond
