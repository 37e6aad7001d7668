"""Haar ("db1") single-level 2-D DWT / IDWT and the conditioning assembly, CPU restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference calls PyWavelets (`pywt.wavedec2(x, "db1", level=1)`), a third-party C
extension that is NOT vendored in /root/reference, NOT version-pinned by it (no
requirements file) and NOT installed here:
  /root/reference/dataset/pan_dataset.py:75-80,97-102 and dataset/hisr.py:50-55.
This file restates PyWavelets' published db1 filter bank
  dec_lo = [1/sqrt2, 1/sqrt2],  dec_hi = [-1/sqrt2, 1/sqrt2]
with `dwt2` returning (cA, (cH, cV, cD)) = (aa, (da, ad, dd)).  For even sizes (all the
reference uses) the default 'symmetric' boundary mode never engages.
PARITY UNPINNED beyond the two PyWavelets documentation known answers checked in
tests/test_oracle_golden.py:
  pywt.dwt([1,2,3,4],'db1') -> ([2.12132034, 4.94974747], [-0.70710678, -0.70710678])
  [[1,2],[3,4]] -> cA 5, cH -2, cV -1, cD 0.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def haar_dwt1(x: np.ndarray):
    """1-D db1 analysis along the last axis (even length)."""
    x = np.asarray(x, dtype=np.float64)
    a, b = x[..., 0::2], x[..., 1::2]
    s = 1.0 / np.sqrt(2.0)
    return (a + b) * s, (a - b) * s


def haar_dwt2(x: np.ndarray):
    """cA, (cH, cV, cD) over the last two axes; per 2x2 block [[a,b],[c,d]]:
    cA=(a+b+c+d)/2, cH=(a+b-c-d)/2, cV=(a-b+c-d)/2, cD=(a-b-c+d)/2."""
    x = np.asarray(x, dtype=np.float64)
    a = x[..., 0::2, 0::2]
    b = x[..., 0::2, 1::2]
    c = x[..., 1::2, 0::2]
    d = x[..., 1::2, 1::2]
    return (a + b + c + d) / 2, ((a + b - c - d) / 2, (a - b + c - d) / 2, (a - b - c + d) / 2)


def haar_idwt2(cA, cH, cV, cD):
    """Inverse of haar_dwt2.  The reference has no IDWT call site; parity = round trip."""
    cA, cH, cV, cD = (np.asarray(v, dtype=np.float64) for v in (cA, cH, cV, cD))
    out = np.empty(cA.shape[:-2] + (cA.shape[-2] * 2, cA.shape[-1] * 2), dtype=np.float64)
    out[..., 0::2, 0::2] = (cA + cH + cV + cD) / 2
    out[..., 0::2, 1::2] = (cA + cH - cV - cD) / 2
    out[..., 1::2, 0::2] = (cA - cH + cV - cD) / 2
    out[..., 1::2, 1::2] = (cA - cH - cV + cD) / 2
    return out


def wavelet_channels(lms_dn: np.ndarray, pan_dn: np.ndarray, division: float, order: str = "pan") -> torch.Tensor:
    """Wavelet conditioning channels.

    order="pan":  [LL(lms), pan_cH, pan_cD, pan_cV] / division  (pan_dataset.py:127-142 — note h, d, v)
    order="hisr": [LL(hsi_up), rgb_cH, rgb_cV, rgb_cD]          (hisr.py:57-59, no division)"""
    lms_main, _ = haar_dwt2(lms_dn)
    _, (ph, pv, pd) = haar_dwt2(pan_dn)
    parts = [lms_main, ph, pd, pv] if order == "pan" else [lms_main, ph, pv, pd]
    return torch.cat([torch.tensor(p / division, dtype=torch.float32) for p in parts], dim=1)


def assemble_cond(lms: torch.Tensor, pan: torch.Tensor, wavelets: torch.Tensor) -> torch.Tensor:
    """cond = cat([lms, pan, bilinear(wavelets -> W)]) (diffusion_engine.py:221-228,441-444)."""
    up = F.interpolate(wavelets, size=lms.shape[-1], mode="bilinear")
    return torch.cat([lms, pan, up], dim=1)
