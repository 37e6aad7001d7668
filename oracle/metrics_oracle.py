"""SAM / ERGAS / PSNR restated from the reference's validation metrics.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
/root/reference/utils/_metric_legacy.py:299-346,365 (`analysis_accu`) as called by
utils/metric.py:24-30 (ratio 4).  These are the parity yard-stick for final fused
images (ΔPSNR <= 0.1 dB, ΔSAM / ΔERGAS <= 0.05).
"""
from __future__ import annotations

import math
from typing import Dict

import torch


def sam_ergas_psnr(gt_chw: torch.Tensor, out_chw: torch.Tensor, ratio: int = 4) -> Dict[str, float]:
    """One image, [C,H,W] each.  Quirks kept: bounds cut is `[0:-1]` on both spatial axes
    (dim_cut=1 -> slice(0,-1), :300-302); pi ~ 3.14159256 and the mean angle is rounded to 6
    decimals (:328-330); PSNR = mean over bands of -20*log10(1/rmse) (:342-346,365)."""
    a = gt_chw.permute(1, 2, 0)[0:-1, 0:-1, :].to(torch.float32)
    b = out_chw.permute(1, 2, 0)[0:-1, 0:-1, :].to(torch.float32)
    s1 = (a * b).sum(2)
    t = ((a * a).sum(2) * (b * b).sum(2)) ** 0.5
    num = (t > 0).sum()
    ang = torch.acos(s1 / t)
    tot = torch.where(torch.isnan(ang), torch.zeros_like(ang), ang).sum()
    aver = tot if num == 0 else tot / num
    aver = (aver * 10**6).round() / 10**6
    sam = aver * 180 / 3.14159256
    summ = 0.0
    for i in range(a.shape[2]):
        mse_i = torch.mean((a[:, :, i] - b[:, :, i]) ** 2)
        m = torch.mean(a[:, :, i])
        summ = summ + mse_i / (m * m)
    ergas = 100 * (1 / ratio) * ((summ / a.shape[2]) ** 0.5)
    rmse = torch.mean(torch.mean((a - b) ** 2, 0), 0) ** 0.5
    psnr = torch.mean(-20 * (torch.log(1 / rmse) / math.log(10)))
    # CC (:367-375, choices=5)
    hw = a.shape[0] * a.shape[1]
    ma, mb = torch.mean(torch.mean(a, 0), 0), torch.mean(torch.mean(b, 0), 0)
    c1 = torch.sum(torch.sum(a * b, 0), 0) - hw * (ma * mb)
    c2 = torch.sum(torch.sum(b ** 2, 0), 0) - hw * (mb ** 2)
    c3 = torch.sum(torch.sum(a ** 2, 0), 0) - hw * (ma ** 2)
    cc = torch.mean(c1 / ((c2 * c3) ** 0.5))
    return {"SAM": float(sam), "ERGAS": float(ergas), "PSNR": float(psnr), "CC": float(cc)}


def batch_metrics(gt: torch.Tensor, out: torch.Tensor, ratio: int = 4) -> Dict[str, float]:
    acc = {"SAM": 0.0, "ERGAS": 0.0, "PSNR": 0.0, "CC": 0.0}
    for g, o in zip(gt, out):
        m = sam_ergas_psnr(g, o, ratio)
        for k in acc:
            acc[k] += m[k] / gt.shape[0]
    return acc
