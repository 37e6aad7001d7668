"""On-device validation metrics: SAM / ERGAS / PSNR / CC of a batch of fused images.

Replaces the per-image CPU loop of /root/reference/utils/_metric_legacy.py:299-379 (`analysis_accu`) as driven by
`AnalysisPanAcc` (/root/reference/utils/metric.py:24-99, ERGAS ratio 4): ONE kernel launch reduces every image of the batch
to 2 + 6*C partial sums in fp64 (`ddif_metrics_f32`, csrc/prep_post.cu); the closed-form finish below is a few scalars
per image.  The reference's quirks are kept: the `[0:-1]` bounds cut on both spatial axes (:300-302), pi ~ 3.14159256 and
the mean spectral angle rounded to 6 decimals (:328-330), PSNR = mean over bands of -20*log10(1/rmse) (:342-346,365).
SSIM (skimage in the reference) is not part of the path's tolerances and is not computed.  No CPU fallback.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import _lib


def image_sums(gt: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """[B, 2 + 6*C] fp64 partial sums (layout in include/ddif_b200.h, ddif_metrics_t)."""
    if not (gt.is_cuda and out.is_cuda):
        raise RuntimeError("dif_pan_b200.metrics runs on CUDA only (no CPU fallback)")
    if gt.shape != out.shape or gt.dim() != 4:
        raise ValueError(f"gt and out must both be [B, C, H, W], got {tuple(gt.shape)} and {tuple(out.shape)}")
    gt, out = gt.to(torch.float32).contiguous(), out.to(torch.float32).contiguous()
    B, C, H, W = gt.shape
    sums = torch.zeros(B, 2 + 6 * C, dtype=torch.float64, device=gt.device)
    _lib.launch("ddif_metrics_t", _lib.stream_of(gt.device), gt=gt.data_ptr(), out=out.data_ptr(),
                sums=sums.data_ptr(), batch=B, c=C, h=H, w=W)
    return sums


def per_image_metrics(gt: torch.Tensor, out: torch.Tensor, ratio: int = 4) -> Dict[str, torch.Tensor]:
    """SAM (degrees), ERGAS, PSNR (dB), CC per image as fp64 device tensors of shape [B]."""
    B, C, H, W = gt.shape
    s = image_sums(gt, out)
    n = float((H - 1) * (W - 1))
    tot, num = s[:, 0], s[:, 1]
    band = s[:, 2:].reshape(B, C, 6)
    sse, sa, sb, saa, sbb, sab = (band[:, :, i] for i in range(6))
    aver = torch.where(num == 0, tot, tot / num.clamp_min(1.0))
    aver = torch.round(aver * 10 ** 6) / 10 ** 6
    sam = aver * 180 / 3.14159256
    mse = sse / n
    mean_a = sa / n
    ergas = 100 * (1 / ratio) * torch.sqrt((mse / (mean_a * mean_a)).sum(1) / C)
    psnr = (-20 * (torch.log(1 / torch.sqrt(mse)) / math.log(10))).mean(1)
    c1 = sab - n * (sa / n) * (sb / n)
    c2 = sbb - n * (sb / n) ** 2
    c3 = saa - n * (sa / n) ** 2
    cc = (c1 / torch.sqrt(c2 * c3)).mean(1)
    return {"SAM": sam, "ERGAS": ergas, "PSNR": psnr, "CC": cc}


def batch_metrics(gt: torch.Tensor, out: torch.Tensor, ratio: int = 4) -> Dict[str, float]:
    """Mean over the batch (what `AnalysisPanAcc.sam_ergas_psnr_cc_batch` returns, metric.py:73-82)."""
    return {k: float(v.mean()) for k, v in per_image_metrics(gt, out, ratio).items()}


class AnalysisPanAcc:
    """Running average over calls, like /root/reference/utils/metric.py:24-99 (without SSIM)."""

    def __init__(self, ergas_ratio: int = 4):
        self.ratio = ergas_ratio
        self.clear_history()

    @property
    def last_acc(self):
        return self._acc_d

    def clear_history(self):
        self._acc_d: Dict[str, float] = {}
        self._call_n = 0
        self.acc_ave = {"SAM": 0.0, "ERGAS": 0.0, "PSNR": 0.0, "CC": 0.0}

    def once_batch_call(self, b_gt, b_pred):
        self._acc_d = batch_metrics(b_gt, b_pred, self.ratio)
        return self._acc_d

    def __call__(self, b_gt, b_pred):
        n = b_gt.shape[0]
        now = self.once_batch_call(b_gt, b_pred)
        for k in self.acc_ave:
            self.acc_ave[k] = (self.acc_ave[k] * self._call_n + now[k] * n) / (self._call_n + n)
        self._call_n += n
        return self.acc_ave

    def print_str(self, acc_d=None):
        acc_d = self.acc_ave if acc_d is None else acc_d
        return ", ".join(f"{k}: {v:.6f}" for k, v in acc_d.items())
