"""Seeded synthetic inputs and weights for the DDIF hot path (no datasets, no checkpoints here).

Shapes and value conventions follow SURVEY.md §8(d):
  * WV3  C=8,  P=1, division 2047  (diffusion_engine.py:107)
  * GF2  C=4,  P=1, division 1023
  * QB   C=4,  P=1, division 2047
  * CAVE C=31, P=3, division 1 (already in [0,1]; hisr.py:43,113-119), HISR wavelet order
`make_state_dict` produces a reference-compatible `state_dict` (same 702 names and shapes as
models/sr3_dwt.py builds) from a seed, so the reference, the oracle and the CUDA path can share
weights without shipping a 40 MB checkpoint.  CondInjection's zero-initialised last conv
(sr3_dwt.py:386-387) is given non-zero values so the CSM branch is exercised.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class DatasetSpec:
    name: str
    bands: int
    pan_bands: int
    division: float
    wavelet_order: str  # "pan" -> [LL, H, D, V];  "hisr" -> [LL, H, V, D]


DATASETS = {
    "wv3": DatasetSpec("wv3", 8, 1, 2047.0, "pan"),
    "gf2": DatasetSpec("gf2", 4, 1, 1023.0, "pan"),
    "qb": DatasetSpec("qb", 4, 1, 2047.0, "pan"),
    "cave": DatasetSpec("cave", 31, 3, 1.0, "hisr"),
}


def _haar_dwt2_np(x: np.ndarray):
    a, b = x[..., 0::2, 0::2], x[..., 0::2, 1::2]
    c, d = x[..., 1::2, 0::2], x[..., 1::2, 1::2]
    return (a + b + c + d) / 2, (a + b - c - d) / 2, (a - b + c - d) / 2, (a - b - c + d) / 2


def make_batch(dataset: str, batch: int, size: int = 64, seed: int = 1234, dwt2=_haar_dwt2_np) -> Dict[str, torch.Tensor]:
    """Image-like synthetic sample: hr (GT), ms, lms, pan, wavelets, cond — all fp32 CPU tensors.

    hr = clamp(smooth low-frequency field + 0.05*randn, 0, 1); pan = band mean (per pan band a
    different band subset); ms = 4x area down-sample; lms = bicubic x4 up-sample.  Pan-type sets are
    quantised to integer DN before the float64 DWT, like the reference's h5 data."""
    ds = DATASETS[dataset]
    g = torch.Generator().manual_seed(seed)
    C, P = ds.bands, ds.pan_bands
    low = torch.rand(batch, C, 8, 8, generator=g)
    hr = F.interpolate(low, size=(size, size), mode="bicubic", align_corners=False)
    hr = (hr + 0.05 * torch.randn(batch, C, size, size, generator=g)).clamp(0, 1)
    groups = torch.chunk(torch.arange(C), P)
    pan = torch.stack([hr[:, idx].mean(1) for idx in groups], dim=1)
    ms = F.avg_pool2d(hr, 4)
    lms = F.interpolate(ms, size=(size, size), mode="bicubic", align_corners=False).clamp(0, 1)
    if ds.division != 1.0:
        hr, pan, ms, lms = (torch.round(v * ds.division) for v in (hr, pan, ms, lms))
    lms_dn = lms.double().numpy()
    pan_dn = pan.double().numpy()
    ll, _, _, _ = dwt2(lms_dn)
    _, ph, pv, pd = dwt2(pan_dn)
    parts = [ll, ph, pd, pv] if ds.wavelet_order == "pan" else [ll, ph, pv, pd]
    wave = torch.cat([torch.tensor(p / ds.division, dtype=torch.float32) for p in parts], dim=1)
    hr, pan, ms, lms = (v / ds.division for v in (hr, pan, ms, lms))
    up = F.interpolate(wave, size=size, mode="bilinear")
    cond = torch.cat([lms, pan, up], dim=1)
    return dict(hr=hr, ms=ms, lms=lms, pan=pan, wavelets=wave, cond=cond,
                lms_dn=torch.tensor(lms_dn), pan_dn=torch.tensor(pan_dn))


def unet_kwargs(dataset: str, image_size: int = 64) -> dict:
    """Production hyper-parameters (diffusion_engine.py:121-133,381-393)."""
    ds = DATASETS[dataset]
    return dict(in_channel=ds.bands, out_channel=ds.bands, lms_channel=ds.bands, pan_channel=ds.pan_bands,
                inner_channel=32, norm_groups=1, channel_mults=(1, 2, 2, 4), attn_res=(8,), dropout=0.2,
                image_size=image_size, self_condition=True)


def _shapes(in_channel, out_channel, inner_channel, lms_channel, pan_channel, channel_mults, attn_res,
            res_blocks, image_size, self_condition) -> Dict[str, Tuple[int, ...]]:
    """Names and shapes of the reference state_dict (models/sr3_dwt.py:52-163)."""
    sh: Dict[str, Tuple[int, ...]] = {}
    ic = inner_channel

    def conv(p, co, ci, k, bias=True, groups=1):
        sh[p + ".weight"] = (co, ci // groups, k, k)
        if bias:
            sh[p + ".bias"] = (co,)

    def lin(p, co, ci):
        sh[p + ".weight"] = (co, ci)
        sh[p + ".bias"] = (co,)

    def gn(p, c):
        sh[p + ".weight"] = (c,)
        sh[p + ".bias"] = (c,)

    def resblock(p, d):
        lin(p + ".noise_func.noise_func.0", d, ic)
        gn(p + ".block1.block.0", d)
        conv(p + ".block1.block.3", d, d, 3)
        gn(p + ".block2.block.0", d)
        conv(p + ".block2.block.3", d, d, 3)

    def attn(p, d):
        gn(p + ".norm", d)
        conv(p + ".qkv", 3 * d, d, 1, bias=False)
        conv(p + ".out", d, d, 1)

    lin("noise_level_mlp.1", 4 * ic, ic)
    lin("noise_level_mlp.3", ic, 4 * ic)
    cin = in_channel + (out_channel if self_condition else 0)
    conv("downs.0", ic, cin, 3)
    nd = 1
    pre, feat, res = ic, [ic], image_size
    n = len(channel_mults)
    ce, cdec = lms_channel + pan_channel, lms_channel + 3 * pan_channel
    for lvl in range(n):
        ch = ic * channel_mults[lvl]
        for _ in range(res_blocks):
            p = f"downs.{nd}"
            conv(p + ".cond_inj.body.0", 4 * ch, ce, 3, bias=False)
            gn(p + ".cond_inj.body.1", 4 * ch)
            conv(p + ".cond_inj.body.3", 2 * ch, 4 * ch, 1)
            conv(p + ".cond_inj.x_conv", ch, pre, 1)
            resblock(p + ".res_block", ch)
            if res in attn_res:
                attn(p + ".attn", ch)
            feat.append(ch)
            pre = ch
            nd += 1
        if lvl != n - 1:
            conv(f"downs.{nd}.conv", pre, pre, 3)
            feat.append(pre)
            res //= 2
            nd += 1
    resblock("mid.0.res_block", pre)
    attn("mid.0.attn", pre)
    resblock("mid.1.res_block", pre)
    nu = 0
    for lvl in reversed(range(n)):
        ch = ic * channel_mults[lvl]
        for _ in range(res_blocks + 1):
            p = f"ups.{nu}"
            dim = pre + feat.pop()
            q = p + ".cond_inj"
            gn(q + ".prenorm_x", dim)
            conv(q + ".q.0", dim, dim, 3, bias=False, groups=dim)
            conv(q + ".q.1", dim, dim, 1)
            conv(q + ".kv.0", cdec, cdec, 3, bias=False, groups=cdec)
            conv(q + ".kv.1", 2 * dim, cdec, 1)
            conv(q + ".attn_out", ch, dim, 1)
            if dim != ch:
                conv(q + ".attn_res", ch, dim, 1)
            conv(q + ".ffn.0", 2 * ch, ch, 3, bias=False)
            conv(q + ".ffn.2", ch, 2 * ch, 3, bias=False)
            conv(q + ".ffn.3", ch, ch, 1)
            resblock(p + ".res_block", ch)
            if res in attn_res:
                attn(p + ".attn", ch)
            pre = ch
            nu += 1
        if lvl >= 1:
            conv(f"ups.{nu}.conv", pre, pre, 3)
            res *= 2
            nu += 1
    gn("final_conv.block.0", pre)
    conv("final_conv.block.3", out_channel, pre, 3)
    return sh


def make_state_dict(seed: int = 0, *, in_channel=8, out_channel=8, inner_channel=32, lms_channel=8, pan_channel=1,
                    channel_mults=(1, 2, 2, 4), attn_res=(8,), res_blocks=3, image_size=64, self_condition=True,
                    **_ignored) -> Dict[str, torch.Tensor]:
    """Seeded weights with PyTorch-default-like scales: conv/linear U(-1/sqrt(fan_in), +); GroupNorm
    gamma = 1 + 0.1 n, beta = 0.1 n (so the affine is exercised)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    shapes = _shapes(in_channel, out_channel, inner_channel, lms_channel, pan_channel, tuple(channel_mults),
                     tuple(attn_res), res_blocks, image_size, self_condition)
    for name, shp in shapes.items():
        is_norm = any(k in name for k in (".block.0.", ".norm.", ".body.1.", ".prenorm_x."))
        if is_norm:
            base = 1.0 if name.endswith("weight") else 0.0
            sd[name] = base + 0.1 * torch.randn(shp, generator=g)
        else:
            if name.endswith("weight"):
                fan_in = int(np.prod(shp[1:]))
            else:
                w = shapes[name[: -len("bias")] + "weight"]
                fan_in = int(np.prod(w[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            sd[name] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    return sd
