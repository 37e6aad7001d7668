"""Functional CPU restatement of the SR3-DWT conditional UNet forward.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
/root/reference/models/sr3_dwt.py; line numbers below refer to that file.
The function consumes a plain reference `state_dict` (702 entries for the
production hyper-parameters) so the same weights drive the reference, this
oracle and the CUDA path.

Only inference semantics are restated (Dropout / DropPath are identity in
eval mode, sr3_dwt.py:295,534).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class UNetCfg:
    """Constructor arguments of UNetSR3 (sr3_dwt.py:31-51) that matter at inference."""

    in_channel: int = 8
    out_channel: int = 8
    inner_channel: int = 32
    lms_channel: int = 8
    pan_channel: int = 1
    norm_groups: int = 1
    channel_mults: Tuple[int, ...] = (1, 2, 2, 4)
    attn_res: Tuple[int, ...] = (8,)
    res_blocks: int = 3
    image_size: int = 64
    self_condition: bool = True


@dataclass
class _Blk:
    kind: str  # "conv0" | "enc" | "down" | "mid" | "dec" | "up"
    name: str  # state-dict prefix
    dim: int = 0
    dim_out: int = 0
    attn: bool = False


def topology(cfg: UNetCfg) -> Tuple[List[_Blk], List[_Blk], List[_Blk]]:
    """Module order of downs / mid / ups (sr3_dwt.py:86-159)."""
    downs: List[_Blk] = [_Blk("conv0", "downs.0")]
    pre = cfg.inner_channel
    feat = [pre]
    res = cfg.image_size
    n = len(cfg.channel_mults)
    for lvl in range(n):
        ch = cfg.inner_channel * cfg.channel_mults[lvl]
        for _ in range(cfg.res_blocks):
            downs.append(_Blk("enc", f"downs.{len(downs)}", pre, ch, res in cfg.attn_res))
            feat.append(ch)
            pre = ch
        if lvl != n - 1:
            downs.append(_Blk("down", f"downs.{len(downs)}", pre, pre))
            feat.append(pre)
            res //= 2
    mid = [_Blk("mid", "mid.0", pre, pre, True), _Blk("mid", "mid.1", pre, pre, False)]
    ups: List[_Blk] = []
    for lvl in reversed(range(n)):
        ch = cfg.inner_channel * cfg.channel_mults[lvl]
        for _ in range(cfg.res_blocks + 1):
            ups.append(_Blk("dec", f"ups.{len(ups)}", pre + feat.pop(), ch, res in cfg.attn_res))
            pre = ch
        if lvl >= 1:
            ups.append(_Blk("up", f"ups.{len(ups)}", pre, pre))
            res *= 2
    return downs, mid, ups


def _swish(x):
    return x * torch.sigmoid(x)  # sr3_dwt.py:261-263


def _gn(x, sd, p, groups):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps=1e-5)


def time_embedding(sd, cfg: UNetCfg, time: torch.Tensor) -> torch.Tensor:
    """PositionalEncoding + noise_level_mlp (sr3_dwt.py:57-64, 223-238).

    `step` is built in the dtype of `time` (long for DDPM/DDIM, float for
    DPM-Solver) exactly as the reference does (:230-233)."""
    count = cfg.inner_channel // 2
    step = torch.arange(count, dtype=time.dtype, device=time.device) / count
    enc = time.unsqueeze(1) * torch.exp(-math.log(1e4) * step.unsqueeze(0))
    enc = torch.cat([torch.sin(enc), torch.cos(enc)], dim=-1)
    h = F.linear(enc, sd["noise_level_mlp.1.weight"], sd["noise_level_mlp.1.bias"])
    h = _swish(h)
    return F.linear(h, sd["noise_level_mlp.3.weight"], sd["noise_level_mlp.3.bias"])


def _block(x, sd, p, groups):
    """Block = GN -> Swish -> (Dropout) -> Conv3x3 (sr3_dwt.py:288-300)."""
    h = _swish(_gn(x, sd, p + ".block.0", groups))
    return F.conv2d(h, sd[p + ".block.3.weight"], sd[p + ".block.3.bias"], padding=1)


def resnet_block(x, t_emb, sd, p, groups):
    """ResnetBlock (sr3_dwt.py:303-327) with additive FiLM (:241-258)."""
    h = _block(x, sd, p + ".block1", groups)
    film = F.linear(t_emb, sd[p + ".noise_func.noise_func.0.weight"], sd[p + ".noise_func.noise_func.0.bias"])
    h = h + film[:, :, None, None]
    h = _block(h, sd, p + ".block2", groups)
    if (p + ".res_conv.weight") in sd:
        x = F.conv2d(x, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])
    return h + x


def self_attention(x, sd, p, groups, n_head=8):
    """SelfAttention (sr3_dwt.py:330-360).  Scale uses the FULL channel count (:352)."""
    b, c, h, w = x.shape
    hd = c // n_head
    qkv = F.conv2d(_gn(x, sd, p + ".norm", groups), sd[p + ".qkv.weight"])
    qkv = qkv.view(b, n_head, hd * 3, h * w)
    q, k, v = qkv[:, :, :hd], qkv[:, :, hd : 2 * hd], qkv[:, :, 2 * hd :]
    att = torch.einsum("bncq,bnck->bnqk", q, k) / math.sqrt(c)
    att = torch.softmax(att, dim=-1)
    out = torch.einsum("bnqk,bnck->bncq", att, v).reshape(b, c, h, w)
    out = F.conv2d(out, sd[p + ".out.weight"], sd[p + ".out.bias"])
    return out + x


def csm_modulation(c, sd, p, groups):
    """cond-only part of CondInjection: body(cond) -> (scale, shift) (sr3_dwt.py:379-391)."""
    h = F.conv2d(c, sd[p + ".body.0.weight"], None, padding=1)
    h = F.silu(_gn(h, sd, p + ".body.1", groups))
    h = F.conv2d(h, sd[p + ".body.3.weight"], sd[p + ".body.3.bias"])
    return h.chunk(2, dim=1)


def cond_injection(x, c, sd, p, groups):
    """CondInjection / CSM (sr3_dwt.py:376-396)."""
    scale, shift = csm_modulation(c, sd, p, groups)
    x = F.conv2d(x, sd[p + ".x_conv.weight"], sd[p + ".x_conv.bias"])
    return x * (1 + scale) + shift


def fwm_context(c, sd, p, n_head=8):
    """cond-only part of FastAttnCondInjection: k,v -> softmax_W(k) -> context (sr3_dwt.py:541,546,563)."""
    cd = c.shape[1]
    kv = F.conv2d(c, sd[p + ".kv.0.weight"], None, padding=1, groups=cd)
    kv = F.conv2d(kv, sd[p + ".kv.1.weight"], sd[p + ".kv.1.bias"])
    k, v = kv.chunk(2, dim=1)
    k = k.softmax(dim=-1)
    b, dim, h, w = k.shape
    d = dim // n_head
    k = k.reshape(b, n_head, d, h * w)
    v = v.reshape(b, n_head, d, h * w)
    return torch.einsum("bhdn,bhen->bhde", k, v)


def fwm_injection(x, c, sd, p, groups, n_head=8):
    """FastAttnCondInjection / FWM (sr3_dwt.py:493-577), eval mode (DropPath = identity)."""
    xh = _gn(x, sd, p + ".prenorm_x", groups)
    dim = xh.shape[1]
    q = F.conv2d(xh, sd[p + ".q.0.weight"], None, padding=1, groups=dim)
    q = F.conv2d(q, sd[p + ".q.1.weight"], sd[p + ".q.1.bias"])
    ctx = fwm_context(c, sd, p, n_head)
    q = q.softmax(dim=-2)
    b, _, h, w = q.shape
    d = dim // n_head
    q = q.reshape(b, n_head, d, h * w) * (1.0 / math.sqrt(dim // n_head))
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, dim, h, w)
    y = F.conv2d(out, sd[p + ".attn_out.weight"], sd[p + ".attn_out.bias"])
    if (p + ".attn_res.weight") in sd:
        y = y + F.conv2d(xh, sd[p + ".attn_res.weight"], sd[p + ".attn_res.bias"])
    else:
        y = y + xh
    f = F.conv2d(y, sd[p + ".ffn.0.weight"], None, padding=1)
    f = F.silu(f)
    f = F.conv2d(f, sd[p + ".ffn.2.weight"], None, padding=1)
    f = F.conv2d(f, sd[p + ".ffn.3.weight"], sd[p + ".ffn.3.bias"])
    return f + y


def resize_cond(c, size):
    """F.interpolate(cond, size, 'bilinear') (sr3_dwt.py:661-663): align_corners=False, no antialias."""
    return F.interpolate(c, size=size, mode="bilinear")


def unet_forward(
    sd: Dict[str, torch.Tensor],
    cfg: UNetCfg,
    x: torch.Tensor,
    time: torch.Tensor,
    cond: torch.Tensor,
    self_cond: Optional[torch.Tensor] = None,
    taps: Optional[Dict[str, torch.Tensor]] = None,
) -> torch.Tensor:
    """UNetSR3.forward (sr3_dwt.py:169-219).

    `taps`, if given, receives the output of every block keyed by its
    state-dict prefix ("downs.3", "mid.0", "ups.7", ...) for layer-by-layer
    debugging of the CUDA path."""
    g = cfg.norm_groups
    if cfg.self_condition:
        sc = x if self_cond is None else self_cond
        x = torch.cat([sc, x], dim=1)
    t = time_embedding(sd, cfg, time)
    downs, mid, ups = topology(cfg)
    c_enc = cond[:, : cfg.lms_channel + cfg.pan_channel]
    c_dec = cond[:, -(cfg.lms_channel + 3 * cfg.pan_channel) :]
    feats = []
    for blk in downs:
        p = blk.name
        if blk.kind == "conv0":
            x = F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif blk.kind == "down":
            x = F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], stride=2, padding=1)
        else:
            x = cond_injection(x, resize_cond(c_enc, x.shape[-2:]), sd, p + ".cond_inj", g)
            x = resnet_block(x, t, sd, p + ".res_block", g)
            if blk.attn:
                x = self_attention(x, sd, p + ".attn", g)
        feats.append(x)
        if taps is not None:
            taps[p] = x
    for blk in mid:
        p = blk.name
        x = resnet_block(x, t, sd, p + ".res_block", g)
        if blk.attn:
            x = self_attention(x, sd, p + ".attn", g)
        if taps is not None:
            taps[p] = x
    for blk in ups:
        p = blk.name
        if blk.kind == "up":
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
        else:
            x = torch.cat((x, feats.pop()), dim=1)
            x = fwm_injection(x, resize_cond(c_dec, x.shape[-2:]), sd, p + ".cond_inj", g)
            x = resnet_block(x, t, sd, p + ".res_block", g)
            if blk.attn:
                x = self_attention(x, sd, p + ".attn", g)
        if taps is not None:
            taps[p] = x
    return _block(x, sd, "final_conv", g)


def count_flops(cfg: UNetCfg, hw: int = 64) -> float:
    """Conv+bmm+addmm FLOPs of one forward at hw x hw (2*MACs), for roofline arithmetic.

    Reproduces BASELINE.md's 8.378 GFLOP for the WV3 configuration."""
    C, P = cfg.lms_channel, cfg.pan_channel
    downs, mid, ups = topology(cfg)
    res = hw
    fl = 0.0

    def conv(ci, co, k, r):
        return 2.0 * r * r * ci * co * k * k

    def resblk(d, r):
        return 2 * conv(d, d, 3, r) + 2.0 * cfg.inner_channel * d

    def attn(d, r):
        n = r * r
        return conv(d, 3 * d, 1, r) + conv(d, d, 1, r) + 2 * (2.0 * n * n * d)

    in_ch = cfg.in_channel + (cfg.out_channel if cfg.self_condition else 0)
    for blk in downs:
        if blk.kind == "conv0":
            fl += conv(in_ch, cfg.inner_channel, 3, res)
        elif blk.kind == "down":
            res //= 2
            fl += conv(blk.dim, blk.dim, 3, res)
        else:
            d = blk.dim_out
            fl += conv(C + P, 4 * d, 3, res) + conv(4 * d, 2 * d, 1, res) + conv(blk.dim, d, 1, res)
            fl += resblk(d, res) + (attn(d, res) if blk.attn else 0)
    for blk in mid:
        fl += resblk(blk.dim, res) + (attn(blk.dim, res) if blk.attn else 0)
    for blk in ups:
        if blk.kind == "up":
            res *= 2
            fl += conv(blk.dim, blk.dim, 3, res)
        else:
            dim, o, cd = blk.dim, blk.dim_out, C + 3 * P
            n = res * res
            fl += 2.0 * n * dim * 9 + conv(dim, dim, 1, res)  # q.0 (depthwise), q.1
            fl += 2.0 * n * cd * 9 + conv(cd, 2 * dim, 1, res)  # kv.0, kv.1
            fl += 2 * (2.0 * n * dim * (dim // 8))  # context + apply
            fl += conv(dim, o, 1, res) + (conv(dim, o, 1, res) if dim != o else 0)
            fl += conv(o, 2 * o, 3, res) + conv(2 * o, o, 3, res) + conv(o, o, 1, res)
            fl += resblk(o, res) + (attn(o, res) if blk.attn else 0)
    fl += conv(cfg.inner_channel, cfg.out_channel, 3, res)
    fl += 2.0 * (cfg.inner_channel * 4 * cfg.inner_channel) * 2
    return fl
