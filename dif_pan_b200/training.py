"""Training-mode forward / backward of the SR3-DWT UNet — FIRST SLICE (BASELINE configs[4]).

The reference trains with plain autograd: `diffusion(res, cond=cond)` -> `p_losses` -> `loss.backward()` -> `clip_grad_norm_(0.003)` -> AdamW
-> EMA (/root/reference/diffusion/diffusion_ddpm_pan.py:692-766, /root/reference/diffusion_engine.py:219-248).  What this module provides in
round 2:

  * every DENSE convolution of the network (3x3 / 1x1, stride 1 / 2: 264 - 32 depthwise = 232 per forward, > 99 % of the FLOPs) runs on this
    repo's kernels in all three directions, wrapped as ONE `torch.autograd.Function` (`conv2d`):
        forward   tcgen05 implicit GEMM (csrc/conv3x3_halo.cu, csrc/gemm_tc.cu), bf16 operands, fp32 accumulation
        dgrad     the same kernels on the flipped / transposed weights (a stride-2 conv's data gradient = zero insertion + 3x3 conv)
        wgrad     csrc/backward.cu `wgrad_kernel` (split-K over pixels, ldmatrix.trans + mma.sync, fp32 atomics) and `colsum_kernel` (bias)
  * everything BETWEEN the convolutions (GroupNorm, Swish / SiLU, FiLM add, CSM modulation, the two FWM softmaxes and its d x d context
    products, the 64-token attention core, bilinear cond resize, nearest x2, Dropout / DropPath, depthwise 3x3) is expressed with torch
    tensor ops in fp32 in this slice, so their backward comes from autograd.  They are < 1 % of the FLOPs but most of the launches; hand-written
    fused forward+backward kernels for them (the inference path already fuses their forward) are the next slice.

The inference path does not go through this module and stays free of ATen arithmetic.  Numerics: convolution operands and their incoming
gradients are rounded to bf16 (fp32 accumulation), the residual stream / normalisations / gradients in between stay fp32 -- i.e. the mixed
precision `torch.autocast(bfloat16)` would give the reference, with this repo's kernels in place of cuDNN.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

from . import _lib
from .unet import _ceil, _pack_conv


def _nhwc_bf16(x: torch.Tensor, c_pad: int) -> torch.Tensor:
    """logical [B, C, H, W] (any layout, fp32 / bf16) -> contiguous [B, H, W, c_pad] bf16, zero padded channels."""
    B, C, H, W = x.shape
    if C == c_pad:
        return x.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
    out = torch.zeros(B, H, W, c_pad, dtype=torch.bfloat16, device=x.device)
    out[..., :C] = x.permute(0, 2, 3, 1)
    return out


def _launch_gemm(xa: torch.Tensor, wp: torch.Tensor, n_valid: int, taps: int, stride: int, bias: Optional[torch.Tensor], out: torch.Tensor,
                 per_sample: bool = False) -> None:
    """DDIF_OP_GEMM on NHWC bf16 tensors (include/ddif_b200.h): out[B, oh, ow, n_pad] = conv(xa[B, H, W, Cp], wp[taps | B][n_pad][Cp]) (+ bias)."""
    B, H, W, Cp = xa.shape
    _lib.launch("ddif_gemm_t", _lib.stream_of(xa.device), a=[xa.data_ptr(), None], a_ld=[Cp, 0], a_c=[Cp, 0], a_h=[H, 0], a_w=[W, 0],
                w=[wp.data_ptr(), None], w_s=[wp.shape[0], 0], w_k=[wp.shape[2], 0], taps=[taps, 0], w_per_sample=[1 if per_sample else 0, 0], nseg=1,
                stride=stride, batch=B, out_h=out.shape[1], out_w=out.shape[2], n_pad=out.shape[3], n_valid=n_valid,
                bias=bias.data_ptr() if bias is not None else None, film=None, film_ld=0, mod=None, residual=None, res_ld=0, act=0,
                out=out.data_ptr(), out_ld=out.shape[3], out_nchw=None, stats=None, gn_stats=None, gn_gamma=None, gn_beta=None, gn_eps=1e-5,
                gn_act=0, a_up=0, force_tma=0, gn_stats2=None, dw_w=None, dw_n=0)


class _Conv2d(torch.autograd.Function):
    """F.conv2d(x, w, b, stride, padding = k // 2) for k in {1, 3}: forward, data gradient and weight gradient on this repo's CUDA kernels."""

    @staticmethod
    def forward(ctx, x, w, b, stride):
        if not x.is_cuda:
            raise RuntimeError("dif_pan_b200.training runs on CUDA only (no CPU fallback)")
        B, C, H, W = x.shape
        O, I, kh, kw = w.shape
        if I != C or kh != kw or kh not in (1, 3) or stride not in (1, 2) or H % stride or W % stride:
            raise ValueError(f"conv2d: unsupported shape x {tuple(x.shape)} w {tuple(w.shape)} stride {stride}")
        Cp, Op = _ceil(C, 16), _ceil(O, 16)
        xa = _nhwc_bf16(x.detach(), Cp)
        wp = _pack_conv(w.detach(), Cp)
        bias = b.detach().to(torch.float32).contiguous() if b is not None else None
        out = torch.zeros(B, H // stride, W // stride, Op, dtype=torch.bfloat16, device=x.device)
        _launch_gemm(xa, wp, O, kh * kw, stride, bias, out)
        ctx.save_for_backward(xa, w)
        ctx.meta = (C, O, kh, stride, b is not None, H, W)
        return out[..., :O].permute(0, 3, 1, 2).float()

    @staticmethod
    def backward(ctx, gy):
        xa, w = ctx.saved_tensors
        C, O, k, stride, has_bias, H, W = ctx.meta
        B, Cp, Op, taps = xa.shape[0], xa.shape[3], _ceil(O, 16), k * k
        dev = xa.device
        ga = _nhwc_bf16(gy, Op)                                       # [B, oh, ow, Op] bf16
        st = _lib.stream_of(dev)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            # data gradient = convolution of dY with the flipped, transposed weights; stride 2: zero insertion first
            wt = _pack_conv(w.detach().flip(2, 3).permute(1, 0, 2, 3), Op)  # [taps][Cp][Op]
            src = ga
            if stride == 2:
                src = torch.zeros(B, H, W, Op, dtype=torch.bfloat16, device=dev)
                src[:, ::2, ::2] = ga
            dxa = torch.zeros(B, H, W, Cp, dtype=torch.bfloat16, device=dev)
            _launch_gemm(src, wt, C, taps, 1, None, dxa)
            gx = dxa[..., :C].permute(0, 3, 1, 2).float()
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(taps, Op, Cp, dtype=torch.float32, device=dev)
            _lib.launch("ddif_wgrad_t", st, x=xa.data_ptr(), x_ld=Cp, cin=Cp, dy=ga.data_ptr(), dy_ld=Op, cout=Op, dw=dw.data_ptr(), batch=B, in_h=H, in_w=W,
                        out_h=ga.shape[1], out_w=ga.shape[2], taps=taps, stride=stride, per_sample=0)
            gw = dw[:, :O, :C].permute(1, 2, 0).reshape(O, C, k, k).contiguous()
        if has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(Op, dtype=torch.float32, device=dev)
            _lib.launch("ddif_colsum_t", st, dy=ga.data_ptr(), out=db.data_ptr(), batch=B, hw=ga.shape[1] * ga.shape[2], c=Op, ld=Op, per_sample=0)
            gb = db[:O]
        return gx, gw, gb, None


def conv2d(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None, stride: int = 1) -> torch.Tensor:
    """Dense convolution (kernel 1 or 3, padding k // 2) on the CUDA kernels, differentiable in x, w and b."""
    return _Conv2d.apply(x, w, b, stride)


# ----------------------------------------------------------------------------------------------------------------
# the network (parameters are read from a dif_pan_b200.UNetSR3; structure follows /root/reference/models/sr3_dwt.py)
# ----------------------------------------------------------------------------------------------------------------
def _swish(x):
    return x * torch.sigmoid(x)                                       # sr3_dwt.py:261-263


def _drop_path(x, p: float, training: bool):
    """timm DropPath(scale_by_keep=True): one Bernoulli(keep) draw per sample, divided by keep (sr3_dwt.py:534,576)."""
    if p == 0.0 or not training:
        return x
    keep = 1.0 - p
    mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
    if keep > 0.0:
        mask.div_(keep)
    return x * mask


def _gn(x, m):
    return F.group_norm(x, m.num_groups, m.weight, m.bias, m.eps)


def _block(x, blk, p_drop: float, training: bool):
    """Block = GroupNorm -> Swish -> Dropout -> Conv3x3 (sr3_dwt.py:288-300)."""
    h = _swish(_gn(x, blk.block[0]))
    if p_drop:
        h = F.dropout(h, p_drop, training)
    c = blk.block[3]
    return conv2d(h, c.weight, c.bias)


def _res_block(x, t_emb, rb, dropout: float, training: bool):
    """ResnetBlock with additive FiLM (sr3_dwt.py:241-258,303-327): dropout only in block2."""
    h = _block(x, rb.block1, 0.0, training)
    lin = rb.noise_func.noise_func[0]
    h = h + F.linear(t_emb, lin.weight, lin.bias)[:, :, None, None]
    h = _block(h, rb.block2, dropout, training)
    if hasattr(rb, "res_conv"):
        x = conv2d(x, rb.res_conv.weight, rb.res_conv.bias)
    return h + x


def _self_attention(x, a, n_head: int):
    """SelfAttention (sr3_dwt.py:330-360); the scale uses the full channel count (:352)."""
    b, c, h, w = x.shape
    hd = c // n_head
    qkv = conv2d(_gn(x, a.norm), a.qkv.weight, None).reshape(b, n_head, hd * 3, h * w)
    q, k, v = qkv[:, :, :hd], qkv[:, :, hd:2 * hd], qkv[:, :, 2 * hd:]
    att = torch.softmax(torch.einsum("bncq,bnck->bnqk", q, k) / math.sqrt(c), dim=-1)
    out = torch.einsum("bnqk,bnck->bncq", att, v).reshape(b, c, h, w)
    return conv2d(out, a.out.weight, a.out.bias) + x


def _csm(x, c, ci):
    """CondInjection (sr3_dwt.py:376-396): x_conv(x) * (1 + scale) + shift, (scale, shift) = body(cond)."""
    h = conv2d(c, ci.body[0].weight, None)
    h = F.silu(_gn(h, ci.body[1]))
    scale, shift = conv2d(h, ci.body[3].weight, ci.body[3].bias).chunk(2, dim=1)
    return conv2d(x, ci.x_conv.weight, ci.x_conv.bias) * (1 + scale) + shift


def _fwm(x, c, ci, n_head: int, drop_path: float, training: bool):
    """FastAttnCondInjection (sr3_dwt.py:493-577)."""
    xh = _gn(x, ci.prenorm_x)
    dim = xh.shape[1]
    q = conv2d(F.conv2d(xh, ci.q[0].weight, None, padding=1, groups=dim), ci.q[1].weight, ci.q[1].bias)
    cd = c.shape[1]
    k, v = conv2d(F.conv2d(c, ci.kv[0].weight, None, padding=1, groups=cd), ci.kv[1].weight, ci.kv[1].bias).chunk(2, dim=1)
    q = q.softmax(dim=-2)
    k = k.softmax(dim=-1)
    b, _, h, w = q.shape
    d = dim // n_head
    q = q.reshape(b, n_head, d, h * w) * (1.0 / math.sqrt(d))
    ctx = torch.einsum("bhdn,bhen->bhde", k.reshape(b, n_head, d, h * w), v.reshape(b, n_head, d, h * w))
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, dim, h, w)
    y = conv2d(out, ci.attn_out.weight, ci.attn_out.bias)
    y = y + (conv2d(xh, ci.attn_res.weight, ci.attn_res.bias) if hasattr(ci, "attn_res") else xh)
    f = conv2d(F.silu(conv2d(y, ci.ffn[0].weight, None)), ci.ffn[2].weight, None)
    f = conv2d(f, ci.ffn[3].weight, ci.ffn[3].bias)
    return _drop_path(f, drop_path, training) + y


def _time_embedding(net, time: torch.Tensor) -> torch.Tensor:
    """PositionalEncoding + noise_level_mlp (sr3_dwt.py:57-64,223-238); `step` in the dtype of `time` like the reference."""
    count = net.inner_channel // 2
    step = torch.arange(count, dtype=time.dtype, device=time.device) / count
    enc = time.unsqueeze(1) * torch.exp(-math.log(1e4) * step.unsqueeze(0))
    enc = torch.cat([torch.sin(enc), torch.cos(enc)], dim=-1)
    mlp = net.noise_level_mlp
    return F.linear(_swish(F.linear(enc, mlp[1].weight, mlp[1].bias)), mlp[3].weight, mlp[3].bias)


FWM_DROP_PATH = 0.2  # FastAttnCondInjection's default drop_path_prob (sr3_dwt.py:493-505)


def unet_forward(net, x: torch.Tensor, time: torch.Tensor, cond: torch.Tensor, self_cond: Optional[torch.Tensor] = None) -> torch.Tensor:
    """UNetSR3.forward (sr3_dwt.py:169-219) with an autograd graph; Dropout / DropPath are active iff `net.training`."""
    training, p_drop = net.training, float(net.dropout)
    C, P = net.lms_channel, net.pan_channel
    x = x.to(torch.float32)
    if net.self_condition:
        x = torch.cat([x if self_cond is None else self_cond.to(torch.float32), x], dim=1)
    x = x.contiguous(memory_format=torch.channels_last)
    t = _time_embedding(net, time)
    c_enc, c_dec = cond[:, :C + P], cond[:, -(C + 3 * P):]
    resize = lambda c, like: c if c.shape[-2:] == like.shape[-2:] else F.interpolate(c, size=like.shape[-2:], mode="bilinear")
    feats = []
    for i, m in enumerate(net.downs):
        kind = getattr(m, "kind", None)
        if i == 0:
            x = conv2d(x, m.weight, m.bias)
        elif kind == "down":
            x = conv2d(x, m.conv.weight, m.conv.bias, stride=2)
        else:
            x = _csm(x, resize(c_enc, x), m.cond_inj)
            x = _res_block(x, t, m.res_block, p_drop, training)
            if m.with_attn:
                x = _self_attention(x, m.attn, net.N_HEADS)
        feats.append(x)
    for m in net.mid:
        x = _res_block(x, t, m.res_block, p_drop, training)
        if m.with_attn:
            x = _self_attention(x, m.attn, net.N_HEADS)
    for m in net.ups:
        if getattr(m, "kind", None) == "up":
            x = conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), m.conv.weight, m.conv.bias)
            continue
        x = torch.cat((x, feats.pop()), dim=1)
        x = _fwm(x, resize(c_dec, x), m.cond_inj, net.N_HEADS, FWM_DROP_PATH, training)
        x = _res_block(x, t, m.res_block, p_drop, training)
        if m.with_attn:
            x = _self_attention(x, m.attn, net.N_HEADS)
    return _block(x, net.final_conv, 0.0, training)


# ----------------------------------------------------------------------------------------------------------------
# whole-iteration CUDA graphs
# ----------------------------------------------------------------------------------------------------------------
class GraphedLossStep:
    """`loss = p_losses(x0, cond); loss.backward()` of a `GaussianDiffusion` captured into CUDA graphs for a fixed batch shape.

    One training iteration of this slice issues ~4 000 small launches (232 convolutions x (forward + dgrad + wgrad + weight repacking) plus the
    autograd ops between them) and is bound by the host at batch 32 (154 ms per iteration against ~25 ms of GPU work); replaying a captured
    graph removes the host from the loop -- the same remedy the inference path uses for its denoise step.  Two graphs are captured, with and
    without the no-grad self-conditioning pass (the reference draws `random.random() < 0.5` on the host per iteration,
    /root/reference/diffusion/diffusion_ddpm_pan.py:702).  Randomness inside the graphs (t, the q_sample noise, Dropout / DropPath masks) comes
    from torch's graph-safe CUDA generator, so every replay draws fresh numbers.  Gradients land in the parameters' `.grad` (allocate them
    first, e.g. through `ddp.GradientAllReducer`, whose bucket views are then written in place); gradient hooks do not run at replay, so
    reduce the buckets after `run()` (`GradientAllReducer.finish()` does that for buckets whose hooks did not fire).
    """

    def __init__(self, diffusion, x0: torch.Tensor, cond: torch.Tensor, warmup: int = 2):
        import random as _random
        self.dif, self._random = diffusion, _random
        self.x0, self.cond = x0.clone(), cond.clone()
        self.loss = {}
        self.graphs = {}
        params = [p for p in diffusion.model.parameters() if p.requires_grad]
        for p in params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        dev = x0.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for sc in (False, True):
                for _ in range(warmup):
                    self._iteration(sc)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for sc in (False, True):
            for p in params:
                p.grad.zero_()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.loss[sc] = self._iteration(sc)
            self.graphs[sc] = g
        for p in params:
            p.grad.zero_()

    def _iteration(self, self_cond: bool) -> torch.Tensor:
        b = self.x0.shape[0]
        t = torch.randint(0, self.dif.num_timesteps, (b,), device=self.x0.device)
        noise = torch.randn_like(self.x0)
        with torch.enable_grad():
            loss, _ = self.dif.p_losses(self.x0, noise=noise, cond=self.cond, t=t, self_cond_draw=0.0 if self_cond else 1.0)
            loss.backward()
        return loss.detach()

    def run(self, x0: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
        """Copies the batch into the graphs' static buffers, replays one iteration (gradients are ACCUMULATED into `.grad`: zero them
        first) and returns the (device) loss."""
        self.x0.copy_(x0)
        self.cond.copy_(cond)
        sc = self.dif.self_condition and self._random.random() < 0.5
        self.graphs[sc].replay()
        return self.loss[sc]
