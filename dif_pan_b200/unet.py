"""UNetSR3 — drop-in for /root/reference/models/sr3_dwt.py:30-219 driven by sm_100a CUDA kernels.

Same constructor keywords, same parameter names / shapes (`state_dict()` of the reference loads unchanged, 702
entries for the production configuration), same `forward(x, time, cond=None, self_cond=None)` contract (fp32 NCHW
in, new fp32 NCHW tensor out, inputs never mutated).  Underneath, the forward is a recorded plan of ~280 kernel
launches (include/ddif_b200.h) over NHWC bf16 activations:

  * every dense conv -> tcgen05/TMEM/TMA implicit GEMM with fused epilogue (bias, FiLM, CSM modulation, residual,
    SiLU, GroupNorm statistics)                                                            [csrc/gemm_tc.cu]
  * GroupNorm(1 group)+Swish, depthwise 3x3, FWM softmaxes, self-attention core, time MLP [csrc/elementwise.cu]
  * everything that depends only on `cond` (CSM scale/shift, FWM context folded into per-sample attn_out weights,
    the 4-level cond pyramid) is computed ONCE per `cond` and cached.

There is no PyTorch / CPU fallback: without the CUDA library or a CUDA device, forward raises.
In train() mode, or when an autograd graph is requested, forward goes through training.py (first slice of the training path).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _lib
from .plan import Buf, PlanBuilder


def _ceil(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ----------------------------------------------------------------------------------------------------------------
# parameter containers (structure == state_dict names of the reference; no forward of their own)
# ----------------------------------------------------------------------------------------------------------------
class _Box(nn.Module):
    """Plain container: holds parameters under the reference's attribute names."""


class _Marker(nn.Module):
    """Parameter-less placeholder keeping nn.Sequential indices aligned with the reference (e.g. Swish, SiLU)."""


def _seq(*mods) -> nn.Sequential:
    return nn.Sequential(*mods)


def _block(dim, dim_out, groups):
    b = _Box()
    b.block = _seq(nn.GroupNorm(groups, dim), _Marker(), _Marker(), nn.Conv2d(dim, dim_out, 3, padding=1))
    return b


def _res_block(dim, dim_out, emb_dim, groups):
    r = _Box()
    r.noise_func = _Box()
    r.noise_func.noise_func = _seq(nn.Linear(emb_dim, dim_out))
    r.block1 = _block(dim, dim_out, groups)
    r.block2 = _block(dim_out, dim_out, groups)
    if dim != dim_out:
        r.res_conv = nn.Conv2d(dim, dim_out, 1)
    return r


def _attn(dim, groups):
    a = _Box()
    a.norm = nn.GroupNorm(groups, dim)
    a.qkv = nn.Conv2d(dim, dim * 3, 1, bias=False)
    a.out = nn.Conv2d(dim, dim, 1)
    return a


def _csm(fea_dim, cond_dim, hidden, groups):
    c = _Box()
    c.body = _seq(nn.Conv2d(cond_dim, hidden * 4, 3, padding=1, bias=False), nn.GroupNorm(groups, hidden * 4), _Marker(),
                  nn.Conv2d(hidden * 4, hidden * 2, 1, bias=True))
    c.x_conv = nn.Conv2d(fea_dim, hidden, 1, bias=True)
    nn.init.zeros_(c.body[-1].weight)  # sr3_dwt.py:386-387
    nn.init.zeros_(c.body[-1].bias)
    return c


def _fwm(fea_dim, cond_dim, qkv_dim, dim_out, groups):
    f = _Box()
    f.prenorm_x = nn.GroupNorm(groups, fea_dim)
    f.q = _seq(nn.Conv2d(fea_dim, fea_dim, 3, 1, 1, bias=False, groups=fea_dim), nn.Conv2d(fea_dim, qkv_dim, 1, bias=True))
    f.kv = _seq(nn.Conv2d(cond_dim, cond_dim, 3, 1, 1, bias=False, groups=cond_dim), nn.Conv2d(cond_dim, qkv_dim * 2, 1, bias=True))
    f.attn_out = nn.Conv2d(qkv_dim, dim_out, 1, bias=True)
    if fea_dim != dim_out:
        f.attn_res = nn.Conv2d(fea_dim, dim_out, 1, bias=True)
    f.ffn = _seq(nn.Conv2d(dim_out, dim_out * 2, 3, 1, 1, bias=False), _Marker(), nn.Conv2d(dim_out * 2, dim_out, 3, 1, 1, bias=False),
                 nn.Conv2d(dim_out, dim_out, 1, bias=True))
    return f


def _stage(dim, dim_out, *, cond_dim, emb_dim, groups, with_attn, encoder):
    s = _Box()
    s.kind = "enc" if (cond_dim is not None and encoder) else ("dec" if cond_dim is not None else "mid")
    s.dim, s.dim_out, s.with_attn = dim, dim_out, with_attn
    s.res_block = _res_block(dim_out if cond_dim is not None else dim, dim_out, emb_dim, groups)
    if with_attn:
        s.attn = _attn(dim_out, groups)
    if cond_dim is not None:
        s.cond_inj = _csm(dim, cond_dim, dim_out, groups) if encoder else _fwm(dim, cond_dim, dim, dim_out, groups)
    return s


def _resample(dim, kind):
    r = _Box()
    r.kind = kind
    r.dim = r.dim_out = dim
    r.conv = nn.Conv2d(dim, dim, 3, 2 if kind == "down" else 1, 1)
    return r


# ----------------------------------------------------------------------------------------------------------------
class UNetSR3(nn.Module):
    N_HEADS = 8

    def __init__(self, in_channel=8, out_channel=3, inner_channel=32, lms_channel=8, pan_channel=1, norm_groups=32,
                 channel_mults=(1, 2, 4, 8, 8), attn_res=(8,), res_blocks=3, dropout=0, with_noise_level_emb=True,
                 image_size=128, self_condition=False, fourier_features=False, fourier_min=7, fourier_max=8,
                 fourier_step=1, pred_var=False):
        super().__init__()
        if fourier_features or pred_var:
            raise NotImplementedError("fourier_features / pred_var are never enabled by the reference engine "
                                      "(diffusion_ddpm_pan.py:184 asserts pred_var == False)")
        if not with_noise_level_emb:
            raise NotImplementedError("with_noise_level_emb=False is not used by the reference engine")
        if norm_groups != 1:
            raise NotImplementedError("the CUDA path implements norm_groups=1 (diffusion_engine.py:127,387)")
        self.lms_channel, self.pan_channel = lms_channel, pan_channel
        self.in_channel, self.out_channel, self.inner_channel = in_channel, out_channel, inner_channel
        self.res_blocks, self.self_condition, self.pred_var = res_blocks, self_condition, pred_var
        self.fourier_features = False
        self.image_size, self.dropout = image_size, dropout
        ic = inner_channel
        self.noise_level_mlp = _seq(_Marker(), nn.Linear(ic, ic * 4), _Marker(), nn.Linear(ic * 4, ic))
        cin = in_channel + (out_channel if self_condition else 0)
        downs: List[nn.Module] = [nn.Conv2d(cin, ic, 3, padding=1)]
        pre, feat, res = ic, [ic], image_size
        n = len(channel_mults)
        for lvl in range(n):
            ch = ic * channel_mults[lvl]
            for _ in range(res_blocks):
                downs.append(_stage(pre, ch, cond_dim=lms_channel + pan_channel, emb_dim=ic, groups=norm_groups,
                                    with_attn=res in attn_res, encoder=True))
                feat.append(ch)
                pre = ch
            if lvl != n - 1:
                downs.append(_resample(pre, "down"))
                feat.append(pre)
                res //= 2
        self.downs = nn.ModuleList(downs)
        self.mid = nn.ModuleList([
            _stage(pre, pre, cond_dim=None, emb_dim=ic, groups=norm_groups, with_attn=True, encoder=True),
            _stage(pre, pre, cond_dim=None, emb_dim=ic, groups=norm_groups, with_attn=False, encoder=True)])
        ups: List[nn.Module] = []
        for lvl in reversed(range(n)):
            ch = ic * channel_mults[lvl]
            for _ in range(res_blocks + 1):
                ups.append(_stage(pre + feat.pop(), ch, cond_dim=lms_channel + pan_channel * 3, emb_dim=ic,
                                  groups=norm_groups, with_attn=res in attn_res, encoder=False))
                pre = ch
            if lvl >= 1:
                ups.append(_resample(pre, "up"))
                res *= 2
        self.ups = nn.ModuleList(ups)
        self.final_conv = _block(pre, out_channel if out_channel is not None else in_channel, norm_groups)
        # runtime state (not parameters)
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._packed_key = None
        self._rt: Optional["_Runtime"] = None

    # -- weights ---------------------------------------------------------------------------------------------
    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def pack_weights(self, force: bool = False) -> Dict[str, torch.Tensor]:
        """One-time repack of the fp32 OIHW parameters into kernel layouts (bf16 [tap][Cout_pad][Cin_pad] etc.)."""
        key = self._weights_key()
        if not force and self._packed is not None and key == self._packed_key:
            return self._packed
        sd = {k: v.detach() for k, v in self.state_dict().items()}
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("dif_pan_b200.UNetSR3 runs on CUDA only (no CPU fallback): call .cuda() first")
        self._packed = pack_state_dict(sd, self)
        self._packed_key = key
        if self._rt is not None:  # the recorded plans / captured graph point at the old packs
            self._rt.close()
        self._rt = None
        return self._packed

    # -- runtime ---------------------------------------------------------------------------------------------
    def runtime(self, batch: int, height: int, width: int) -> "_Runtime":
        self.pack_weights()
        rt = self._rt
        if rt is None or (rt.B, rt.H, rt.W) != (batch, height, width):
            if rt is not None:
                rt.close()
            rt = _Runtime(self, batch, height, width)
            self._rt = rt
        return rt

    def forward(self, x, time, cond=None, self_cond=None):
        if cond is None:
            raise ValueError("UNetSR3.forward needs cond=[lms, pan, wavelets] (sr3_dwt.py:197,214)")
        if not x.is_cuda:
            raise RuntimeError("dif_pan_b200.UNetSR3 has no CPU fallback; inputs must be CUDA tensors")
        if self.training or (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())):
            # train() mode (Dropout / DropPath active, like the reference's nn.Module) or an autograd graph is wanted: the training path
            # (training.py: convolutions forward / dgrad / wgrad on the CUDA kernels).  Sampling = .eval() under torch.no_grad().
            from . import training
            return training.unet_forward(self, x, time, cond, self_cond)
        B, C, H, W = x.shape
        rt = self.runtime(B, H, W)
        rt.set_cond(cond)
        rt.x_buf.copy_(x)
        rt.t_buf.copy_(time.reshape(-1).to(torch.float32).expand(B))
        if self.self_condition and self_cond is not None:
            rt.sc_buf.copy_(self_cond)
            rt.step(explicit_self_cond=True)
        else:
            rt.step()
        return rt.out_buf.clone()


# ----------------------------------------------------------------------------------------------------------------
# weight packing (torch ops, one-time plumbing)
# ----------------------------------------------------------------------------------------------------------------
def _pack_conv(w: torch.Tensor, cin_pad: Optional[int] = None) -> torch.Tensor:
    """OIHW fp32 -> [taps][O_pad16][I_pad] bf16 (K-major rows, tap = ky*3+kx)."""
    O, I, kh, kw = w.shape
    ip = cin_pad if cin_pad is not None else _ceil(I, 16)
    out = torch.zeros(kh * kw, _ceil(O, 16), ip, dtype=torch.bfloat16, device=w.device)
    out[:, :O, :I] = w.permute(2, 3, 0, 1).reshape(kh * kw, O, I).to(torch.bfloat16)
    return out.contiguous()


def pack_state_dict(sd: Dict[str, torch.Tensor], net: "UNetSR3") -> Dict[str, torch.Tensor]:
    P: Dict[str, torch.Tensor] = {}
    f32 = lambda t: t.to(torch.float32).contiguous()
    for k in ("noise_level_mlp.1.weight", "noise_level_mlp.1.bias", "noise_level_mlp.3.weight", "noise_level_mlp.3.bias"):
        P[k] = f32(sd[k])
    film_w, film_b, off = [], [], 0
    P["_film_offsets"] = {}

    def resblock(p):
        nonlocal off
        w, b = sd[p + ".noise_func.noise_func.0.weight"], sd[p + ".noise_func.noise_func.0.bias"]
        P["_film_offsets"][p] = off
        off += w.shape[0]
        film_w.append(w)
        # block1's conv bias is folded into the FiLM vector (both are per-(sample,)channel additive terms): one add less
        # per output element in the conv epilogue
        film_b.append(b + sd[p + ".block1.block.3.bias"])
        for blk in ("block1", "block2"):
            P[f"{p}.{blk}.gamma"] = f32(sd[f"{p}.{blk}.block.0.weight"])
            P[f"{p}.{blk}.beta"] = f32(sd[f"{p}.{blk}.block.0.bias"])
            P[f"{p}.{blk}.w"] = _pack_conv(sd[f"{p}.{blk}.block.3.weight"])
            P[f"{p}.{blk}.b"] = f32(sd[f"{p}.{blk}.block.3.bias"])
        if (p + ".res_conv.weight") in sd:
            P[p + ".res_conv.w"] = _pack_conv(sd[p + ".res_conv.weight"])
            P[p + ".res_conv.b"] = f32(sd[p + ".res_conv.bias"])

    def attn(p):
        P[p + ".gamma"], P[p + ".beta"] = f32(sd[p + ".norm.weight"]), f32(sd[p + ".norm.bias"])
        P[p + ".qkv.w"] = _pack_conv(sd[p + ".qkv.weight"])
        P[p + ".out.w"] = _pack_conv(sd[p + ".out.weight"])
        P[p + ".out.b"] = f32(sd[p + ".out.bias"])
        C = sd[p + ".out.weight"].shape[1]
        if C == 128:  # fused attention block (csrc/attn_block.cu): out weight with its K axis permuted for 64-byte fragment loads
            pos = torch.arange(C, device=sd[p + ".out.weight"].device)
            j_, t_, w_, e_ = pos // 32, (pos % 32) // 8, (pos % 8) // 2, pos % 2   # chunk 4j + t of the row, word w of the chunk
            src = 16 * (2 * j_ + w_ // 2) + 8 * (w_ % 2) + 2 * t_ + e_
            P[p + ".out.wp"] = sd[p + ".out.weight"].to(torch.float32)[:, :, 0, 0][:, src].to(torch.bfloat16).contiguous()

    cin = net.in_channel + (net.out_channel if net.self_condition else 0)
    P["downs.0.w"] = _pack_conv(sd["downs.0.weight"], _ceil(cin, 16))
    P["downs.0.b"] = f32(sd["downs.0.bias"])
    ce_pad = _ceil(net.lms_channel + net.pan_channel, 16)
    for grp in ("downs", "mid", "ups"):
        for i, m in enumerate(getattr(net, grp)):
            p = f"{grp}.{i}"
            kind = getattr(m, "kind", None)
            if kind in ("down", "up"):
                P[p + ".w"] = _pack_conv(sd[p + ".conv.weight"])
                P[p + ".b"] = f32(sd[p + ".conv.bias"])
                continue
            if kind is None:
                continue
            resblock(p + ".res_block")
            if m.with_attn:
                attn(p + ".attn")
            q = p + ".cond_inj"
            if kind == "enc":
                P[q + ".body0.w"] = _pack_conv(sd[q + ".body.0.weight"], ce_pad)
                P[q + ".body.gamma"], P[q + ".body.beta"] = f32(sd[q + ".body.1.weight"]), f32(sd[q + ".body.1.bias"])
                P[q + ".body3.w"] = _pack_conv(sd[q + ".body.3.weight"])
                P[q + ".body3.b"] = f32(sd[q + ".body.3.bias"])
                P[q + ".x_conv.w"] = _pack_conv(sd[q + ".x_conv.weight"])
                P[q + ".x_conv.b"] = f32(sd[q + ".x_conv.bias"])
            elif kind == "dec":
                dim = m.dim
                P[q + ".gamma"], P[q + ".beta"] = f32(sd[q + ".prenorm_x.weight"]), f32(sd[q + ".prenorm_x.bias"])
                P[q + ".q0"] = f32(sd[q + ".q.0.weight"].reshape(dim, 9).t())  # [9][dim]
                P[q + ".q1.w"] = _pack_conv(sd[q + ".q.1.weight"])
                # q = Conv1x1(DW3x3(.)) composed once in fp32 into one dense 3x3 (sr3_dwt.py:509-512): W[o,c,ky,kx] = W1[o,c] * Wdw[c,ky,kx]
                P[q + ".qc.w"] = _pack_conv(torch.einsum("oc,ckl->ockl", sd[q + ".q.1.weight"].to(torch.float32)[:, :, 0, 0],
                                                          sd[q + ".q.0.weight"].to(torch.float32)[:, 0]))
                P[q + ".q1.b"] = f32(sd[q + ".q.1.bias"])
                if (q + ".attn_res.weight") in sd:
                    # [q ; attn_res] as ONE 3x3 conv over x_hat: rows dim.. hold attn_res (1x1 = centre tap only), so the
                    # q-path kernel also emits r = attn_res(x_hat) and x_hat itself never goes through HBM (sr3_dwt.py:573)
                    o_ = sd[q + ".attn_res.weight"].shape[0]
                    wq = torch.einsum("oc,ckl->ockl", sd[q + ".q.1.weight"].to(torch.float32)[:, :, 0, 0], sd[q + ".q.0.weight"].to(torch.float32)[:, 0])
                    wr = torch.zeros(o_, dim, 3, 3, dtype=torch.float32, device=wq.device)
                    wr[:, :, 1, 1] = sd[q + ".attn_res.weight"].to(torch.float32)[:, :, 0, 0]
                    P[q + ".qcr.w"] = _pack_conv(torch.cat([wq, wr], 0))
                    # depthwise mode of the same kernel: 1x1 weights [q.1 ; attn_res], the depthwise 3x3 runs inside the kernel
                    P[q + ".q1r.w"] = _pack_conv(torch.cat([sd[q + ".q.1.weight"].to(torch.float32), sd[q + ".attn_res.weight"].to(torch.float32)], 0))
                    P[q + ".qcr.b"] = f32(torch.cat([sd[q + ".q.1.bias"].to(torch.float32), torch.zeros(o_, dtype=torch.float32, device=wq.device)]))
                cd = sd[q + ".kv.0.weight"].shape[0]
                P[q + ".kv0"] = f32(sd[q + ".kv.0.weight"].reshape(cd, 9))
                P[q + ".kv1.w"] = f32(sd[q + ".kv.1.weight"].reshape(2 * dim, cd))
                P[q + ".kv1.b"] = f32(sd[q + ".kv.1.bias"])
                P[q + ".attn_out.w32"] = f32(sd[q + ".attn_out.weight"].reshape(m.dim_out, dim))
                bias = sd[q + ".attn_out.bias"].clone()
                if (q + ".attn_res.weight") in sd:
                    P[q + ".attn_res.w"] = _pack_conv(sd[q + ".attn_res.weight"])
                    bias = bias + sd[q + ".attn_res.bias"]
                P[q + ".attn.b"] = f32(bias)
                P[q + ".ffn0.w"] = _pack_conv(sd[q + ".ffn.0.weight"])
                # ffn.2 (3x3, no bias) followed by ffn.3 (1x1): composed once in fp32 into one 3x3 conv (sr3_dwt.py:531-532)
                w2 = sd[q + ".ffn.2.weight"].to(torch.float32)
                w3 = sd[q + ".ffn.3.weight"].to(torch.float32)[:, :, 0, 0]
                P[q + ".ffn23.w"] = _pack_conv(torch.einsum("om,mckl->ockl", w3, w2))
                P[q + ".ffn23.b"] = f32(sd[q + ".ffn.3.bias"])
    P["film.w"] = f32(torch.cat(film_w, 0))
    P["film.b"] = f32(torch.cat(film_b, 0))
    P["final.gamma"], P["final.beta"] = f32(sd["final_conv.block.0.weight"]), f32(sd["final_conv.block.0.bias"])
    P["final.w"] = _pack_conv(sd["final_conv.block.3.weight"])
    P["final.b"] = f32(sd["final_conv.block.3.bias"])
    return P


# ----------------------------------------------------------------------------------------------------------------
# schedule construction (pure bookkeeping; unit-tested on CPU with fake addresses)
# ----------------------------------------------------------------------------------------------------------------
class Act:
    """An NHWC bf16 activation inside a plan."""

    def __init__(self, buf, B, H, W, C, stats=None):
        self.buf, self.B, self.H, self.W, self.C, self.stats = buf, B, H, W, C, stats


class Schedule:
    """Builds the cond-cache plan and the per-step forward plan of a UNetSR3 for a fixed (B, H, W)."""

    def __init__(self, net: UNetSR3, addr: Dict[str, int], B: int, H: int, W: int, io: Dict[str, int], *, use_qconv: bool = True,
                 use_dwq: bool = True, use_attn_block: bool = True, use_cs_gemm: bool = True, use_fwm_front: bool = True):
        """addr: packed-weight name -> device address; io: x, sc, t, out, cond addresses (fixed buffers).
        use_qconv / use_dwq / use_attn_block / use_cs_gemm / use_fwm_front: A/B switches for tools/ (composed q conv, in-kernel depthwise q path, fused
        attention block, softmax-over-H fused into the attn_out GEMM, fused FWM front at 8x8); the product always builds the schedule with all of them on."""
        if H % 8 or W % 8 or min(H, W) < 8 * 2 ** (self._levels(net) - 1):
            raise ValueError(f"UNetSR3 needs H, W multiples of 8 and >= {8 * 2 ** (self._levels(net) - 1)}, got {H}x{W}")
        self.net, self.addr, self.B, self.H, self.W, self.io = net, addr, B, H, W, io
        self.fwd = PlanBuilder()
        self.cnd = PlanBuilder()
        self.cache = PlanBuilder()  # only used as a packer for the persistent cond-cache buffers
        self.n_stats = 0
        self.stats_buf = self.fwd.buf("stats", 0, persistent=True)
        self.cstats_buf = self.cnd.buf("cond_stats", 0, persistent=True)
        self.n_cstats = 0
        self.mod: Dict[str, Buf] = {}
        self.weff: Dict[str, Buf] = {}
        self.weff_k: Dict[str, int] = {}
        self.film_offsets = None
        self.first_body_op = 0
        self.use_qconv, self.use_dwq, self.use_attn_block, self.use_cs_gemm = use_qconv, use_dwq, use_attn_block, use_cs_gemm
        self.use_fwm_front = use_fwm_front

    @staticmethod
    def _levels(net) -> int:
        return 1 + sum(1 for m in net.downs if getattr(m, "kind", None) == "down")

    # -- helpers ------------------------------------------------------------------------------------------
    def _act(self, pb: PlanBuilder, name, B, H, W, C, stats=False) -> Act:
        buf = pb.buf(name, B * H * W * C * 2)
        st = None
        if stats:
            if pb is self.fwd:
                st = (self.stats_buf, self.n_stats * B * 16)
                self.n_stats += 1
            else:
                st = (self.cstats_buf, self.n_cstats * B * 16)
                self.n_cstats += 1
        return Act(buf, B, H, W, C, st)

    def _gemm(self, pb, label, srcs, weights, n_valid, out: Optional[Act], *, taps, stride=1, bias=None, film=None,
              film_ld=0, mod=None, residual: Optional[Act] = None, act=0, out_nchw=None, per_sample=(0, 0), w_s=None, out_hw=None,
              gn=None, a_up=0, w_k=None, ref_flops=-1.0, residual_off=0, dw=None, softmax_h=False, a_c=None):
        """gn = (gamma_addr, beta_addr, act): fuse GroupNorm(+Swish) of the source(s) into the conv's loader, using the
        sources' own statistics; a_up is reserved (must be 0)."""
        a0 = srcs[0]
        oh, ow = out_hw if out_hw else ((a0.H << a_up) // stride, (a0.W << a_up) // stride)
        n_pad = _ceil(n_valid, 16)
        nseg = len(srcs)
        k_total = sum((a_c[i] if a_c else s.C) * taps[i] for i, s in enumerate(srcs))
        fields = dict(
            a=[s.buf for s in srcs] + [None] * (2 - nseg), a_ld=[s.C for s in srcs] + [0] * (2 - nseg),
            a_c=(list(a_c) if a_c else [s.C for s in srcs]) + [0] * (2 - nseg), a_h=[s.H for s in srcs] + [0] * (2 - nseg),
            a_w=[s.W for s in srcs] + [0] * (2 - nseg), w=list(weights) + [None] * (2 - nseg),
            w_s=[(w_s[i] if w_s else taps[i]) for i in range(nseg)] + [0] * (2 - nseg),
            w_k=(list(w_k) if w_k else [s.C for s in srcs]) + [0] * (2 - nseg), taps=list(taps) + [0] * (2 - nseg),
            w_per_sample=list(per_sample)[:nseg] + [0] * (2 - nseg), nseg=nseg, stride=stride, batch=a0.B, out_h=oh, out_w=ow,
            n_pad=n_pad, n_valid=n_valid, bias=bias, film=film, film_ld=film_ld, mod=mod,
            residual=((residual.buf, residual_off * 2) if residual_off else residual.buf) if residual else None,
            res_ld=residual.C if residual else 0, act=act,
            out=out.buf if out else None, out_ld=out.C if out else 0, out_nchw=out_nchw,
            stats=out.stats if (out is not None and out.stats is not None) else None,
            gn_stats=a0.stats if gn else None, gn_gamma=gn[0] if gn else None, gn_beta=gn[1] if gn else None, gn_eps=1e-5,
            gn_act=gn[2] if gn else 0, a_up=a_up, force_tma=0,
            gn_stats2=srcs[1].stats if (gn and nseg == 2) else None, dw_w=dw[0] if dw else None, dw_n=dw[1] if dw else 0,
            a_softmax_h=1 if softmax_h else 0)
        if gn:
            assert all(s.stats is not None for s in srcs), label
        m = a0.B * oh * ow
        nbytes = sum(s.B * s.H * s.W * (a_c[i] if a_c else s.C) * 2 for i, s in enumerate(srcs)) + m * n_valid * (2 if out else 4)
        if residual:
            nbytes += m * n_valid * 2
        if mod is not None:
            nbytes += m * n_valid * 4
        flops = 2.0 * m * n_valid * k_total
        if dw:  # in-kernel depthwise 3x3 (CUDA cores) + ONE tap on the tensor cores
            cin = sum(s.C for s in srcs)
            flops = 2.0 * m * cin * (9 + n_valid)
        return pb.add("ddif_gemm_t", label=label, flops=flops, traffic=nbytes, ref_flops=ref_flops, **fields)

    def _gn(self, pb, label, src: Act, gamma, beta, act, src2: Optional[Act] = None, dw_w=None, name="gn"):
        C = src.C + (src2.C if src2 else 0)
        out = self._act(pb, name, src.B, src.H, src.W, C)
        out_dw = self._act(pb, name + "_dw", src.B, src.H, src.W, C) if dw_w is not None else None
        pb.add("ddif_gn_apply_t", label=label, traffic=src.B * src.H * src.W * C * 2 * (2 + (1 if dw_w else 0)),
               src1=src.buf, c1=src.C, src2=src2.buf if src2 else None, c2=src2.C if src2 else 0,
               stats1=src.stats, stats2=src2.stats if src2 else None, gamma=gamma, beta=beta, out=out.buf,
               dw_w=dw_w, out_dw=out_dw.buf if out_dw else None, batch=src.B, h=src.H, w=src.W, act=act, eps=1e-5)
        return out, out_dw

    @staticmethod
    def _fusable(x: Act) -> bool:
        """The halo kernel with the GN+Swish prologue tiles the image in 8 (x) x 16 (y) pixel boxes (csrc/conv3x3_halo.cu);
        at the 8x8 level half of every tile is padding, still cheaper than a stand-alone normalisation launch."""
        return x.W >= 8 and x.H >= 8 and x.C <= 256

    def _gn_conv3(self, label, x: Act, gamma, beta, w, n_valid, out, **kw):
        """Block = GN -> Swish -> Conv3x3 (sr3_dwt.py:288-300): one fused kernel where the tile shape allows, else
        gn_apply + the generic TMA conv."""
        pb = self.fwd
        if self._fusable(x):
            return self._gemm(pb, label + ".gn+conv", [x], [w], n_valid, out, taps=[9], gn=(gamma, beta, 1), **kw)
        n, _ = self._gn(pb, label + ".gn", x, gamma, beta, 1, name=label + ".n")
        return self._gemm(pb, label + ".conv", [n], [w], n_valid, out, taps=[9], **kw)

    def _res_block(self, x: Act, p: str) -> Act:
        A, pb = self.addr, self.fwd
        d = x.C
        h1 = self._act(pb, p + ".h1", x.B, x.H, x.W, d, stats=True)
        film_off = self.film_offsets[p]
        self._gn_conv3(p + ".block1", x, A[p + ".block1.gamma"], A[p + ".block1.beta"], A[p + ".block1.w"], d, h1,
                       film=(self.film_buf, film_off * 4), film_ld=self.nfilm)  # conv bias lives in the FiLM vector
        out = self._act(pb, p + ".out", x.B, x.H, x.W, d, stats=True)
        self._gn_conv3(p + ".block2", h1, A[p + ".block2.gamma"], A[p + ".block2.beta"], A[p + ".block2.w"], d, out,
                       bias=A[p + ".block2.b"], residual=x)
        return out

    def _attention(self, x: Act, p: str) -> Act:
        A, pb = self.addr, self.fwd
        C = x.C
        ntok = x.H * x.W
        if self.use_attn_block and ntok == 64 and C == 128 and self.net.N_HEADS == 8 and (p + ".out.wp") in A and x.stats is not None:
            # GN + qkv + 64-token attention + out + residual + statistics in ONE kernel, one CTA per sample (csrc/attn_block.cu)
            out = self._act(pb, p + ".out", x.B, x.H, x.W, C, stats=True)
            pb.add("ddif_attn_block_t", label=p + ".block", flops=2.0 * x.B * ntok * C * (3 * C + C) + 4.0 * x.B * ntok * ntok * C,
                   traffic=x.B * ntok * C * 4, x=x.buf, stats_in=x.stats, gamma=A[p + ".gamma"], beta=A[p + ".beta"], wqkv=A[p + ".qkv.w"],
                   wout=A[p + ".out.wp"], bout=A[p + ".out.b"], out=out.buf, stats_out=out.stats, batch=x.B, ntok=ntok, c=C,
                   heads=self.net.N_HEADS, scale=1.0 / math.sqrt(C), eps=1e-5)
            return out
        n, _ = self._gn(pb, p + ".norm", x, A[p + ".gamma"], A[p + ".beta"], 0, name=p + ".n")
        qkv = self._act(pb, p + ".qkv", x.B, x.H, x.W, 3 * C)
        self._gemm(pb, p + ".qkv", [n], [A[p + ".qkv.w"]], 3 * C, qkv, taps=[1])
        a = self._act(pb, p + ".a", x.B, x.H, x.W, C)
        pb.add("ddif_attn_t", label=p + ".core", flops=4.0 * x.B * ntok * ntok * C, traffic=x.B * ntok * C * 8,
               qkv=qkv.buf, out=a.buf, batch=x.B, ntok=ntok, c=C, heads=self.net.N_HEADS, scale=1.0 / math.sqrt(C))
        out = self._act(pb, p + ".out", x.B, x.H, x.W, C, stats=True)
        self._gemm(pb, p + ".out", [a], [A[p + ".out.w"]], C, out, taps=[1], bias=A[p + ".out.b"], residual=x)
        return out

    # -- cond-only plan -----------------------------------------------------------------------------------
    def build_cond(self) -> None:
        net, A, pb, B = self.net, self.addr, self.cnd, self.B
        C, P = net.lms_channel, net.pan_channel
        ce, cd = C + P, C + 3 * P
        ce_pad = _ceil(ce, 16)
        ctot = 2 * C + 4 * P
        levels = self._levels(net)
        pb.add("ddif_memset_t", label="cond.stats0", ptr=self.cstats_buf, bytes=0)  # size patched in finish()
        cenc, cdec = [], []
        for l in range(levels):
            h, w = self.H >> l, self.W >> l
            e = Act(pb.buf(f"cenc{l}", B * h * w * ce_pad * 2), B, h, w, ce_pad)
            dbuf = pb.buf(f"cdec{l}", B * cd * h * w * 4)
            pb.add("ddif_resize_t", label=f"cond.enc{l}", traffic=B * h * w * ce * 6, src=self.io["cond"], batch=B, c_total=ctot, c0=0, c=ce,
                   h=self.H, w=self.W, out_h=h, out_w=w, dst_nhwc=e.buf, c_pad=ce_pad, dst_nchw=None)
            pb.add("ddif_resize_t", label=f"cond.dec{l}", traffic=B * h * w * cd * 8, src=self.io["cond"], batch=B, c_total=ctot, c0=ctot - cd, c=cd,
                   h=self.H, w=self.W, out_h=h, out_w=w, dst_nhwc=None, c_pad=0, dst_nchw=dbuf)
            cenc.append(e)
            cdec.append((dbuf, h, w))
        lvl = 0
        for i, m in enumerate(net.downs):
            kind = getattr(m, "kind", None)
            if kind == "down":
                lvl += 1
            if kind != "enc":
                continue
            p = f"downs.{i}.cond_inj"
            d = m.dim_out
            e = cenc[lvl]
            hb = self._act(pb, p + ".hb", B, e.H, e.W, 4 * d, stats=True)
            self._gemm(pb, p + ".body0", [e], [A[p + ".body0.w"]], 4 * d, hb, taps=[9])
            hn, _ = self._gn(pb, p + ".body.gn", hb, A[p + ".body.gamma"], A[p + ".body.beta"], 1, name=p + ".hn")
            mod = self.cache.buf(p + ".mod", B * e.H * e.W * 2 * d * 2, persistent=True)
            self.mod[f"downs.{i}"] = mod
            self._gemm(pb, p + ".body3", [hn], [A[p + ".body3.w"]], 2 * d, Act(("cache", mod), B, e.H, e.W, 2 * d), taps=[1], bias=A[p + ".body3.b"])
        for i, m in enumerate(net.ups):
            kind = getattr(m, "kind", None)
            if kind == "up":
                lvl -= 1
            if kind != "dec":
                continue
            p = f"ups.{i}.cond_inj"
            dim, o = m.dim, m.dim_out
            dbuf, h, w = cdec[lvl]
            dh = dim // net.N_HEADS
            ctx = pb.buf(p + ".ctx", B * dim * dh * 4)
            pb.add("ddif_fwm_context_t", label=p + ".context", flops=2.0 * B * h * w * (cd * 9 + 2 * dim * cd + dim * dh),
                   traffic=B * h * w * cd * 4, c_dec=dbuf, kv0_w=A[p + ".kv0"], kv1_w=A[p + ".kv1.w"], kv1_b=A[p + ".kv1.b"],
                   ctx=ctx, batch=B, h=h, w=w, cd=cd, dim=dim, heads=net.N_HEADS)
            o_pad = _ceil(o, 16)
            # K padded to a multiple of 64 (dim 96 -> 128; the cache arena is zero-initialised and these columns are never written): the
            # column-softmax GEMM then reads the [q | attn_res] tensor as two 64-channel slabs instead of three 32-channel ones -- the padding
            # channels are the first attn_res channels of the same pixel, multiplied by zero weights
            k_pad = _ceil(dim, 64) if _ceil(dim, 64) <= dim + o else dim
            weff = self.cache.buf(p + ".weff", B * o_pad * k_pad * 2, persistent=True)
            self.weff[f"ups.{i}"] = weff
            self.weff_k[f"ups.{i}"] = k_pad
            pb.add("ddif_fwm_weff_t", label=p + ".weff", flops=2.0 * B * o * dim * dh, traffic=B * o * dim * 2, ctx=ctx,
                   w_out=A[p + ".attn_out.w32"], weff=("cache", weff), batch=B, o=o, dim=dim, heads=net.N_HEADS, o_pad=o_pad,
                   k_pad=k_pad, scale=1.0 / math.sqrt(dh))
        self.cstats_buf.nbytes = max(self.n_cstats * B * 16, 16)
        pb.ops[0].fields["bytes"] = self.cstats_buf.nbytes

    # -- per-step forward plan ----------------------------------------------------------------------------
    def build_forward(self, film_offsets: Dict[str, int], nfilm: int) -> None:
        net, A, pb, B, H, W = self.net, self.addr, self.fwd, self.B, self.H, self.W
        self.film_offsets, self.nfilm = film_offsets, nfilm
        ic = net.inner_channel
        self.film_buf = pb.buf("film", B * nfilm * 4)
        pb.add("ddif_memset_t", label="stats0", ptr=self.stats_buf, bytes=0)  # size patched below
        pb.add("ddif_time_embed_t", label="time_embed", flops=2.0 * B * (ic * 4 * ic * 2 + nfilm * ic), traffic=nfilm * ic * 4,
               time=self.io["t"], w1=A["noise_level_mlp.1.weight"], b1=A["noise_level_mlp.1.bias"], w2=A["noise_level_mlp.3.weight"],
               b2=A["noise_level_mlp.3.bias"], wf=A["film.w"], bf=A["film.b"], film=self.film_buf, batch=B, inner=ic, nfilm=nfilm)
        cin = net.in_channel + (net.out_channel if net.self_condition else 0)
        cin_pad = _ceil(cin, 16)
        xin = self._act(pb, "xin", B, H, W, cin_pad)
        self.in_convert_op = pb.add("ddif_in_convert_t", label="in_convert", traffic=B * H * W * (cin * 4 + cin_pad * 2), x=self.io["x"],
                                    self_cond=self.io["x"] if net.self_condition else None, out=xin.buf, batch=B, c=net.in_channel,
                                    h=H, w=W, c_pad=cin_pad)
        self.xin = xin
        x = self._act(pb, "downs.0", B, H, W, ic, stats=True)
        self._gemm(pb, "downs.0", [xin], [A["downs.0.w"]], ic, x, taps=[9], bias=A["downs.0.b"])
        self.taps = {"downs.0": (x, len(pb.ops))}
        feats = [x]
        for i, m in enumerate(net.downs):
            kind = getattr(m, "kind", None)
            p = f"downs.{i}"
            if kind == "down":
                y = self._act(pb, p, B, x.H // 2, x.W // 2, x.C, stats=True)
                self._gemm(pb, p + ".conv", [x], [A[p + ".w"]], x.C, y, taps=[9], stride=2, bias=A[p + ".b"])
                x = y
            elif kind == "enc":
                xc = self._act(pb, p + ".csm", B, x.H, x.W, m.dim_out, stats=True)
                self._gemm(pb, p + ".cond_inj.x_conv", [x], [A[p + ".cond_inj.x_conv.w"]], m.dim_out, xc, taps=[1],
                           bias=A[p + ".cond_inj.x_conv.b"], mod=("cache", self.mod[p]))
                x = self._res_block(xc, p + ".res_block")
                if m.with_attn:
                    x = self._attention(x, p + ".attn")
            else:
                continue
            self.taps[p] = (x, len(pb.ops))
            feats.append(x)
        for i, m in enumerate(net.mid):
            p = f"mid.{i}"
            x = self._res_block(x, p + ".res_block")
            if m.with_attn:
                x = self._attention(x, p + ".attn")
            self.taps[p] = (x, len(pb.ops))
        for i, m in enumerate(net.ups):
            kind = getattr(m, "kind", None)
            p = f"ups.{i}"
            if kind == "up":
                y = self._act(pb, p, B, x.H * 2, x.W * 2, x.C, stats=True)
                # nearest x2 as its own (one-load-four-stores) kernel + the halo conv: faster at every level than a conv whose loader
                # folds the up-sampling (round 1: 45 -> 41, 59 -> 45, 176 -> 127 us at B = 256)
                up = self._act(pb, p + ".up", B, x.H * 2, x.W * 2, x.C)
                pb.add("ddif_upsample2x_t", label=p + ".nearest", traffic=B * x.H * x.W * x.C * 10, **{"in": x.buf}, out=up.buf, batch=B,
                       h=x.H, w=x.W, c=x.C)
                self._gemm(pb, p + ".conv", [up], [A[p + ".w"]], x.C, y, taps=[9], bias=A[p + ".b"])
                x = y
                self.taps[p] = (x, len(pb.ops))
                continue
            skip = feats.pop()
            q = p + ".cond_inj"
            dim, o = m.dim, m.dim_out
            assert dim == x.C + skip.C, (p, dim, x.C, skip.C)
            has_res = (q + ".attn_res.w") in A
            qconv = x.H >= 16 and x.W >= 8 and dim <= 192 and self.use_qconv
            if qconv and has_res and dim + o <= 192:
                # prenorm_x -> [DW3x3 -> Conv1x1 | attn_res] as ONE tensor-core 3x3 conv over the virtual concat (x, skip) with
                # the GroupNorm fused into its loader: channels [0, dim) = q, [dim, dim + o) = r = attn_res(x_hat)
                qr = self._act(pb, q + ".qr", B, x.H, x.W, dim + o)
                # depthwise 3x3 inside the kernel + ONE tensor-core tap (q from dw(x_hat), r from x_hat) when the composed dense 3x3 would
                # waste 9x the 1x1 FLOPs on K = 9*dim >= 864; at dim = 64 the composed conv is the cheaper one (measured per layer at
                # B = 256: dim 64 @64^2 104 vs 143 us; dim 96 @64^2 238 vs 219; dim 128 @32^2 147 vs 77; dim 128 @16^2 49 vs 31)
                if self.use_dwq and dim >= 96:
                    self._gemm(pb, q + ".qconv", [x, skip], [A[q + ".q1r.w"], A[q + ".q1r.w"] + 2 * x.C], dim + o, qr, taps=[9, 9], w_s=[1, 1],
                               bias=A[q + ".qcr.b"], gn=(A[q + ".gamma"], A[q + ".beta"], 0), w_k=[dim, dim], dw=(A[q + ".q0"], dim),
                               ref_flops=2.0 * B * x.H * x.W * dim * (9 + dim + o))
                else:             # composed dense 3x3 (9x the 1x1 FLOPs on the tensor cores)
                    self._gemm(pb, q + ".qconv", [x, skip], [A[q + ".qcr.w"], A[q + ".qcr.w"] + 2 * x.C], dim + o, qr, taps=[9, 9], bias=A[q + ".qcr.b"],
                               gn=(A[q + ".gamma"], A[q + ".beta"], 0), w_k=[dim, dim],
                               ref_flops=2.0 * B * x.H * x.W * dim * (9 + dim + o))  # reference: depthwise 3x3 + 1x1 (q) + 1x1 (attn_res)
                y = self._act(pb, q + ".y", B, x.H, x.W, o)
                if self.use_cs_gemm and x.H in (16, 32, 64) and x.W % (128 // x.H) == 0 and dim % 32 == 0 and _ceil(o, 16) <= 128:
                    # q.softmax(dim=-2) inside the attn_out GEMM's loader (cs_gemm_tc_kernel: a tile = 128/H columns x all H lines); the
                    # normalised q tensor never goes to HBM
                    self._gemm(pb, q + ".softmax_h+attn_out", [qr], [("cache", self.weff[p])], o, y, taps=[1], bias=A[q + ".attn.b"],
                               per_sample=(1,), w_s=[B], w_k=[self.weff_k[p]], a_c=[self.weff_k[p]], residual=qr, residual_off=dim,
                               softmax_h=True, ref_flops=2.0 * B * x.H * x.W * o * dim)
                else:
                    qs = self._act(pb, q + ".qs", B, x.H, x.W, dim)
                    pb.add("ddif_softmax_h_t", label=q + ".softmax_h", traffic=B * x.H * x.W * dim * 4, **{"in": qr.buf}, out=qs.buf, batch=B,
                           h=x.H, w=x.W, c=dim, scale=1.0, in_ld=dim + o)
                    self._gemm(pb, q + ".attn_out", [qs], [("cache", self.weff[p])], o, y, taps=[1], bias=A[q + ".attn.b"], per_sample=(1,),
                               w_s=[B], w_k=[self.weff_k[p]], residual=qr, residual_off=dim, ref_flops=2.0 * B * x.H * x.W * o * dim)
            elif (self.use_fwm_front and not qconv and has_res and (x.H, x.W) == (8, 8) and dim in (192, 256) and o == 128
                  and x.stats is not None and skip.stats is not None):
                # the whole front (prenorm_x, DW3x3, q 1x1, softmax over H, attn_out with the per-sample W_eff, attn_res, bias) in ONE launch, one
                # CTA per sample (csrc/fwm_front.cu) instead of gn_apply+dw / 1x1 GEMM / softmax_h / two-segment 1x1 GEMM
                y = self._act(pb, q + ".y", B, x.H, x.W, o)
                pb.add("ddif_fwm_front_t", label=q + ".front", flops=2.0 * B * 64 * dim * (9 + dim + 2 * o), traffic=B * 64 * (dim + o) * 2,
                       ref_flops=2.0 * B * 64 * dim * (9 + dim + 2 * o), x=x.buf, skip=skip.buf, c1=x.C, c2=skip.C, stats1=x.stats,
                       stats2=skip.stats, gamma=A[q + ".gamma"], beta=A[q + ".beta"], eps=1e-5, dw_w=A[q + ".q0"], w1=A[q + ".q1.w"], w1_ld=dim,
                       b1=A[q + ".q1.b"], weff=("cache", self.weff[p]), weff_ld=self.weff_k[p], weff_rows=_ceil(o, 16),
                       wres=A[q + ".attn_res.w"], wres_ld=dim, bias=A[q + ".attn.b"], out=y.buf, out_ld=o, batch=B, h=x.H, w=x.W, o=o)
            else:
                qt = self._act(pb, q + ".q", B, x.H, x.W, dim)
                if qconv:
                    xh, _ = self._gn(pb, q + ".prenorm", x, A[q + ".gamma"], A[q + ".beta"], 0, src2=skip, name=q + ".xh")
                    self._gemm(pb, q + ".qconv", [x, skip], [A[q + ".qc.w"], A[q + ".qc.w"] + 2 * x.C], dim, qt, taps=[9, 9], bias=A[q + ".q1.b"],
                               gn=(A[q + ".gamma"], A[q + ".beta"], 0), w_k=[dim, dim],
                               ref_flops=2.0 * B * x.H * x.W * dim * (9 + dim))  # reference: depthwise 3x3 + 1x1
                else:
                    xh, xdw = self._gn(pb, q + ".prenorm+dw", x, A[q + ".gamma"], A[q + ".beta"], 0, src2=skip, dw_w=A[q + ".q0"], name=q + ".xh")
                    self._gemm(pb, q + ".q1", [xdw], [A[q + ".q1.w"]], dim, qt, taps=[1], bias=A[q + ".q1.b"])
                qs = self._act(pb, q + ".qs", B, x.H, x.W, dim)
                pb.add("ddif_softmax_h_t", label=q + ".softmax_h", traffic=B * x.H * x.W * dim * 4, **{"in": qt.buf}, out=qs.buf, batch=B,
                       h=x.H, w=x.W, c=dim, scale=1.0, in_ld=0)
                y = self._act(pb, q + ".y", B, x.H, x.W, o)
                weff = ("cache", self.weff[p])
                if has_res:
                    self._gemm(pb, q + ".attn_out+res", [qs, xh], [weff, A[q + ".attn_res.w"]], o, y, taps=[1, 1], bias=A[q + ".attn.b"],
                               per_sample=(1, 0), w_s=[B, 1], w_k=[self.weff_k[p], xh.C])
                else:
                    self._gemm(pb, q + ".attn_out", [qs], [weff], o, y, taps=[1], bias=A[q + ".attn.b"], per_sample=(1,), w_s=[B],
                               w_k=[self.weff_k[p]], residual=xh)
            f1 = self._act(pb, q + ".f1", B, x.H, x.W, 2 * o)
            self._gemm(pb, q + ".ffn0", [y], [A[q + ".ffn0.w"]], 2 * o, f1, taps=[9], act=1)
            z = self._act(pb, q + ".z", B, x.H, x.W, o, stats=True)
            self._gemm(pb, q + ".ffn23", [f1], [A[q + ".ffn23.w"]], o, z, taps=[9], bias=A[q + ".ffn23.b"], residual=y,
                       ref_flops=2.0 * B * x.H * x.W * o * (2 * o * 9 + o))  # reference: ffn.2 (3x3, 2o -> o) + ffn.3 (1x1)
            x = self._res_block(z, p + ".res_block")
            if m.with_attn:
                x = self._attention(x, p + ".attn")
            self.taps[p] = (x, len(pb.ops))
        self._gn_conv3("final", x, A["final.gamma"], A["final.beta"], A["final.w"], net.out_channel, None, bias=A["final.b"],
                       out_nchw=self.io["out"])
        self.stats_buf.nbytes = max(self.n_stats * B * 16, 16)
        pb.ops[0].fields["bytes"] = self.stats_buf.nbytes


def _resolve_cache(pb: PlanBuilder, cache_base: int) -> None:
    """Replace ('cache', Buf) references (cond-cache arena) by absolute addresses."""
    def fix(v):
        if isinstance(v, tuple) and len(v) == 2 and v[0] == "cache":
            return cache_base + v[1].offset
        if isinstance(v, list):
            return [fix(e) for e in v]
        return v
    for op in pb.ops:
        for k in list(op.fields):
            op.fields[k] = fix(op.fields[k])


# ----------------------------------------------------------------------------------------------------------------
class _Runtime:
    """Device buffers + recorded plans of one UNetSR3 at a fixed (B, H, W)."""

    def __init__(self, net: UNetSR3, B: int, H: int, W: int, use_side_branch: bool = True, **schedule_options):
        _lib.load()
        P = net._packed
        dev = next(net.parameters()).device
        self.net, self.B, self.H, self.W, self.dev = net, B, H, W, dev
        C = net.in_channel
        ctot = 2 * net.lms_channel + 4 * net.pan_channel
        f32 = dict(dtype=torch.float32, device=dev)
        self.x_buf = torch.zeros(B, C, H, W, **f32)
        self.sc_buf = torch.zeros(B, net.out_channel, H, W, **f32)
        self.t_buf = torch.zeros(B, **f32)
        self.out_buf = torch.zeros(B, net.out_channel, H, W, **f32)
        self.cond_buf = torch.zeros(B, ctot, H, W, **f32)
        addr = {k: v.data_ptr() for k, v in P.items() if isinstance(v, torch.Tensor)}
        io = dict(x=self.x_buf.data_ptr(), sc=self.sc_buf.data_ptr(), t=self.t_buf.data_ptr(), out=self.out_buf.data_ptr(),
                  cond=self.cond_buf.data_ptr())
        sch = Schedule(net, addr, B, H, W, io, **schedule_options)
        sch.build_cond()
        sch.build_forward(P["_film_offsets"], int(P["film.w"].shape[0]))
        self.sch = sch
        cache_bytes = sch.cache.layout()
        self.cache = torch.zeros(max(cache_bytes, 16), dtype=torch.uint8, device=dev)
        ws_bytes = max(sch.cnd.layout(), sch.fwd.layout())
        self.ws = torch.zeros(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        for pb in (sch.cnd, sch.fwd):
            _resolve_cache(pb, self.cache.data_ptr())
            pb.finalize(self.ws.data_ptr(), device=dev)
        # The time embedding + FiLM vectors (one latency-bound launch, ~30 us) run as a side branch of the step graph, concurrently with the
        # input conversion and the first convs; the first op that adds a FiLM vector joins it.  Its output buffer is live from op 1 to the last
        # FiLM consumer, so the liveness packer never aliases it with anything those ops touch.
        te = next(i for i, op in enumerate(sch.fwd.ops) if op.struct == "ddif_time_embed_t")
        users = [i for i, op in enumerate(sch.fwd.ops) if i > te and op.fields.get("film") is not None]
        if use_side_branch and users and users[0] > te + 1:
            sch.fwd.set_side_branch(te, te + 1, users[0])
        self.cond_key = None
        self.graph_ready = False
        self.use_graph = True

    @property
    def stream(self) -> int:
        return _lib.stream_of(self.dev)

    def set_cond(self, cond: torch.Tensor, force: bool = False) -> None:
        """Rebuild the cond cache unless `cond` is the very tensor object cached last time and has not been written
        since (object identity + version counter; the object is kept alive so its address cannot be recycled)."""
        if not force and self.cond_key is not None and self.cond_key[0] is cond and self.cond_key[1] == cond._version:
            return
        if tuple(cond.shape) != tuple(self.cond_buf.shape):
            raise ValueError(f"cond must be {tuple(self.cond_buf.shape)}, got {tuple(cond.shape)}")
        self.cond_buf.copy_(cond)
        self.sch.cnd.run(self.stream)
        self.cond_key = (cond, cond._version)

    def step(self, explicit_self_cond: bool = False) -> None:
        """One UNet forward on the fixed buffers (x_buf, t_buf[, sc_buf]) -> out_buf."""
        fwd = self.sch.fwd
        if explicit_self_cond:
            k = self.sch.in_convert_op
            fwd.run(self.stream, 0, k)
            _lib.launch("ddif_in_convert_t", self.stream, x=self.x_buf.data_ptr(), self_cond=self.sc_buf.data_ptr(),
                        out=self.ws.data_ptr() + self.sch.xin.buf.offset, batch=self.B, c=self.net.in_channel, h=self.H, w=self.W,
                        c_pad=self.sch.xin.C)
            fwd.run(self.stream, k + 1, -1)
            return
        if self.use_graph:
            if not self.graph_ready:
                side = torch.cuda.Stream(self.dev)
                side.wait_stream(torch.cuda.current_stream(self.dev))
                with torch.cuda.stream(side):
                    fwd.graph_build(_lib.Stream(side.cuda_stream, self.dev))
                torch.cuda.current_stream(self.dev).wait_stream(side)
                self.graph_ready = True
            fwd.graph_launch(self.stream)
        else:
            fwd.run(self.stream)

    def debug_taps(self, x, time, cond) -> Dict[str, torch.Tensor]:
        """Run the forward op by op and return every block output (fp32 NCHW) keyed by its state-dict prefix."""
        self.set_cond(cond)
        self.x_buf.copy_(x)
        self.t_buf.copy_(time.reshape(-1).to(torch.float32).expand(self.B))
        fwd, cur, out = self.sch.fwd, 0, {}
        for name, (act, end) in self.sch.taps.items():
            fwd.run(self.stream, cur, end)
            cur = end
            torch.cuda.synchronize(self.dev)
            n = act.B * act.H * act.W * act.C * 2
            t = self.ws[act.buf.offset: act.buf.offset + n].view(torch.bfloat16).view(act.B, act.H, act.W, act.C)
            out[name] = t.float().permute(0, 3, 1, 2).contiguous()
        fwd.run(self.stream, cur, -1)
        return out

    def launches_per_step(self) -> int:
        return len(self.sch.fwd)

    def close(self) -> None:
        self.sch.fwd.close()
        self.sch.cnd.close()
