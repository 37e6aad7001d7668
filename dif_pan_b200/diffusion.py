"""GaussianDiffusion — drop-in for /root/reference/diffusion/diffusion_ddpm_pan.py:143-778 (sampling + q_sample).

Same constructor, same buffers (14 fp32 `[T]` tensors computed in float64 numpy exactly like the reference),
same `forward(x, mode="ddpm_sample" | "ddim_sample" | "train")` dispatch.  The per-step arithmetic runs as ONE
fused CUDA kernel per step (csrc/sampler.cu) instead of ~25 elementwise ATen kernels; when the denoiser is a
`dif_pan_b200.UNetSR3` the loop works in place on the model's fixed device buffers and every step is one
CUDA-graph launch + one fused sampler kernel (no per-step host<->device traffic).

Noise: the reference draws `torch.randn` on the device (diffusion_ddpm_pan.py:86,484).  Here the loops accept
`noise=[x_T, n_1, ...]` (injected tensors, for parity tests); otherwise noise comes from an in-kernel
Philox4x32-10 generator seeded by `seed`.
"""
from __future__ import annotations

import math
import random
from functools import partial
from typing import Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import _lib
from .unet import UNetSR3

PRED_MODES = {"x_start": 0, "noise": 1, "pred_v": 2}


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """diffusion_ddpm_pan.py:26-57 (float64)."""
    if schedule == "quad":
        return np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2
    if schedule == "linear":
        return np.linspace(linear_start, linear_end, n_timestep, dtype=np.float64)
    if schedule in ("warmup10", "warmup50"):
        frac = 0.1 if schedule == "warmup10" else 0.5
        betas = linear_end * np.ones(n_timestep, dtype=np.float64)
        k = int(n_timestep * frac)
        betas[:k] = np.linspace(linear_start, linear_end, k, dtype=np.float64)
        return betas
    if schedule == "const":
        return linear_end * np.ones(n_timestep, dtype=np.float64)
    if schedule == "jsd":
        return 1.0 / np.linspace(n_timestep, 1, n_timestep, dtype=np.float64)
    if schedule == "cosine":
        ts = torch.arange(n_timestep + 1, dtype=torch.float64) / n_timestep + cosine_s
        al = torch.cos(ts / (1 + cosine_s) * math.pi / 2).pow(2)
        al = al / al[0]
        return (1 - al[1:] / al[:-1]).clamp(max=0.999)
    raise NotImplementedError(schedule)


def _stream(dev) -> int:
    return _lib.stream_of(dev)


class GaussianDiffusion(nn.Module):
    def __init__(self, denoise_fn, image_size, channels=3, loss_type="l2", conditional=True, schedule_opt=None,
                 device="cuda:0", clamp_range=(-1.0, 1.0), clamp_type="abs", pred_mode="noise",
                 p2_loss_weight_gamma=0.0, p2_loss_weight_k=1):
        super().__init__()
        assert clamp_type in ["abs", "dynamic"]
        assert pred_mode in ["noise", "x_start", "pred_v"]
        assert loss_type in ["l1", "l2", "l1ssim"]
        self.channels, self.image_size, self.model = channels, image_size, denoise_fn
        self.conditional, self.loss_type, self.device = conditional, loss_type, device
        self.clamp_range, self.clamp_type = clamp_range, clamp_type
        self.p2_loss_weight_gamma, self.p2_loss_weight_k = p2_loss_weight_gamma, p2_loss_weight_k
        if schedule_opt is not None:
            self.set_new_noise_schedule(schedule_opt, device)
        self.pred_mode = pred_mode
        self.self_condition = self.model.self_condition
        self.pred_var = self.model.pred_var
        assert self.pred_var == False, "not supported yet"  # noqa: E712  (diffusion_ddpm_pan.py:184)
        self.seed = 0
        # In-kernel noise is a function of (seed, step, GLOBAL element index): a rank that samples patches [lo, lo + b) of a batch of
        # `total` sets noise_shard = (lo, total) (sharding.sample_sharded does) and draws exactly the numbers the single-process run
        # draws for those patches, so Philox results do not depend on the number of ranks.  None = (0, local batch).
        self.noise_shard = None
        self._coef = {}

    # -- schedule (diffusion_ddpm_pan.py:199-276) --------------------------------------------------------------
    def set_new_noise_schedule(self, schedule_opt=None, device="cpu", *, betas=None):
        to_torch = partial(torch.tensor, dtype=torch.float32, device=device)
        if schedule_opt is not None:
            betas = make_beta_schedule(schedule=schedule_opt["schedule"], n_timestep=schedule_opt["n_timestep"],
                                       linear_start=schedule_opt["linear_start"], linear_end=schedule_opt["linear_end"])
        betas = betas.detach().cpu().numpy() if isinstance(betas, torch.Tensor) else np.asarray(betas)
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        ac_next = np.append(ac[1:], 0.0)
        self.num_timesteps = int(betas.shape[0])
        pv = betas * (1.0 - ac_prev) / (1.0 - ac)
        bufs = dict(
            betas=betas, alphas_cumprod=ac, alphas_cumprod_prev=ac_prev, alphas_cumprod_next=ac_next,
            sqrt_alphas_cumprod=np.sqrt(ac), sqrt_one_minus_alphas_cumprod=np.sqrt(1.0 - ac),
            log_one_minus_alphas_cumprod=np.log(1.0 - ac), sqrt_recip_alphas_cumprod=np.sqrt(1.0 / ac),
            sqrt_recipm1_alphas_cumprod=np.sqrt(1.0 / ac - 1), posterior_variance=pv,
            posterior_log_variance_clipped=np.log(np.maximum(pv, 1e-20)),
            posterior_mean_coef1=betas * np.sqrt(ac_prev) / (1.0 - ac),
            posterior_mean_coef2=(1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
            p2_loss_weight=(self.p2_loss_weight_k + ac / (1 - ac)) ** -self.p2_loss_weight_gamma)
        for k, v in bufs.items():
            self.register_buffer(k, to_torch(v))
        self._coef = {}

    def _coef_table(self, kind: str) -> torch.Tensor:
        """fp32 [T][8] device table consumed by the fused step kernels (include/ddif_b200.h)."""
        t = self._coef.get(kind)
        if t is not None and t.device == self.betas.device:
            return t
        z = torch.zeros_like(self.betas)
        if kind == "ddpm":
            cols = [self.posterior_mean_coef1, self.posterior_mean_coef2, self.posterior_log_variance_clipped,
                    self.sqrt_recip_alphas_cumprod, self.sqrt_recipm1_alphas_cumprod, self.sqrt_alphas_cumprod,
                    self.sqrt_one_minus_alphas_cumprod, z]
        else:
            cols = [self.alphas_cumprod, self.alphas_cumprod_prev, self.sqrt_recip_alphas_cumprod,
                    self.sqrt_recipm1_alphas_cumprod, self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod, z, z]
        t = torch.stack(cols, dim=1).contiguous()
        self._coef[kind] = t
        return t

    # -- fused steps (public, usable with any denoiser) ----------------------------------------------------------
    def _step(self, kind, x, model_out, cond, t: int, noise=None, time_out=None, eta=0.0, clip=True, offset=0):
        """In-place x <- step(x, model_out) for the whole batch at integer timestep t."""
        if not x.is_cuda:
            raise RuntimeError("dif_pan_b200 sampler kernels run on CUDA only (no CPU fallback)")
        B, C = x.shape[0], x.shape[1]
        hw = x.shape[2] * x.shape[3]
        lo, hi = (self.clamp_range if self.clamp_range is not None else (0.0, 0.0))
        if clip and self.clamp_type != "abs":
            raise NotImplementedError("clamp_type='dynamic' (torch.quantile thresholding) is not on the CUDA path")
        common = dict(x=x.data_ptr(), model_out=model_out.data_ptr(), cond=cond.data_ptr() if cond is not None else None,
                      noise=noise.data_ptr() if noise is not None else None, coef=self._coef_table(kind).data_ptr(),
                      time_out=time_out.data_ptr() if time_out is not None else None, batch=B, c=C, hw=hw,
                      cond_c=cond.shape[1] if cond is not None else 0, t=int(t), pred_mode=PRED_MODES[self.pred_mode],
                      clip=int(bool(clip)), clamp_lo=float(lo), clamp_hi=float(hi), seed=int(self.seed), offset=int(offset))
        if kind == "ddpm":
            _lib.launch("ddif_ddpm_step_t", _stream(x.device), **common)
        else:
            _lib.launch("ddif_ddim_step_t", _stream(x.device), eta=float(eta), **common)
        return x

    def q_sample(self, x_start, t, noise=None):
        """diffusion_ddpm_pan.py:668-681."""
        if noise is None:
            noise = device_randn(x_start.shape, x_start.device, self.seed, 1 << 40)
        # contiguous copies are bound to locals: a temporary freed before the launch could hand its block to the next one
        x0c, nzc, tc = x_start.contiguous(), noise.contiguous(), t.to(torch.int64).contiguous()
        out = torch.empty_like(x0c)
        B = x_start.shape[0]
        _lib.launch("ddif_q_sample_t", _stream(x_start.device), x0=x0c.data_ptr(), noise=nzc.data_ptr(),
                    out=out.data_ptr(), sa=self.sqrt_alphas_cumprod.data_ptr(), s1ma=self.sqrt_one_minus_alphas_cumprod.data_ptr(),
                    t=tc.data_ptr(), batch=B, chw=x_start[0].numel())
        return out

    # -- loops -------------------------------------------------------------------------------------------------
    def _fast(self, cond):
        return isinstance(self.model, UNetSR3) and cond.is_cuda

    def _run_loop(self, kind, cond, n_steps, noise: Optional[Sequence[torch.Tensor]], eta=0.0, clip=True, continous=False):
        dev = self.betas.device
        b, (h, w) = cond.shape[0], cond.shape[-2:]
        shape = (b, self.channels, h, w)
        sample_inter = 1 | (self.num_timesteps // 10)
        lo, total = self.noise_shard if self.noise_shard is not None else (0, b)
        per = self.channels * h * w
        if lo < 0 or lo + b > total:
            raise ValueError(f"noise_shard {self.noise_shard} does not contain a local batch of {b}")
        # Philox counter of local float4 i at step k: k * (total elements) + (lo * per) / 4 + i  (csrc/sampler.cu: counter = offset + i)
        off = lambda k: k * total * per + (lo * per) // 4
        if self._fast(cond):
            rt = self.model.runtime(b, h, w)
            rt.set_cond(cond)
            x, out, tbuf = rt.x_buf, rt.out_buf, rt.t_buf
            if noise is not None:
                x.copy_(noise[0])
            else:
                device_randn_(x, self.seed, off(0))
            tbuf.fill_(float(n_steps - 1))
            ret = [x.clone()] if continous else None
            for k, i in enumerate(reversed(range(n_steps))):
                rt.step()
                self._step(kind, x, out, rt.cond_buf, i, noise[1 + k] if noise is not None else None, time_out=tbuf, eta=eta,
                           clip=clip, offset=off(k + 1))
                if continous and i % sample_inter == 0:
                    ret.append(x.clone())
            return torch.cat(ret, 0) if continous else x.clone()
        # generic denoiser (any nn.Module with the reference call signature)
        img = noise[0].clone() if noise is not None else device_randn(shape, dev, self.seed, off(0))
        ret = [img.clone()] if continous else None
        x_start = None
        for k, i in enumerate(reversed(range(n_steps))):
            t = torch.full((b,), i, device=dev, dtype=torch.long)
            sc = x_start if (self.self_condition and kind == "ddpm") else None
            out = self.model(img, t, cond, sc).contiguous()
            self._step(kind, img, out, cond.contiguous(), i, noise[1 + k] if noise is not None else None, eta=eta, clip=clip,
                       offset=off(k + 1))
            x_start = img
            if continous and i % sample_inter == 0:
                ret.append(img.clone())
        return torch.cat(ret, 0) if continous else img

    @torch.no_grad()
    def p_sample_loop(self, x_in, continous=False, get_interm_fm=False, noise=None):
        """diffusion_ddpm_pan.py:444-507 (conditional branch): x_in is `cond`; self_cond == current image."""
        if not self.conditional:
            raise NotImplementedError("unconditional sampling is not used by the DDIF engine")
        clip = self.clamp_range is not None
        return self._run_loop("ddpm", x_in, self.num_timesteps, noise, clip=clip, continous=continous)

    @staticmethod
    def space_timesteps(num_timesteps, section_counts):
        """diffusion_ddpm_pan.py:529-581."""
        if isinstance(section_counts, str):
            if section_counts.startswith("ddim"):
                want = int(section_counts[len("ddim"):])
                for i in range(1, num_timesteps):
                    if len(range(0, num_timesteps, i)) == want:
                        return set(range(0, num_timesteps, i))
                raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
            section_counts = [int(x) for x in section_counts.split(",")]
        size_per, extra = divmod(num_timesteps, len(section_counts))
        start, steps = 0, []
        for i, cnt in enumerate(section_counts):
            size = size_per + (1 if i < extra else 0)
            if size < cnt:
                raise ValueError(f"cannot divide section of {size} steps into {cnt}")
            stride = 1 if cnt <= 1 else (size - 1) / (cnt - 1)
            cur = 0.0
            for _ in range(cnt):
                steps.append(start + round(cur))
                cur += stride
            start += size
        return set(steps)

    def space_new_betas(self, use_timesteps):
        """diffusion_ddpm_pan.py:583-592 — overwrites this module's schedule permanently, like the reference."""
        last, new_betas = 1.0, []
        for i, ac in enumerate(self.alphas_cumprod.cpu()):
            if i in use_timesteps:
                new_betas.append((1 - ac / last).item())
                last = ac
        self.set_new_noise_schedule(betas=np.array(new_betas), device=self.betas.device)

    @torch.no_grad()
    def ddim_sample_loop(self, x_in, section_counts="ddim300", eta=0.0, noise=None):
        """diffusion_ddpm_pan.py:623-666: respace, then len(betas) steps with self_cond=None, clip_denoised=False."""
        if not self.conditional:
            raise NotImplementedError("unconditional sampling is not used by the DDIF engine")
        assert isinstance(x_in, torch.Tensor)
        self.space_new_betas(self.space_timesteps(self.num_timesteps, section_counts))
        return self._run_loop("ddim", x_in, len(self.betas), noise, eta=eta, clip=False)

    def _axpby(self, ca: torch.Tensor, x: torch.Tensor, cb: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """out[b] = ca[b]*x[b] + cb[b]*y[b] with per-sample fp32 coefficients gathered from the schedule buffers
        (`extract`, diffusion_ddpm_pan.py:73-76): predict_start_from_noise / predict_v / predict_start_from_v (:284-312)."""
        xc, yc = x.contiguous(), y.contiguous()
        cac, cbc = ca.to(torch.float32).contiguous(), cb.to(torch.float32).contiguous()
        out = torch.empty_like(xc)
        _lib.launch("ddif_axpby_t", _stream(x.device), x=xc.data_ptr(), y=yc.data_ptr(), ca=cac.data_ptr(), cb=cbc.data_ptr(), out=out.data_ptr(),
                    batch=x.shape[0], chw=x[0].numel())
        return out

    def p_losses(self, x_start, noise=None, cond=None, *, t=None, self_cond_draw=None):
        """Training objective (diffusion_ddpm_pan.py:692-766): t ~ U{0..T-1}, x_t = q_sample, the optional no-grad self-conditioning
        pass (probability 0.5, Python `random.random()` like the reference, :702), the denoiser, then
        mean(loss_func(target, prediction)) * mean(p2_loss_weight[t]) and recon_x0.  Returns (loss, recon_x0) like the reference.
        With autograd enabled and a denoiser whose parameters require grad, `loss` carries the graph (`loss.backward()` reaches the
        UNet's parameters through training.py); otherwise the objective is evaluated forward-only with the fused loss kernel.
        `t` / `self_cond_draw` pin the two random draws for tests."""
        if self.loss_type not in ("l1", "l2"):
            raise NotImplementedError("loss_type='l1ssim' is not on the CUDA path")
        if not x_start.is_cuda:
            raise RuntimeError("dif_pan_b200 kernels run on CUDA only (no CPU fallback)")
        want_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.model.parameters())
        b = x_start.shape[0]
        dev = x_start.device
        x_start = x_start.detach().to(torch.float32).contiguous()
        if t is None:
            t = torch.randint(0, self.num_timesteps, (b,), device=dev).long()
        if noise is None:
            noise = device_randn(x_start.shape, dev, self.seed, 1 << 41)
        noise = noise.detach().to(torch.float32).contiguous()
        ex = lambda buf: buf.gather(-1, t)
        with torch.no_grad():
            x_noisy = self.q_sample(x_start, t, noise)
            x_self_cond = None
            draw = random.random() if self_cond_draw is None else self_cond_draw
            if self.self_condition and draw < 0.5:
                mo = self.model(x_noisy, t, cond=cond, self_cond=None) if self.conditional else self.model(x_noisy, t, self_cond=None)
                mo = mo.to(torch.float32).contiguous()
                if self.pred_mode == "noise":
                    x_self_cond = self._axpby(ex(self.sqrt_recip_alphas_cumprod), x_noisy, -ex(self.sqrt_recipm1_alphas_cumprod), mo)
                elif self.pred_mode == "x_start":
                    x_self_cond = mo
                else:
                    x_self_cond = self._axpby(ex(self.sqrt_alphas_cumprod), x_noisy, -ex(self.sqrt_one_minus_alphas_cumprod), mo)
        with torch.set_grad_enabled(want_grad):
            if self.conditional:
                pred = self.model(x_noisy, t, cond=cond, self_cond=x_self_cond)
            else:
                pred = self.model(x_noisy, t, self_cond=x_self_cond)
        with torch.no_grad():
            pd = pred.detach().to(torch.float32).contiguous()
            if self.pred_mode == "noise":
                recon = self._axpby(ex(self.sqrt_recip_alphas_cumprod), x_noisy, -ex(self.sqrt_recipm1_alphas_cumprod), pd)
                target = noise
            elif self.pred_mode == "x_start":
                recon, target = pd, x_start
            else:
                target = self._axpby(ex(self.sqrt_alphas_cumprod), noise, -ex(self.sqrt_one_minus_alphas_cumprod), x_start)
                recon = self._axpby(ex(self.sqrt_alphas_cumprod), x_noisy, -ex(self.sqrt_one_minus_alphas_cumprod), target)
        # nn.L1Loss() / nn.MSELoss() reduce to a scalar first (:189-193), so the reference's objective is
        # mean(l) * mean_b(p2_loss_weight[t_b]) (:759-762), not a per-sample weighting
        if want_grad:
            l = (target - pred).abs().mean() if self.loss_type == "l1" else ((target - pred) ** 2).mean()
            return (l * ex(self.p2_loss_weight)).mean(), recon
        with torch.no_grad():
            acc = torch.zeros(1, dtype=torch.float64, device=dev)
            _lib.launch("ddif_loss_t", _stream(dev), a=target.data_ptr(), b=pd.data_ptr(), weight=None, out=acc.data_ptr(),
                        batch=b, chw=x_start[0].numel(), squared=0 if self.loss_type == "l1" else 1)
            loss = ((acc / float(x_start.numel())).to(torch.float32) * ex(self.p2_loss_weight)).mean()
        return loss, recon

    def forward(self, x, mode="train", *args, **kwargs):
        if mode == "train":
            return self.p_losses(x, *args, **kwargs)
        elif mode == "ddpm_sample":
            with torch.no_grad():
                return self.p_sample_loop(x, *args, **kwargs)
        elif mode == "ddim_sample":
            with torch.no_grad():
                return self.ddim_sample_loop(x, *args, **kwargs)
        else:
            raise NotImplementedError("mode should be train or sample")


def device_randn_(out: torch.Tensor, seed: int, offset: int) -> torch.Tensor:
    """Fill `out` (fp32, numel % 4 == 0) with N(0,1) from the in-kernel Philox generator."""
    _lib.launch("ddif_randn_t", _stream(out.device), out=out.data_ptr(), n=out.numel(), seed=int(seed), offset=int(offset))
    return out


def device_randn(shape, device, seed: int, offset: int) -> torch.Tensor:
    return device_randn_(torch.empty(shape, dtype=torch.float32, device=device), seed, offset)


def fuse_output(sample: torch.Tensor, cond: torch.Tensor, lo=0.0, hi=1.0) -> torch.Tensor:
    """sr = clip(sample + lms, 0, 1) (diffusion_engine.py:446-447); lms = first C channels of cond."""
    sc, cc = sample.contiguous(), cond.contiguous()
    out = torch.empty_like(sc)
    B, C = sample.shape[:2]
    _lib.launch("ddif_axpby_clip_t", _stream(sample.device), x=sc.data_ptr(), cond=cc.data_ptr(),
                out=out.data_ptr(), batch=B, c=C, hw=sample.shape[2] * sample.shape[3], cond_c=cond.shape[1], lo=float(lo), hi=float(hi))
    return out
