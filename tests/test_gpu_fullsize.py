"""Parity at BASELINE.json's FULL sizes through size-independent properties (the CPU oracle cannot run 256 patches x 500 steps):

  * configs[1] size (WV3, batch 256): every sample of a batch is an independent chain, so the batch-256 output
    of samples {0, 1, 127, 255} must equal (a) the same samples run as a batch of 4 and (b) the CPU oracle on those samples;
    permuting the batch permutes the output; two runs are bit-stable up to the fp64 statistics-atomics order.
  * configs[3] size (CAVE, batch 128): first / last sample against the oracle.
  * sampler kernels and DWT at full size: step kernel == oracle on a strided subset, idwt(dwt(x)) == x, DDPM loop equivariance.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import dif_pan_b200 as dp  # noqa: E402
from dif_pan_b200 import synth  # noqa: E402
from oracle import sampler_oracle as so, unet_oracle as uo  # noqa: E402

DEV = "cuda:0"
torch.set_grad_enabled(False)


def _net(dataset):
    kw = synth.unet_kwargs(dataset)
    net = dp.UNetSR3(**kw)
    sd = synth.make_state_dict(0, **kw)
    net.load_state_dict(sd)
    kw2 = dict(kw); kw2.pop("dropout")
    return net.to(DEV).eval(), sd, uo.UNetCfg(**kw2)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def _big_batch(dataset, B, C):
    base = synth.make_batch(dataset, 8, seed=77)["cond"]
    gen = torch.Generator().manual_seed(78)
    # 8 distinct scenes tiled to B samples with per-sample brightness jitter, so that no two samples are identical
    cond = base.repeat((B + 7) // 8, 1, 1, 1)[:B] * (0.8 + 0.4 * torch.rand(B, 1, 1, 1, generator=gen))
    x = torch.randn(B, C, 64, 64, generator=gen)
    t = torch.randint(0, 500, (B,), generator=gen)
    return x, t, cond.contiguous()


def test_unet_batch256_sample_independence_and_oracle():
    net, sd, cfg = _net("wv3")
    B = 256
    x, t, cond = _big_batch("wv3", B, 8)
    y = net(x.to(DEV), t.to(DEV), cond.to(DEV)).cpu()
    assert torch.isfinite(y).all()
    pick = [0, 1, 127, 255]
    y4 = net(x[pick].to(DEV), t[pick].to(DEV), cond[pick].contiguous().to(DEV)).cpu()
    ref = uo.unet_forward(sd, cfg, x[pick], t[pick], cond[pick])
    for i, b in enumerate(pick):
        assert _rel(y[b], y4[i]) <= 1e-3, (b, _rel(y[b], y4[i]))      # same kernels, different tile -> CTA assignment and atomics order
        assert _rel(y[b], ref[i]) <= 1e-2, (b, _rel(y[b], ref[i]))    # north-star per-step tolerance at the full batch size
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(3))
    yp = net(x[perm].to(DEV), t[perm].to(DEV), cond[perm].contiguous().to(DEV)).cpu()
    assert _rel(yp, y[perm]) <= 1e-3
    y2 = net(x.to(DEV), t.to(DEV), cond.to(DEV)).cpu()
    assert _rel(y2, y) <= 1e-4


def test_unet_cave_batch128_against_oracle():
    net, sd, cfg = _net("cave")
    B = 128
    x, t, cond = _big_batch("cave", B, 31)
    y = net(x.to(DEV), t.to(DEV), cond.to(DEV)).cpu()
    pick = [0, 127]
    ref = uo.unet_forward(sd, cfg, x[pick], t[pick], cond[pick])
    for i, b in enumerate(pick):
        assert _rel(y[b], ref[i]) <= 1e-2, (b, _rel(y[b], ref[i]))


def test_ddpm_loop_batch256_equivariance_and_subset_oracle():
    """T = 3 DDPM loop at batch 256 with injected noise: samples {0, 255} equal the oracle's loop on those samples within the image
    tolerance of the north star (PSNR of the difference), and a batch permutation permutes the result."""
    net, sd, cfg = _net("wv3")
    B, T = 256, 3
    _, _, cond = _big_batch("wv3", B, 8)
    gen = torch.Generator().manual_seed(5)
    noises = [torch.randn(B, 8, 64, 64, generator=gen) for _ in range(T + 1)]

    def run(c, nz):
        dif = dp.GaussianDiffusion(net, image_size=64, channels=8, pred_mode="x_start", loss_type="l1", device=DEV, clamp_range=(0, 1))
        dif.set_new_noise_schedule(betas=dp.make_beta_schedule("cosine", T), device=DEV)
        return dif.to(DEV)(c.to(DEV), mode="ddpm_sample", noise=[n.to(DEV) for n in nz]).cpu()

    out = run(cond, noises)
    pick = [0, 255]
    model = lambda xx, tt, c, sc: uo.unet_forward(sd, cfg, xx, tt, c, self_cond=sc)
    ref = so.ddpm_sample_loop(model, so.schedule_buffers(so.make_beta_schedule("cosine", T)), cond[pick], 8, [n[pick] for n in noises])
    for i, b in enumerate(pick):
        mse = float(((out[b] - ref[i]) ** 2).mean())
        assert mse < 1e-3 * float((ref[i] ** 2).mean()), (b, mse)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(4))
    outp = run(cond[perm].contiguous(), [n[perm].contiguous() for n in noises])
    assert _rel(outp, out[perm]) <= 2e-3


def test_sampler_kernels_and_dwt_full_size():
    B = 256
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(B, 8, 64, 64, generator=gen)
    mo = torch.randn(B, 8, 64, 64, generator=gen) * 0.3
    nz = torch.randn(B, 8, 64, 64, generator=gen)
    cond = torch.rand(B, 20, 64, 64, generator=gen)
    sb = so.schedule_buffers(so.make_beta_schedule("cosine", 500))
    dif = dp.GaussianDiffusion(type("M", (), {"self_condition": True, "pred_var": False})(), image_size=64, channels=8, pred_mode="x_start",
                               loss_type="l1", device=DEV, clamp_range=(0, 1))
    dif.set_new_noise_schedule(betas=dp.make_beta_schedule("cosine", 500), device=DEV)
    for t in (0, 250, 499):
        xd = x.to(DEV).clone()
        dif._step("ddpm", xd, mo.to(DEV), cond.to(DEV), t, noise=nz.to(DEV))
        sub = slice(0, B, 37)
        ref = so.ddpm_step(sb, x[sub], torch.full((len(range(0, B, 37)),), t, dtype=torch.long), mo[sub], cond[sub, :8], nz[sub])
        assert float((xd.cpu()[sub] - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))
    big = torch.randn(B, 8, 64, 64, generator=gen).to(DEV)
    cA, co = dp.haar_dwt2(big)
    assert float((dp.haar_idwt2(cA, co) - big).abs().max()) <= 1e-6
    # Parseval: the orthonormal Haar transform preserves energy
    e0 = float((big.double() ** 2).sum())
    e1 = float(sum((b.double() ** 2).sum() for b in (cA, *co)))
    assert abs(e0 - e1) <= 1e-6 * e0
