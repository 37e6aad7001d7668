#!/usr/bin/env python
"""Secondary measurements on one B200 (the BASELINE configs that are parity cases rather than the bench line, plus the memory-bound
kernels against the HBM roofline).  Prints one JSON object per line; tools/collect_evidence.sh stores it under profiles/.
  config 1: WV3 64x64, batch 1, DPM-Solver++ 2M 20 steps (latency case)
  config 3: GF2 512x512 scene, tiled into 64 patches of 64x64 + DPM-Solver++ 25 steps, and whole-scene DDIM-25 (what test_fn does)
  config 4: CAVE 64x64 (31 bands + RGB), batch 128, DDIM-25
  kernels : Haar DWT, fused cond prep, DDPM / DPM++ / singlestep update, q_sample, metrics -- algorithmic bytes / CUDA-event time
"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dif_pan_b200 as dp
from dif_pan_b200 import synth, _lib, metrics as dm

DEV = "cuda:0"
torch.set_grad_enabled(False)
PEAK = 6650.0
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = json.load(open(p))["hbm_gbs"]


def timed(fn, warm=2, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def net_for(ds):
    kw = synth.unet_kwargs(ds)
    net = dp.UNetSR3(**kw)
    net.load_state_dict(synth.make_state_dict(0, **kw))
    return net.to(DEV).eval(), kw


def emit(**kw):
    print(json.dumps(kw), flush=True)


def config1():
    net, kw = net_for("wv3")
    cond = synth.make_batch("wv3", 1, seed=1)["cond"].to(DEV)
    x_T = torch.randn(1, 8, 64, 64, device=DEV)
    ms = timed(lambda: dp.sample_cond(net, cond, 8, "dpm20", x_T=x_T))
    emit(config="configs[0]: WV3 64x64, batch 1, DPM-Solver++ 2M, 20 steps", ms_per_sampling=ms, patches_per_s=1000.0 / ms, ms_per_denoise_step=ms / 20,
         note="latency case: 20 x (CUDA-graph UNet forward + 1 fused solver kernel), host loop included")


def config3():
    net, kw = net_for("gf2")
    d = synth.make_batch("gf2", 1, size=512, seed=2)
    lms, pan = d["lms_dn"].float().to(DEV), d["pan_dn"].float().to(DEV)
    div = synth.DATASETS["gf2"].division
    ms_t = timed(lambda: dp.fuse_scene(net, lms, pan, div, sampler="dpm25", patch=64, tile_batch=64), warm=1, reps=2)
    emit(config="configs[2]: GF2 512x512 scene -> 64 tiles of 64x64, DPM-Solver++ 2M 25 steps, stitched", ms_per_scene=ms_t, scenes_per_s=1000.0 / ms_t,
         patches_per_s=64000.0 / ms_t, includes="fused cond prep, tiling, sampling, clip(+lms), stitching")
    ms_w = timed(lambda: dp.fuse_scene(net, lms, pan, div, sampler="ddim25"), warm=1, reps=2)
    emit(config="reference test_fn mode: GF2 512x512 scene sampled WHOLE (B=1, H=W=512), DDIM-25", ms_per_scene=ms_w, scenes_per_s=1000.0 / ms_w,
         ms_per_denoise_step=ms_w / 25)


def config4():
    net, kw = net_for("cave")
    cond = synth.make_batch("cave", 8, seed=3)["cond"].repeat(16, 1, 1, 1).contiguous().to(DEV)
    ms = timed(lambda: dp.sample_cond(net, cond, 31, "ddim25"), warm=1, reps=2)
    emit(config="configs[3]: CAVE 64x64 (31-band HSI + RGB cond, 74 channels), batch 128, DDIM-25", ms_per_sampling=ms, patches_per_s=128000.0 / ms,
         ms_per_denoise_step=ms / 25)


def kernels():
    B, C, H, W = 256, 8, 64, 64
    n = B * C * H * W
    x = torch.randn(B, C, H, W, device=DEV)
    rows = []

    def row(name, bytes_, fn):
        ms = timed(fn, warm=3, reps=5)
        gbs = bytes_ / ms / 1e6
        rows.append(dict(kernel=name, algorithmic_bytes=bytes_, us=ms * 1e3, achieved_gbs=gbs, frac_of_hbm_peak=gbs / PEAK))

    big = torch.randn(64, 8, 512, 512, device=DEV)  # 537 MB: larger than L2
    row("haar_dwt2 (64x8x512x512)", big.numel() * 8, lambda: dp.haar_dwt2(big))
    lms, pan = torch.rand(64, 8, 256, 256, device=DEV) * 2047, torch.rand(64, 1, 256, 256, device=DEV) * 2047
    row("make_cond fused (64 x WV3 256x256)", (lms.numel() + pan.numel()) * 4 + 64 * 20 * 256 * 256 * 4, lambda: dp.make_cond(lms, pan, 2047.0))
    xb = torch.randn(2048, 8, 64, 64, device=DEV)  # 268 MB per tensor
    ob, cb = torch.randn_like(xb), torch.rand(2048, 20, 64, 64, device=DEV)
    nz = torch.randn_like(xb)
    dif = dp.GaussianDiffusion(type("M", (), {"self_condition": True, "pred_var": False})(), image_size=64, channels=8, pred_mode="x_start", loss_type="l1",
                               device=DEV, clamp_range=(0, 1))
    dif.set_new_noise_schedule(betas=dp.make_beta_schedule("cosine", 500), device=DEV)
    row("ddpm_step (2048 patches, injected noise)", xb.numel() * 20, lambda: dif._step("ddpm", xb, ob, cb, 250, noise=nz))
    row("ddpm_step (in-kernel Philox noise)", xb.numel() * 16, lambda: dif._step("ddpm", xb, ob, cb, 250))
    row("ddim_step (eta 0)", xb.numel() * 12, lambda: dif._step("ddim", xb, ob, cb, 10, clip=False))
    t = torch.randint(0, 500, (2048,), device=DEV)
    row("q_sample", xb.numel() * 12, lambda: dif.q_sample(xb, t, nz))
    m1, m2 = torch.empty_like(xb), torch.empty_like(xb)
    st = torch.cuda.current_stream().cuda_stream
    row("dpmpp_step order 2", xb.numel() * 20, lambda: _lib.launch("ddif_dpmpp_step_t", st, x=xb.data_ptr(), model_out=ob.data_ptr(), m_cur=m1.data_ptr(),
        m_prev1=m2.data_ptr(), m_prev2=m2.data_ptr(), time_out=None, n=xb.numel(), batch=2048, order=2, model_type=0, alpha_t=0.5, sigma_t=0.8, cx=0.9, ca=0.1,
        cb=0.05, cc=0.0, inv_r0=1.0, inv_r1=1.0, k1=0.5, k2=0.5, t_next_in=0.0, predict=0))
    row("dpm_single stage (mode 1)", xb.numel() * 24, lambda: _lib.launch("ddif_dpm_single_t", st, x_base=nz.data_ptr(), x_eval=xb.data_ptr(), model_out=ob.data_ptr(),
        m_cur=m1.data_ptr(), m_a=m2.data_ptr(), x_out=xb.data_ptr(), time_out=None, n=xb.numel(), batch=2048, model_type=0, predict=0, mode=1, alpha_e=0.5,
        sigma_e=0.8, c0=0.9, c1=0.1, c2=0.05, t_next_in=0.0))
    gt, out = torch.rand(2048, 8, 64, 64, device=DEV), torch.rand(2048, 8, 64, 64, device=DEV)
    row("metrics partial sums (2 passes over gt, out)", gt.numel() * 16, lambda: dm.image_sums(gt, out))
    emit(section="memory-bound kernels vs HBM roofline", peak_gbs=PEAK, rows=rows)


if __name__ == "__main__":
    which = sys.argv[1:] or ["config1", "config3", "config4", "kernels"]
    for w in which:
        globals()[w]()
