"""GPU parity of the rows around the core path: DPM-Solver variants (both algorithm types, multistep + singlestep), the training
objective (forward), on-device metrics, the fused conditioning prep, scene tiling / stitching and the scene driver.
Everything goes through the C ABI (dif_pan_b200._lib); the oracle and the reference-made fixtures are the checkers."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import dif_pan_b200 as dp  # noqa: E402
from dif_pan_b200 import synth, metrics as dmetrics  # noqa: E402
from oracle import metrics_oracle, sampler_oracle as so, unet_oracle as uo  # noqa: E402
from test_oracle_variants import DPM_CASES, case_key, dpm_inputs, loss_inputs, rel_max  # noqa: E402

DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
torch.set_grad_enabled(False)


def _net(dataset):
    kw = synth.unet_kwargs(dataset)
    net = dp.UNetSR3(**kw)
    net.load_state_dict(synth.make_state_dict(0, **kw))
    return net.to(DEV).eval(), kw


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


# ---- DPM-Solver variants -----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", DPM_CASES, ids=[case_key(c) for c in DPM_CASES])
def test_dpm_variants_match_reference_golden(case):
    """Same analytic fp32 denoiser as the fixture generator, so only the solver arithmetic (host scalars + the fused
    ddif_dpm_single_f32 / ddif_dpmpp_step_f32 kernels) is under test.  Also at batch 3, where the reference itself cannot run
    (its [B] x [B,C,H,W] broadcast, dpm_solver.py:299-300): checked against the oracle."""
    g = np.load(os.path.join(GOLDEN, "dpm_variants.npz"))
    x_T, cond = dpm_inputs(int(g["seed"]))
    method, order, steps, skip, algo, mtype = case
    betas = torch.as_tensor(so.make_beta_schedule("cosine", 500), dtype=torch.float32)
    ns = dp.NoiseScheduleVP("discrete", betas=betas.to(DEV))
    wm = dp.model_wrapper(so.analytic_denoiser, ns, model_type=mtype, guidance_type="classifier-free", condition=cond.to(DEV), guidance_scale=1.0)
    got = dp.DPM_Solver(wm, ns, algorithm_type=algo).sample(x_T.to(DEV), steps=steps, order=order, skip_type=skip, method=method)
    e = rel_max(got.cpu().numpy(), g[case_key(case)])
    print(f"{case_key(case)}: max err / max |ref| = {e:.3g}")
    assert e <= 2e-3  # trajectories divide by alpha_T ~ 1e-4: fp32 rounding (and GPU vs CPU sin/cos) is amplified ~1e3x at the first step
    gen = torch.Generator().manual_seed(77)
    x3, c3 = torch.randn(3, 4, 16, 16, generator=gen), torch.rand(3, 12, 16, 16, generator=gen)
    ref = so.dpm_sample(so.analytic_denoiser, so.VPSchedule(betas), x3.clone(), c3, steps=steps, order=order, skip_type=skip, method=method,
                        algorithm=algo, model_type=mtype)
    wm3 = dp.model_wrapper(so.analytic_denoiser, ns, model_type=mtype, guidance_type="classifier-free", condition=c3.to(DEV), guidance_scale=1.0)
    got3 = dp.DPM_Solver(wm3, ns, algorithm_type=algo).sample(x3.to(DEV), steps=steps, order=order, skip_type=skip, method=method)
    assert rel_max(got3.cpu().numpy(), ref.numpy()) <= 2e-3


def test_singlestep_unet_fast_path_equals_generic_path():
    """With a dif_pan_b200.UNetSR3 the singlestep loop runs in place on the model's device buffers (graph replay + one fused kernel per
    evaluation, time label written by the kernel); through a plain wrapper it uses the generic path.  Same kernels, same numbers."""
    net, kw = _net("gf2")
    cond = synth.make_batch("gf2", 2, seed=5)["cond"].to(DEV)
    betas = torch.as_tensor(so.make_beta_schedule("cosine", 500), dtype=torch.float32)
    ns = dp.NoiseScheduleVP("discrete", betas=betas.to(DEV))
    gen = torch.Generator().manual_seed(8)
    x_T = torch.randn(2, 4, 64, 64, generator=gen).to(DEV)

    def plain(x, t, c):
        return net(x, t, c)

    outs = []
    for model in (net, plain):
        wm = dp.model_wrapper(model, ns, model_type="x_start", guidance_type="classifier-free", condition=cond, guidance_scale=1.0)
        outs.append(dp.DPM_Solver(wm, ns).sample(x_T.clone(), steps=7, order=3, skip_type="time_uniform", method="singlestep"))
    assert torch.isfinite(outs[0]).all()
    assert _rel(outs[0], outs[1]) < 1e-4


# ---- training objective (forward) ------------------------------------------------------------------------------------------------
class _LossModel(torch.nn.Module):
    self_condition, pred_var = True, False

    def forward(self, x, t, cond=None, self_cond=None):
        y = so.analytic_denoiser(x, t.to(torch.float32), cond)
        return y if self_cond is None else y + 0.25 * self_cond


@pytest.mark.parametrize("pred_mode", ["x_start", "noise", "pred_v"])
@pytest.mark.parametrize("loss_type", ["l1", "l2"])
@pytest.mark.parametrize("sc", [0, 1])
def test_p_losses_match_reference_golden(pred_mode, loss_type, sc):
    g = np.load(os.path.join(GOLDEN, "losses.npz"))
    x0, nz, cond = loss_inputs(int(g["seed"]))
    dif = dp.GaussianDiffusion(_LossModel().eval(), image_size=16, channels=8, pred_mode=pred_mode, loss_type=loss_type, device=DEV,
                               clamp_range=(0, 1), p2_loss_weight_gamma=0.5 if loss_type == "l2" else 0.0)
    dif.set_new_noise_schedule(betas=dp.make_beta_schedule("cosine", 500), device=DEV)
    loss, recon = dif(x0.to(DEV), "train", noise=nz.to(DEV), cond=cond.to(DEV), t=torch.as_tensor(g["t"]).to(DEV),
                      self_cond_draw=0.1 if sc else 0.9)
    ref = float(g[f"loss_{pred_mode}_{loss_type}_{sc}"])
    assert abs(float(loss) - ref) <= 5e-5 * abs(ref), (float(loss), ref)
    assert rel_max(recon.cpu().numpy(), g[f"recon_{pred_mode}_{loss_type}_{sc}"]) <= 2e-5


def test_p_losses_with_cuda_unet_and_random_draws():
    """The objective through the CUDA UNet (eval mode, no autograd): finite, equal to the oracle's objective on the same draws."""
    net, kw = _net("wv3")
    d = synth.make_batch("wv3", 2, seed=9)
    dif = dp.GaussianDiffusion(net, image_size=64, channels=8, pred_mode="x_start", loss_type="l1", device=DEV, clamp_range=(0, 1))
    dif.set_new_noise_schedule(betas=dp.make_beta_schedule("cosine", 500), device=DEV)
    x0 = (d["hr"] - d["lms"]).to(DEV)
    gen = torch.Generator().manual_seed(4)
    nz = torch.randn(2, 8, 64, 64, generator=gen)
    t = torch.tensor([20, 400])
    loss, recon = dif(x0, "train", noise=nz.to(DEV), cond=d["cond"].to(DEV), t=t.to(DEV), self_cond_draw=0.2)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    kw2 = dict(kw); kw2.pop("dropout")
    cfg = uo.UNetCfg(**kw2)
    model = lambda x, tt, c, sc: uo.unet_forward(sd, cfg, x, tt, c, self_cond=sc)
    sb = so.schedule_buffers(so.make_beta_schedule("cosine", 500))
    ref_loss, ref_recon = so.p_losses(sb, model, x0.cpu(), t, nz, d["cond"], "x_start", "l1", True)
    print("loss cuda", float(loss), "oracle", float(ref_loss))
    assert abs(float(loss) - float(ref_loss)) <= 1e-2 * abs(float(ref_loss))
    assert _rel(recon.cpu(), ref_recon) <= 1.5e-2
    loss2, _ = dif(x0, "train", cond=d["cond"].to(DEV))  # internal draws
    assert torch.isfinite(loss2)
    with torch.enable_grad():  # with autograd on, the loss carries a graph (training.py; gradient parity: tests/test_gpu_training.py)
        loss3, _ = dif(x0, "train", noise=nz.to(DEV), cond=d["cond"].to(DEV), t=t.to(DEV), self_cond_draw=0.2)
        assert loss3.requires_grad and abs(float(loss3) - float(loss)) <= 1e-2 * abs(float(loss))


# ---- metrics -------------------------------------------------------------------------------------------------------------------
def test_metrics_match_reference_golden_and_oracle():
    g = np.load(os.path.join(GOLDEN, "metrics_cc.npz"))
    d = synth.make_batch("wv3", 2, seed=int(g["seed"]))
    gen = torch.Generator().manual_seed(int(g["noise_seed"]))
    out = (d["hr"] + 0.03 * torch.randn(d["hr"].shape, generator=gen)).clamp(0, 1)
    m = dmetrics.per_image_metrics(d["hr"].to(DEV), out.to(DEV))
    got = torch.stack([m["SAM"], m["ERGAS"], m["PSNR"], m["CC"]], 1).cpu().numpy()
    np.testing.assert_allclose(got, g["sam_ergas_psnr_cc"], rtol=2e-5)
    # other shapes (CAVE bands, non-square, degenerate pixels with a zero spectrum) against the oracle
    gen = torch.Generator().manual_seed(2)
    a = torch.rand(3, 31, 40, 24, generator=gen)
    b = (a + 0.05 * torch.randn(a.shape, generator=gen)).clamp(0, 1)
    a[0, :, 3, 4] = 0.0
    b[1, :, 7, 7] = 0.0
    ref = metrics_oracle.batch_metrics(a, b)
    got = dmetrics.batch_metrics(a.to(DEV), b.to(DEV))
    for k in ("SAM", "ERGAS", "PSNR", "CC"):
        assert abs(got[k] - ref[k]) <= 2e-5 * max(1.0, abs(ref[k])), (k, got[k], ref[k])
    acc = dmetrics.AnalysisPanAcc()
    acc(a[:1].to(DEV), b[:1].to(DEV))
    ave = acc(a[1:].to(DEV), b[1:].to(DEV))
    assert abs(ave["PSNR"] - ref["PSNR"]) <= 1e-4
    with pytest.raises(RuntimeError):
        dmetrics.batch_metrics(a, b)


# ---- fused conditioning prep --------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ds", ["wv3", "gf2", "cave"])
def test_make_cond_fused(ds):
    d = synth.make_batch(ds, 2, seed=31)
    spec = synth.DATASETS[ds]
    lms_dn, pan_dn = d["lms_dn"].float().to(DEV), d["pan_dn"].float().to(DEV)
    cond, wav = dp.make_cond(lms_dn, pan_dn, spec.division, spec.wavelet_order, return_wavelets=True)
    wav3 = dp.wavelet_channels(lms_dn, pan_dn, spec.division, spec.wavelet_order)
    cond3 = dp.assemble_cond(lms_dn / spec.division, pan_dn / spec.division, wav3)
    assert float((wav - wav3).abs().max()) <= 2.5e-7                     # reciprocal multiply vs the DWT kernel's IEEE division: <= 1.5 ulp
    assert float((cond - cond3).abs().max()) <= 5e-7                     # + bilinear taps: FMA contraction may differ by an ulp
    assert float((cond.cpu() - d["cond"]).abs().max()) <= 2e-6           # and equal to the float64 dataset pipeline
    assert float((wav.cpu() - d["wavelets"]).abs().max()) <= 1e-6
    big = dp.make_cond(torch.rand(1, 4, 512, 512, device=DEV) * 1023, torch.rand(1, 1, 512, 512, device=DEV) * 1023, 1023.0)
    assert big.shape == (1, 12, 512, 512) and torch.isfinite(big).all()
    with pytest.raises(ValueError):
        dp.make_cond(lms_dn[:, :, :63], pan_dn[:, :, :63], 1.0)


# ---- scene tiling / stitching ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("overlap", [0, 16, 32])
def test_tile_and_stitch(overlap):
    gen = torch.Generator().manual_seed(3)
    p, s = 64, 64 - overlap
    H, W = p + 2 * s, p + 3 * s
    x = torch.rand(2, 5, H, W, generator=gen).to(DEV)
    tiles = dp.tile_scene(x, p, overlap)
    ref = x.unfold(2, p, s).unfold(3, p, s).permute(0, 2, 3, 1, 4, 5).reshape(-1, 5, p, p)
    assert torch.equal(tiles, ref)
    back = dp.stitch_tiles(tiles, (H, W), overlap)
    assert float((back - x).abs().max()) <= 1e-6                         # averaging identical overlaps returns the scene
    noisy = tiles + 0.1 * torch.randn(tiles.shape, generator=gen).to(DEV)
    ny, nx = (H - p) // s + 1, (W - p) // s + 1
    cols = noisy.reshape(2, ny * nx, 5 * p * p).permute(0, 2, 1)
    num = F.fold(cols, (H, W), kernel_size=p, stride=s)
    den = F.fold(torch.ones_like(cols), (H, W), kernel_size=p, stride=s)
    assert float((dp.stitch_tiles(noisy, (H, W), overlap) - num / den).abs().max()) <= 1e-5
    with pytest.raises(ValueError):
        dp.tile_scene(x, 64, 7)


# ---- scene driver ----------------------------------------------------------------------------------------------------------------------
def test_scene_driver_whole_and_tiled():
    """GF2-shaped 128x128 scene, DDIM-5 (T=100): whole-scene sampling (the reference's test_fn feeds scenes whole, diffusion_engine.py:441-447)
    against the oracle's DDIM loop on the same scene and noise; tiled sampling against the driver run tile by tile."""
    net, kw = _net("gf2")
    d = synth.make_batch("gf2", 1, size=128, seed=41)
    spec = synth.DATASETS["gf2"]
    lms_dn, pan_dn = d["lms_dn"].float().to(DEV), d["pan_dn"].float().to(DEV)
    cond = dp.make_cond(lms_dn, pan_dn, spec.division, "pan")
    gen = torch.Generator().manual_seed(6)
    noises = [torch.randn(1, 4, 128, 128, generator=gen) for _ in range(6)]
    dif = dp.GaussianDiffusion(net, image_size=128, channels=4, pred_mode="x_start", loss_type="l1", device=DEV, clamp_range=(0, 1))
    dif.set_new_noise_schedule(betas=dp.make_beta_schedule("cosine", 100), device=DEV)
    sample = dif(cond, mode="ddim_sample", section_counts="ddim5", noise=[n.to(DEV) for n in noises])
    sr = dp.fuse_output(sample, cond).cpu()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    kw2 = dict(kw); kw2.pop("dropout")
    cfg = uo.UNetCfg(**kw2)
    model = lambda x, t, c, sc: uo.unet_forward(sd, cfg, x, t, c, self_cond=sc)
    ref = so.ddim_sample_loop(model, so.make_beta_schedule("cosine", 100), cond.cpu(), 4, noises, section_counts="ddim5")
    ref_sr = (ref + d["lms"]).clip(0, 1)
    m1, m2 = metrics_oracle.batch_metrics(d["hr"], sr), metrics_oracle.batch_metrics(d["hr"], ref_sr)
    print("whole scene: cuda", m1, "oracle", m2)
    assert abs(m1["PSNR"] - m2["PSNR"]) <= 0.1 and abs(m1["SAM"] - m2["SAM"]) <= 0.05 and abs(m1["ERGAS"] - m2["ERGAS"]) <= 0.05
    # driver end to end, whole and tiled (Philox noise): finite, in range, tiled == manual tile loop
    out_w = dp.fuse_scene(net, lms_dn, pan_dn, spec.division, sampler="ddim5", n_timestep=100, seed=3)
    assert out_w.shape == (1, 4, 128, 128) and float(out_w.min()) >= 0 and float(out_w.max()) <= spec.division
    x_T = torch.randn(4, 4, 64, 64, generator=gen).to(DEV)
    out_t = dp.fuse_scene(net, lms_dn, pan_dn, spec.division, sampler="dpm6", patch=64, x_T=x_T)
    tiles = dp.tile_scene(cond, 64, 0)
    manual = dp.fuse_output(dp.sample_cond(net, tiles, 4, "dpm6", x_T=x_T), tiles)
    assert torch.equal(out_t, (dp.stitch_tiles(manual, (128, 128)) * spec.division).clamp(0, spec.division))
    got = dmetrics.batch_metrics(d["hr"].to(DEV), out_t / spec.division)
    assert np.isfinite(list(got.values())).all()


# ---- adaptive DPM-Solver ------------------------------------------------------------------------------------------------------------------
from test_oracle_variants import ADAPTIVE_CASES  # noqa: E402


@pytest.mark.parametrize("order,algo,mtype", ADAPTIVE_CASES)
def test_dpm_adaptive_matches_reference_golden(order, algo, mtype):
    """Step-size control (dpm_solver.py:964-1018): fused stage kernels + one error-norm reduction per iteration; the accept / reject
    sequence must reproduce the reference's, so the final state agrees to the solver kernels' own tolerance."""
    g = np.load(os.path.join(GOLDEN, "dpm_variants.npz"))
    x_T, cond = dpm_inputs(int(g["seed"]))
    betas = torch.as_tensor(so.make_beta_schedule("cosine", 500), dtype=torch.float32)
    ns = dp.NoiseScheduleVP("discrete", betas=betas.to(DEV))
    wm = dp.model_wrapper(so.analytic_denoiser, ns, model_type=mtype, guidance_type="classifier-free", condition=cond.to(DEV), guidance_scale=1.0)
    sol = dp.DPM_Solver(wm, ns, algorithm_type=algo)
    got = sol.sample(x_T.to(DEV), order=order, method="adaptive")
    key = f"adaptive_{order}_{algo.replace('+', 'p')}_{mtype}"
    e = rel_max(got.cpu().numpy(), g[key])
    print(f"{key}: nfe {sol.last_nfe}, max err / max |ref| = {e:.3g}")
    assert e <= 2e-3
    with pytest.raises(ValueError):
        sol.sample(x_T.to(DEV), order=1, method="adaptive")


def test_dpm_adaptive_unet_fast_path():
    net, kw = _net("gf2")
    cond = synth.make_batch("gf2", 2, seed=5)["cond"].to(DEV)
    betas = torch.as_tensor(so.make_beta_schedule("cosine", 500), dtype=torch.float32)
    ns = dp.NoiseScheduleVP("discrete", betas=betas.to(DEV))
    gen = torch.Generator().manual_seed(8)
    x_T = torch.randn(2, 4, 64, 64, generator=gen).to(DEV)
    outs = []
    for model in (net, lambda x, t, c: net(x, t, c)):
        wm = dp.model_wrapper(model, ns, model_type="x_start", guidance_type="classifier-free", condition=cond, guidance_scale=1.0)
        sol = dp.DPM_Solver(wm, ns)
        outs.append(sol.sample(x_T.clone(), order=2, method="adaptive", atol=0.05, rtol=0.1))
        assert 0 < sol.last_nfe < 400
    assert torch.isfinite(outs[0]).all() and _rel(outs[0], outs[1]) < 1e-3


# ---- training-loop surround: multi-tensor EMA / grad clip / AdamW --------------------------------------------------------
def _param_sets(seed):
    """Two UNets with the reference's parameter list (702 entries incl. biases) and synthetic gradients."""
    from dif_pan_b200 import optim as dopt  # noqa: F401
    net, kw = _net("gf2")
    net2 = dp.UNetSR3(**kw).to(DEV)
    gen = torch.Generator().manual_seed(seed)
    for p in net2.parameters():
        p.data = torch.randn(p.shape, generator=gen).to(DEV) * 0.05
    for p in net.parameters():
        p.grad = (torch.randn(p.shape, generator=gen) * 0.01).to(DEV)
    return net, net2


def test_ema_updater_matches_reference_expression():
    from dif_pan_b200.optim import EmaUpdater
    net, ema = _param_sets(3)
    holder = lambda n: type("D", (), {"model": n, "state_dict": n.state_dict, "load_state_dict": n.load_state_dict})()
    up = EmaUpdater(holder(net), holder(ema), decay=0.9999, start_iter=2)
    ref = [pe.data * 0.9999 + p.data * (1 - 0.9999) for p, pe in zip(net.parameters(), ema.parameters())]   # utils/optim_utils.py:49-51
    up.update(5)
    for r, pe in zip(ref, ema.parameters()):
        assert torch.equal(r, pe.data)
    up.update(1)                                                                                             # iteration <= start_iter: copy
    for p, pe in zip(net.parameters(), ema.parameters()):
        assert torch.equal(p.data, pe.data)
    assert len(up.ema_model_state_dict) == len(up.on_fly_model_state_dict) == 702


def test_grad_clip_and_fused_adamw_match_torch():
    from dif_pan_b200.optim import FusedAdamW, grad_clip
    net, _ = _param_sets(4)
    params = list(net.parameters())
    twin = [torch.nn.Parameter(p.data.clone()) for p in params]
    for t, p in zip(twin, params):
        t.grad = p.grad.clone()
    # clip by norm (the engine clips at 0.003, diffusion_engine.py:236) and by value
    n_ref = torch.nn.utils.clip_grad_norm_(twin, max_norm=0.003)
    n_got = grad_clip(params, mode="norm", value=0.003)
    assert abs(float(n_got) - float(n_ref)) <= 1e-5 * float(n_ref)
    for t, p in zip(twin, params):
        assert float((t.grad - p.grad).abs().max()) <= 1e-6 * float(t.grad.abs().max() + 1e-12)
    torch.nn.utils.clip_grad_value_(twin, clip_value=1e-5)
    grad_clip(params, mode="value", value=1e-5)
    for t, p in zip(twin, params):
        assert float(p.grad.abs().max()) <= 1e-5 and float((t.grad - p.grad).abs().max()) <= 1e-11
    # three AdamW steps against torch.optim.AdamW (single-tensor, unfused implementation)
    ref_opt = torch.optim.AdamW(twin, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, foreach=False, fused=False)
    opt = FusedAdamW(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    gen = torch.Generator().manual_seed(9)
    for _ in range(3):
        for t, p in zip(twin, params):
            g = (torch.randn(p.shape, generator=gen) * 0.01).to(DEV)
            t.grad, p.grad = g.clone(), g.clone()
        ref_opt.step()
        opt.step()
    worst = max(float((t.data - p.data).abs().max() / (t.data.abs().max() + 1e-12)) for t, p in zip(twin, params))
    print("AdamW worst relative deviation after 3 steps", worst)
    assert worst <= 2e-6
    with pytest.raises(RuntimeError):
        grad_clip([_cpu_param()], mode="value", value=1.0)
    # a real torch.optim.Optimizer: schedulers attach, and the state round-trips with torch.optim.AdamW's own layout
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[1], gamma=0.2)   # diffusion_engine.py:207-210
    sched.step()
    assert abs(opt.param_groups[0]["lr"] - 2e-4) < 1e-12
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 3.0
    ref_opt.load_state_dict(sd)                                                    # FusedAdamW -> torch AdamW
    opt2 = FusedAdamW(params, lr=1e-3)
    opt2.load_state_dict(ref_opt.state_dict())                                     # torch AdamW -> FusedAdamW
    assert opt2.param_groups[0]["lr"] == ref_opt.param_groups[0]["lr"]
    assert torch.equal(opt2.state[params[5]]["exp_avg_sq"], opt.state[params[5]]["exp_avg_sq"])


def test_raw_pointer_updates_invalidate_the_packed_weights():
    """EmaUpdater / FusedAdamW write parameters through raw device pointers; UNetSR3 keys its packed bf16 weights and captured CUDA
    graph on (data_ptr, _version), so those writers bump the version counters.  The reference's validation flow is exactly
    `ema_updater.update(it)` followed by sampling with the EMA model (diffusion_engine.py:240,273-298)."""
    from dif_pan_b200 import synth
    from dif_pan_b200.optim import EmaUpdater, FusedAdamW
    kw = synth.unet_kwargs("wv3")
    net, ema = dp.UNetSR3(**kw), dp.UNetSR3(**kw)
    net.load_state_dict(synth.make_state_dict(0, **kw))
    ema.load_state_dict(synth.make_state_dict(1, **kw))
    net, ema = net.to(DEV).eval(), ema.to(DEV).eval()

    class Holder:  # EmaUpdater reads `.model.parameters()` (GaussianDiffusion objects in the engine)
        def __init__(self, m):
            self.model = m

    d = synth.make_batch("wv3", 2, seed=11)
    cond = d["cond"].to(DEV)
    x = torch.randn(2, 8, 64, 64, generator=torch.Generator().manual_seed(3)).to(DEV)
    t = torch.tensor([400, 20], device=DEV)
    y_before = ema(x, t, cond)
    up = EmaUpdater(Holder(net), Holder(ema), decay=0.5, start_iter=0)
    up.update(10)                                      # ema <- 0.5 ema + 0.5 net, written through raw pointers
    y_after = ema(x, t, cond)
    fresh = dp.UNetSR3(**kw)
    fresh.load_state_dict({k: v.clone() for k, v in ema.state_dict().items()})
    y_fresh = fresh.to(DEV).eval()(x, t, cond)
    assert float((y_after - y_before).abs().max()) > 1e-3, "the EMA update must change the output"
    assert float((y_after - y_fresh).abs().max()) <= 1e-3 * float(y_fresh.abs().max()), "stale packed weights after EmaUpdater.update"
    # same for an optimizer step
    for p in ema.parameters():
        p.grad = torch.full_like(p, 1e-2)
    FusedAdamW(ema.parameters(), lr=1e-2).step()
    y_opt = ema(x, t, cond)
    fresh.load_state_dict({k: v.clone() for k, v in ema.state_dict().items()})
    y_fresh = fresh(x, t, cond)
    assert float((y_opt - y_after).abs().max()) > 1e-3
    assert float((y_opt - y_fresh).abs().max()) <= 1e-3 * float(y_fresh.abs().max()), "stale packed weights after FusedAdamW.step"


def _cpu_param():
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    return p


from test_oracle_variants import TAYLOR_CASES  # noqa: E402


@pytest.mark.parametrize("case", TAYLOR_CASES, ids=[case_key(c) for c in TAYLOR_CASES])
def test_dpm_taylor_matches_reference_golden(case):
    """solver_type='taylor' only changes host-side scalars of the second-order updates (same fused kernels)."""
    g = np.load(os.path.join(GOLDEN, "dpm_variants.npz"))
    x_T, cond = dpm_inputs(int(g["seed"]))
    method, order, steps, skip, algo, mtype = case
    betas = torch.as_tensor(so.make_beta_schedule("cosine", 500), dtype=torch.float32)
    ns = dp.NoiseScheduleVP("discrete", betas=betas.to(DEV))
    wm = dp.model_wrapper(so.analytic_denoiser, ns, model_type=mtype, guidance_type="classifier-free", condition=cond.to(DEV), guidance_scale=1.0)
    got = dp.DPM_Solver(wm, ns, algorithm_type=algo).sample(x_T.to(DEV), steps=steps, order=order, skip_type=skip, method=method, solver_type="taylor")
    assert rel_max(got.cpu().numpy(), g["taylor_" + case_key(case)]) <= 2e-3
    with pytest.raises(NotImplementedError):
        dp.DPM_Solver(wm, ns, algorithm_type=algo).sample(x_T.to(DEV), steps=9, order=3, method="singlestep", solver_type="taylor")


from test_oracle_variants import DTZ_CASES  # noqa: E402


@pytest.mark.parametrize("case", DTZ_CASES, ids=[case_key(c) for c in DTZ_CASES])
def test_dpm_denoise_to_zero_matches_reference_golden(case):
    g = np.load(os.path.join(GOLDEN, "dpm_variants.npz"))
    x_T, cond = dpm_inputs(int(g["seed"]))
    method, order, steps, skip, algo, mtype = case
    betas = torch.as_tensor(so.make_beta_schedule("cosine", 500), dtype=torch.float32)
    ns = dp.NoiseScheduleVP("discrete", betas=betas.to(DEV))
    wm = dp.model_wrapper(so.analytic_denoiser, ns, model_type=mtype, guidance_type="classifier-free", condition=cond.to(DEV), guidance_scale=1.0)
    got = dp.DPM_Solver(wm, ns, algorithm_type=algo).sample(x_T.to(DEV), steps=steps, order=order, skip_type=skip, method=method, denoise_to_zero=True)
    assert rel_max(got.cpu().numpy(), g["dtz_" + case_key(case)]) <= 2e-3
