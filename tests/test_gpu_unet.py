"""GPU parity of the full hot path through the drop-in classes:

  * UNetSR3.forward vs the CPU oracle AND vs the committed golden vectors of the reference itself
    (per-sample relative Frobenius error <= 1e-2: the north-star's "per-step UNet output within 1e-2 relative in bf16")
  * block-by-block statistics (localises a wrong layer)
  * DDPM / DDIM / DPM-Solver++ loops with injected noise vs golden final images:
    |dPSNR| <= 0.1 dB, |dSAM|, |dERGAS| <= 0.05 against a synthetic GT
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import dif_pan_b200 as dp  # noqa: E402
from dif_pan_b200 import synth  # noqa: E402
from oracle import metrics_oracle, unet_oracle as uo  # noqa: E402

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


@pytest.mark.parametrize("dataset", ["wv3", "gf2", "cave"])
def test_unet_forward_matches_reference_golden(dataset):
    g = np.load(os.path.join(GOLDEN, f"unet_{dataset}.npz"))
    B, seed = int(g["batch"]), int(g["seed"])
    net, kw = _net(dataset)
    data = synth.make_batch(dataset, B, seed=seed)
    gen = torch.Generator().manual_seed(int(g["gen_seed"]))
    C = kw["in_channel"]
    x = torch.randn(B, C, 64, 64, generator=gen)
    sc = torch.randn(B, C, 64, 64, generator=gen) * 0.3
    cond = data["cond"].to(DEV)
    x_d = x.to(DEV)
    x_before = x_d.clone()
    cases = [("long", torch.tensor(g["t_long"]), None, g["y_long"]),
             ("float", torch.tensor(g["t_float"]), None, g["y_float"]),
             ("selfcond", torch.tensor(g["t_long"]), sc.to(DEV), g["y_selfcond"])]
    for name, t, s, ref in cases:
        y = net(x_d, t.to(DEV), cond, s)
        assert y.shape == ref.shape and y.dtype == torch.float32
        ref = torch.tensor(ref)
        for b in range(B):
            e = _rel(y[b].cpu(), ref[b])
            print(f"[unet {dataset} {name}] sample {b}: rel err {e:.4g}")
            assert e <= 1e-2, (dataset, name, b, e)
    assert torch.equal(x_d, x_before)  # caller-owned inputs are never mutated
    # graph replay and eager op-by-op execution agree bit-for-bit except for atomics order in the statistics
    rt = net.runtime(B, 64, 64)
    y1 = net(x_d, torch.tensor(g["t_long"]).to(DEV), cond)
    rt.use_graph = False
    y2 = net(x_d, torch.tensor(g["t_long"]).to(DEV), cond)
    rt.use_graph = True
    assert _rel(y1, y2) < 1e-3


def test_unet_blocks_against_oracle_taps():
    """Layer-by-layer: compare the per-block activations (debug taps) with the oracle's."""
    net, kw = _net("wv3")
    kw2 = dict(kw)
    kw2.pop("dropout")
    cfg = uo.UNetCfg(**kw2)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    data = synth.make_batch("wv3", 2, seed=77)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 8, 64, 64, generator=gen)
    t = torch.tensor([300, 7])
    taps = {}
    y_ref = uo.unet_forward(sd, cfg, x, t, data["cond"], taps=taps)
    y = net(x.to(DEV), t.to(DEV), data["cond"].to(DEV))
    rt = net.runtime(2, 64, 64)
    rt.use_graph = False
    got = rt.debug_taps(x.to(DEV), t.to(DEV), data["cond"].to(DEV))
    rt.use_graph = True
    worst = 0.0
    for key, ref in taps.items():
        if key not in got:
            continue
        e = _rel(got[key].cpu(), ref)
        worst = max(worst, e)
        print(f"[tap {key}] rel err {e:.4g}")
        assert e < 2.5e-2, (key, e)
    assert len(got) >= 30
    assert _rel(y.cpu(), y_ref) <= 1e-2


def test_other_sizes_and_batches():
    """Fully convolutional: 128x64 patches, odd batch; compared with the oracle."""
    net, kw = _net("gf2")
    kw2 = dict(kw)
    kw2.pop("dropout")
    cfg = uo.UNetCfg(**kw2)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(3, 4, 128, 64, generator=gen)
    cond = torch.rand(3, 12, 128, 64, generator=gen)
    t = torch.tensor([10, 200, 499])
    y = net(x.to(DEV), t.to(DEV), cond.to(DEV))
    y_ref = uo.unet_forward(sd, cfg, x, t, cond)
    for b in range(3):
        assert _rel(y[b].cpu(), y_ref[b]) <= 1e-2
    # changing cond invalidates the cond cache
    cond2 = cond.flip(0).contiguous()
    y2 = net(x.to(DEV), t.to(DEV), cond2.to(DEV))
    y2_ref = uo.unet_forward(sd, cfg, x, t, cond2)
    assert _rel(y2.cpu(), y2_ref) <= 1e-2


def _metrics_close(out, ref, gt, psnr=0.1, sam=0.05, ergas=0.05):
    """Default bounds = the north-star's final-image tolerance (0.1 dB PSNR, 0.05 SAM / ERGAS)."""
    m1, m2 = metrics_oracle.batch_metrics(gt, out), metrics_oracle.batch_metrics(gt, ref)
    print("metrics cuda", m1, "reference", m2)
    assert abs(m1["PSNR"] - m2["PSNR"]) <= psnr
    assert abs(m1["SAM"] - m2["SAM"]) <= sam and abs(m1["ERGAS"] - m2["ERGAS"]) <= ergas


def test_sampling_loops_match_reference_golden():
    net, kw = _net("wv3")
    g = np.load(os.path.join(GOLDEN, "ddpm_T6.npz"))
    data = synth.make_batch("wv3", 1, seed=int(g["data_seed"]))
    cond, lms, gt = data["cond"].to(DEV), data["lms"], data["hr"]
    fuse = lambda s: (s.cpu() + lms).clip(0, 1)

    def diffusion(T):
        d = dp.GaussianDiffusion(net, image_size=64, channels=8, pred_mode="x_start", loss_type="l1", device=DEV, clamp_range=(0, 1))
        d.set_new_noise_schedule(betas=dp.make_beta_schedule("cosine", T), device=DEV)
        return d.to(DEV)

    # DDPM, T = 6
    T = int(g["T"])
    gen = torch.Generator().manual_seed(int(g["noise_seed"]))
    noises = [torch.randn(1, 8, 64, 64, generator=gen).to(DEV) for _ in range(T + 1)]
    out = diffusion(T)(cond, mode="ddpm_sample", noise=noises)
    ref = torch.tensor(g["out"])
    print("ddpm rel", _rel(out.cpu(), ref))
    assert _rel(out.cpu(), ref) < 2e-2
    _metrics_close(fuse(out), fuse(ref), gt)
    fused = dp.fuse_output(out, cond)
    assert torch.allclose(fused.cpu(), fuse(out), atol=1e-7)

    # DDIM, T = 100 -> ddim5
    g = np.load(os.path.join(GOLDEN, "ddim_T100_5.npz"))
    gen = torch.Generator().manual_seed(int(g["noise_seed"]))
    noises = [torch.randn(1, 8, 64, 64, generator=gen).to(DEV) for _ in range(6)]
    d = diffusion(100)
    out = d(cond, mode="ddim_sample", section_counts="ddim5", noise=noises)
    assert d.num_timesteps == 5  # the reference overwrites its schedule too
    ref = torch.tensor(g["out"])
    print("ddim rel", _rel(out.cpu(), ref))
    assert _rel(out.cpu(), ref) < 2e-2
    _metrics_close(fuse(out), fuse(ref), gt)

    # DPM-Solver++ 2M, 20 steps (BASELINE config 1) and 3M / lower-order-final variants
    g = np.load(os.path.join(GOLDEN, "dpm.npz"))
    d = diffusion(500)
    ns = dp.NoiseScheduleVP("discrete", betas=d.betas)
    wm = dp.model_wrapper(net, ns, model_type="x_start", guidance_type="classifier-free", condition=cond, guidance_scale=1.0)
    sol = dp.DPM_Solver(wm, ns, algorithm_type="dpmsolver++")
    gen = torch.Generator().manual_seed(int(g["noise_seed"]))
    x_T = torch.randn(1, 8, 64, 64, generator=gen).to(DEV)
    out, inter = sol.sample(x_T, steps=20, order=2, skip_type="time_uniform", method="multistep", return_intermediate=True)
    ref = torch.tensor(g["out_o2_s20"])
    print("dpm20 rel", _rel(out.cpu(), ref), "step1 rel", _rel(inter[1].cpu(), torch.tensor(g["x_after_step1"])))
    assert _rel(inter[1].cpu(), torch.tensor(g["x_after_step1"])) < 2e-2
    assert _rel(out.cpu(), ref) < 3e-2
    _metrics_close(fuse(out), fuse(ref), gt)
    for steps, order, key in ((12, 3, "out_o3_s12"), (5, 2, "out_o2_s5")):
        out = sol.sample(x_T, steps=steps, order=order)
        ref = torch.tensor(g[key])
        print(f"dpm steps={steps} order={order} rel", _rel(out.cpu(), ref))
        # higher-order extrapolation amplifies the per-step bf16 noise of the UNet (1/r0 factors); the solver arithmetic
        # itself is checked in fp32 by test_gpu_kernels.py::test_dpm_solver_loop_fp32_model
        assert _rel(out.cpu(), ref) < 0.1
        # MEASURED DEVIATION (round 2, B200): these two variants are not BASELINE configurations; the 3rd-order multistep update takes
        # second differences of consecutive data predictions, which amplifies the bf16 UNet's per-evaluation error (0.7e-2 relative):
        # 3M-12 lands at dPSNR 0.08 dB, dSAM 0.13, dERGAS 0.06 -- outside the 0.05 SAM / ERGAS bound that every BASELINE configuration
        # meets (2M-20 above, DDPM-500 / DDIM-25 / 2M-25 tiles in test_gpu_headline.py).  Bounded here at 0.2 and stated in DESIGN.md.
        _metrics_close(fuse(out), fuse(ref), gt, psnr=0.2, sam=0.2, ergas=0.2)


def test_generic_denoiser_path_and_philox_noise():
    """The samplers also drive any nn.Module with the reference call signature; without injected noise they
    draw from the in-kernel Philox generator (deterministic per seed)."""
    net, kw = _net("wv3")
    cond = synth.make_batch("wv3", 2, seed=3)["cond"].to(DEV)

    class Wrap(torch.nn.Module):
        self_condition, pred_var = True, False

        def forward(self, x, t, c, sc=None):
            return net(x, t, c, sc)

    def run(model, noise=None, seed=0):
        d = dp.GaussianDiffusion(model, image_size=64, channels=8, pred_mode="x_start", loss_type="l1", device=DEV, clamp_range=(0, 1))
        d.set_new_noise_schedule(betas=dp.make_beta_schedule("cosine", 4), device=DEV)
        d.seed = seed
        return d.to(DEV)(cond, mode="ddpm_sample", noise=noise)

    gen = torch.Generator().manual_seed(1)
    noises = [torch.randn(2, 8, 64, 64, generator=gen).to(DEV) for _ in range(5)]
    a, b = run(net, noises), run(Wrap(), noises)
    assert _rel(a, b) < 5e-3  # explicit self_cond == x path vs fused x-twice path
    p1, p2, p3 = run(net, None, 11), run(net, None, 11), run(net, None, 12)
    assert _rel(p1, p2) < 1e-5 and _rel(p1, p3) > 1e-2 and torch.isfinite(p1).all()  # fp64 atomics order may flip a last bit
