"""Final-image parity at the HEADLINE horizons against the reference's own full runs (tests/golden/make_golden3.py):

  * BASELINE configs[1]: full DDPM sampling, cosine T = 500, one WV3 patch, identical injected noise — final fused image
    clip(x + lms, 0, 1) within 0.1 dB PSNR and 0.05 SAM / ERGAS of the reference's (north_star), plus the trajectory at the
    reference's own intermediates (i = 459, 255, 102, 51) so that an accumulation problem is localised in time
  * the reference's production sampler DDIM-25 on the T = 500 schedule (diffusion_engine.py:445)
  * BASELINE configs[2]: a GF2 scene tiled into 64x64 patches, DPM-Solver++ 2M-25, against the reference run ON THE SAME TILES
  * BASELINE configs[3] shape: one CAVE patch, DDIM-25
  * the same patch inside a batch of 256 (bench shape) gives the same final image as alone (per-sample independence at full horizon)
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import dif_pan_b200 as dp  # noqa: E402
from dif_pan_b200 import synth  # noqa: E402
from dif_pan_b200.scene import stitch_tiles, tile_scene  # noqa: E402
from oracle import metrics_oracle  # noqa: E402

DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
torch.set_grad_enabled(False)

TOL_PSNR, TOL_SAM, TOL_ERGAS = 0.1, 0.05, 0.05   # north_star: final fused images within 0.1 dB PSNR and 0.05 SAM / ERGAS


def _net(dataset):
    kw = synth.unet_kwargs(dataset)
    net = dp.UNetSR3(**kw)
    net.load_state_dict(synth.make_state_dict(0, **kw))
    return net.to(DEV).eval()


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def _metrics_close(tag, out, ref, gt):
    m1, m2 = metrics_oracle.batch_metrics(gt, out), metrics_oracle.batch_metrics(gt, ref)
    d = {k: m1[k] - m2[k] for k in ("PSNR", "SAM", "ERGAS")}
    print(f"[{tag}] cuda {m1}  reference {m2}  delta {d}")
    assert abs(d["PSNR"]) <= TOL_PSNR and abs(d["SAM"]) <= TOL_SAM and abs(d["ERGAS"]) <= TOL_ERGAS, (tag, d)


def _dif(net, channels, T=500):
    d = dp.GaussianDiffusion(net, image_size=64, channels=channels, pred_mode="x_start", loss_type="l1", device=DEV, clamp_range=(0, 1))
    d.set_new_noise_schedule(betas=dp.make_beta_schedule("cosine", T), device=DEV)
    return d.to(DEV)


def _noises(seed, n, shape):
    gen = torch.Generator().manual_seed(seed)
    return [torch.randn(*shape, generator=gen) for _ in range(n)]


def test_full_ddpm_T500_final_image_and_trajectory():
    g = np.load(os.path.join(GOLDEN, "full_ddpm_T500.npz"))
    T = int(g["T"])
    data = synth.make_batch("wv3", 1, seed=int(g["data_seed"]))
    cond, lms, gt = data["cond"].to(DEV), data["lms"], data["hr"]
    noises = [n.to(DEV) for n in _noises(int(g["noise_seed"]), T + 1, (1, 8, 64, 64))]
    net = _net("wv3")
    ret = _dif(net, 8, T)(cond, mode="ddpm_sample", continous=True, noise=noises).cpu()
    assert ret.shape[0] == int(g["n_snapshots"])                       # the reference's `continous` layout: x_T + every 51st step
    snap = {459: 1, 255: 5, 102: 8, 51: 9}
    for k, i in enumerate(g["inter_i"]):
        e = _rel(ret[snap[int(i)]], torch.tensor(g["inter"][k]))
        print(f"[ddpm500] x after step i={int(i)}: rel err {e:.4g}")
        assert e < 5e-2, (int(i), e)
    out, ref = ret[-1:], torch.tensor(g["out"])
    print("[ddpm500] final rel", _rel(out, ref))
    fuse = lambda s: (s + lms).clip(0, 1)
    _metrics_close("ddpm500", fuse(out), fuse(ref), gt)
    assert float((fuse(out) - fuse(ref)).abs().max()) < 0.1

    # the same patch as sample 77 of a bench-shaped batch of 256 (different noise for the other samples)
    B = 256
    gen = torch.Generator(device=DEV).manual_seed(5)
    cond_b = synth.make_batch("wv3", 16, seed=31)["cond"].to(DEV).repeat(16, 1, 1, 1)
    cond_b[77] = cond[0]
    big = []
    for n in noises:
        t = torch.randn(B, 8, 64, 64, generator=gen, device=DEV)
        t[77] = n[0]
        big.append(t)
    out_b = _dif(net, 8, T)(cond_b, mode="ddpm_sample", noise=big)[77:78].cpu()
    print("[ddpm500 in B=256] rel vs alone", _rel(out_b, out), "vs reference", _rel(out_b, ref))
    _metrics_close("ddpm500 B=256", fuse(out_b), fuse(ref), gt)


def test_production_ddim25_on_T500():
    g = np.load(os.path.join(GOLDEN, "full_ddim25_T500.npz"))
    data = synth.make_batch("wv3", 1, seed=int(g["data_seed"]))
    cond, lms, gt = data["cond"].to(DEV), data["lms"], data["hr"]
    noises = [n.to(DEV) for n in _noises(int(g["noise_seed"]), 26, (1, 8, 64, 64))]
    d = _dif(_net("wv3"), 8, 500)
    out = d(cond, mode="ddim_sample", section_counts="ddim25", noise=noises).cpu()
    assert d.num_timesteps == 25
    ref = torch.tensor(g["out"])
    print("[ddim25] final rel", _rel(out, ref))
    assert _rel(out, ref) < 3e-2
    _metrics_close("ddim25", (out + lms).clip(0, 1), (ref + lms).clip(0, 1), gt)


def test_cave_ddim25():
    g = np.load(os.path.join(GOLDEN, "cave_ddim25.npz"))
    data = synth.make_batch("cave", 1, seed=int(g["data_seed"]))
    cond, lms, gt = data["cond"].to(DEV), data["lms"], data["hr"]
    noises = [n.to(DEV) for n in _noises(int(g["noise_seed"]), 26, (1, 31, 64, 64))]
    out = _dif(_net("cave"), 31, 500)(cond, mode="ddim_sample", section_counts="ddim25", noise=noises).cpu()
    ref = torch.tensor(g["out"])
    print("[cave ddim25] final rel", _rel(out, ref))
    assert _rel(out, ref) < 3e-2
    _metrics_close("cave ddim25", (out + lms).clip(0, 1), (ref + lms).clip(0, 1), gt)


def test_gf2_scene_tiles_match_reference_on_same_tiles():
    """configs[2] at test size: 128x128 GF2 scene -> four 64x64 tiles (ddif_tile_f32) -> DPM-Solver++ 2M, 25 steps, batched over the
    tiles -> clip(+lms) -> stitch; the reference ran the same solver on each of the same tiles (tests/golden/make_golden3.py)."""
    g = np.load(os.path.join(GOLDEN, "tiles_gf2_dpm25.npz"))
    data = synth.make_batch("gf2", 1, size=128, seed=int(g["data_seed"]))
    cond, gt = data["cond"].to(DEV), data["hr"]
    tiles = tile_scene(cond, 64, 0)
    assert tiles.shape == (4, 12, 64, 64)
    x_T = torch.randn(4, 4, 64, 64, generator=torch.Generator().manual_seed(int(g["noise_seed"]))).to(DEV)
    net = _net("gf2")
    d = _dif(net, 4, 500)
    ns = dp.NoiseScheduleVP("discrete", betas=d.betas)
    wm = dp.model_wrapper(net, ns, model_type="x_start", guidance_type="classifier-free", condition=tiles, guidance_scale=1.0)
    out = dp.DPM_Solver(wm, ns, algorithm_type="dpmsolver++").sample(x_T, steps=25, order=2, skip_type="time_uniform", method="multistep")
    ref = torch.tensor(g["out_tiles"])
    for k in range(4):
        print(f"[gf2 tile {k}] rel {_rel(out[k].cpu(), ref[k]):.4g}")
    sr = stitch_tiles(dp.fuse_output(out, tiles), (128, 128), 0).cpu()
    _metrics_close("gf2 tiles dpm25", sr, torch.tensor(g["sr_scene"]), gt)
    tile_gt = torch.cat([gt[:, :, y:y + 64, x:x + 64] for y in (0, 64) for x in (0, 64)], 0)
    for k in range(4):  # and tile by tile
        _metrics_close(f"gf2 tile {k}", (out[k:k + 1].cpu() + tiles[k:k + 1, :4].cpu()).clip(0, 1),
                       (ref[k:k + 1] + tiles[k:k + 1, :4].cpu()).clip(0, 1), tile_gt[k:k + 1])
