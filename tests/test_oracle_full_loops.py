"""Pin the CPU oracle against the reference's FULL-horizon runs (tests/golden/make_golden3.py): the production DDIM-25 on the T=500
cosine schedule, the tiled GF2 scene through DPM-Solver++ 2M-25 and one CAVE patch; the full DDPM T=500 loop is pinned at its first 41
steps by the reference's own intermediate (i = 459) and over all 500 steps when DDIF_SLOW_TESTS=1 (36 s of CPU)."""
import os

import numpy as np
import pytest
import torch

from dif_pan_b200 import synth
from oracle import sampler_oracle as so, unet_oracle as uo

torch.set_grad_enabled(False)


def _model(dataset):
    kw = synth.unet_kwargs(dataset)
    sd = synth.make_state_dict(0, **kw)
    kw.pop("dropout")
    cfg = uo.UNetCfg(**kw)
    return lambda x, t, c, sc: uo.unet_forward(sd, cfg, x, t, c, sc)


def _noises(seed, n, shape):
    gen = torch.Generator().manual_seed(seed)
    return [torch.randn(*shape, generator=gen) for _ in range(n)]


def test_ddim25_on_T500(golden_dir):
    g = np.load(os.path.join(golden_dir, "full_ddim25_T500.npz"))
    cond = synth.make_batch("wv3", 1, seed=int(g["data_seed"]))["cond"]
    out = so.ddim_sample_loop(_model("wv3"), so.make_beta_schedule("cosine", 500), cond, 8, _noises(int(g["noise_seed"]), 26, (1, 8, 64, 64)), "ddim25")
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=1e-4)


def test_cave_ddim25(golden_dir):
    g = np.load(os.path.join(golden_dir, "cave_ddim25.npz"))
    cond = synth.make_batch("cave", 1, seed=int(g["data_seed"]))["cond"]
    out = so.ddim_sample_loop(_model("cave"), so.make_beta_schedule("cosine", 500), cond, 31, _noises(int(g["noise_seed"]), 26, (1, 31, 64, 64)), "ddim25")
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=1e-4)


def test_gf2_tiles_dpm25(golden_dir):
    g = np.load(os.path.join(golden_dir, "tiles_gf2_dpm25.npz"))
    cond = synth.make_batch("gf2", 1, size=128, seed=int(g["data_seed"]))["cond"]
    tiles = torch.cat([cond[:, :, y:y + 64, x:x + 64] for y in (0, 64) for x in (0, 64)], 0).contiguous()
    x_T = torch.randn(4, 4, 64, 64, generator=torch.Generator().manual_seed(int(g["noise_seed"])))
    ns = so.VPSchedule(so.schedule_buffers(so.make_beta_schedule("cosine", 500))["betas"])
    out = so.dpmpp_multistep_sample(_model("gf2"), ns, x_T.clone(), tiles, steps=25, order=2)  # the oracle batches what the reference runs per tile
    ref = g["out_tiles"]
    assert np.abs(out.numpy() - ref).max() <= 2e-3 * np.abs(ref).max()


def test_ddpm_T500_prefix_and_full(golden_dir):
    g = np.load(os.path.join(golden_dir, "full_ddpm_T500.npz"))
    T = int(g["T"])
    cond = synth.make_batch("wv3", 1, seed=int(g["data_seed"]))["cond"]
    noises = _noises(int(g["noise_seed"]), T + 1, (1, 8, 64, 64))
    sb = so.schedule_buffers(so.make_beta_schedule("cosine", T))
    full = os.environ.get("DDIF_SLOW_TESTS") == "1"
    model = _model("wv3")
    img, want = noises[0], {int(i): k for k, i in enumerate(g["inter_i"])}
    for k, i in enumerate(reversed(range(T))):
        t = torch.full((1,), i, dtype=torch.long)
        img = so.ddpm_step(sb, img, t, model(img, t, cond, img), cond[:, :8], noises[1 + k], (0.0, 1.0), "x_start")
        if i in want:
            np.testing.assert_allclose(img.numpy(), g["inter"][want[i]:want[i] + 1], rtol=0, atol=2e-4, err_msg=f"i={i}")
            if not full:
                return
    np.testing.assert_allclose(img.numpy(), g["out"], rtol=0, atol=2e-4)
