"""Pin the CPU oracle against the golden vectors produced by the reference itself
(tests/golden/make_golden.py) and against the PyWavelets documentation known answers."""
import os

import numpy as np
import pytest
import torch

from dif_pan_b200 import synth
from oracle import metrics_oracle, sampler_oracle as so, unet_oracle as uo, wavelet_oracle as wo

torch.set_grad_enabled(False)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _cfg(dataset):
    kw = synth.unet_kwargs(dataset)
    kw.pop("dropout")
    return uo.UNetCfg(**kw)


@pytest.mark.parametrize("dataset", ["wv3", "gf2", "cave"])
def test_unet_oracle_matches_reference(golden_dir, dataset):
    g = _load(golden_dir, f"unet_{dataset}.npz")
    B, seed = int(g["batch"]), int(g["seed"])
    kw = synth.unet_kwargs(dataset)
    sd = synth.make_state_dict(0, **kw)
    assert len(sd) == int(g["n_entries"]) == 702
    assert sum(v.numel() for v in sd.values()) == int(g["n_params"])
    data = synth.make_batch(dataset, B, seed=seed)
    gen = torch.Generator().manual_seed(int(g["gen_seed"]))
    C = kw["in_channel"]
    x = torch.randn(B, C, 64, 64, generator=gen)
    sc = torch.randn(B, C, 64, 64, generator=gen) * 0.3
    cfg = _cfg(dataset)
    taps = {}
    y = uo.unet_forward(sd, cfg, x, torch.tensor(g["t_long"]), data["cond"], taps=taps)
    np.testing.assert_allclose(y.numpy(), g["y_long"], rtol=0, atol=2e-5)
    for key, ref in zip(g["tap_keys"], g["tap_stats"]):
        t = taps[str(key)]
        got = np.array([t.mean(), t.std(), t.abs().max()])
        np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-5, err_msg=str(key))
    y = uo.unet_forward(sd, cfg, x, torch.tensor(g["t_float"]), data["cond"])
    np.testing.assert_allclose(y.numpy(), g["y_float"], rtol=0, atol=2e-5)
    y = uo.unet_forward(sd, cfg, x, torch.tensor(g["t_long"]), data["cond"], sc)
    np.testing.assert_allclose(y.numpy(), g["y_selfcond"], rtol=0, atol=2e-5)


def test_param_and_flop_anchors():
    # SURVEY.md §8(c) anchors: 10 397 208 params (WV3); 8.378 GFLOP per 64x64 patch-forward.
    sd = synth.make_state_dict(0, **synth.unet_kwargs("wv3"))
    assert sum(v.numel() for v in sd.values()) == 10_397_208
    assert abs(uo.count_flops(_cfg("wv3")) / 1e9 - 8.378) < 0.01
    assert abs(uo.count_flops(_cfg("gf2")) / 1e9 - 8.127) < 0.01
    assert abs(uo.count_flops(_cfg("cave")) / 1e9 - 9.964) < 0.01


def test_schedule_buffers(golden_dir):
    g = _load(golden_dir, "schedule.npz")
    betas = so.make_beta_schedule("cosine", 500)
    np.testing.assert_array_equal(betas, g["betas64"])
    assert abs(betas[0] - 8.7424e-05) < 1e-8 and betas[-1] == 0.999
    np.testing.assert_array_equal(so.make_beta_schedule("linear", 200), g["linear200"])
    sb = so.schedule_buffers(betas)
    for k in so.SCHEDULE_BUFFERS:
        np.testing.assert_array_equal(sb[k].numpy(), g[k], err_msg=k)
    use = so.space_timesteps(500, "ddim25")
    assert sorted(use) == list(g["ddim25_use"])
    sb25 = so.schedule_buffers(so.spaced_betas(sb["alphas_cumprod"], use))
    for k in so.SCHEDULE_BUFFERS:
        np.testing.assert_array_equal(sb25[k].numpy(), g["ddim25_" + k], err_msg=k)


def test_space_timesteps_sections():
    assert so.space_timesteps(300, "10,15,20") == so.space_timesteps(300, [10, 15, 20])
    assert len(so.space_timesteps(300, "10,15,20")) == 45
    with pytest.raises(ValueError):
        so.space_timesteps(500, "ddim300")


def test_single_steps(golden_dir):
    g = _load(golden_dir, "steps.npz")
    gen = torch.Generator().manual_seed(int(g["seed"]))
    x = torch.randn(3, 8, 16, 16, generator=gen)
    mo = torch.randn(3, 8, 16, 16, generator=gen) * 0.2
    c = torch.rand(3, 20, 16, 16, generator=gen)
    nz = torch.randn(3, 8, 16, 16, generator=gen)
    t = torch.tensor([499, 250, 0])
    sb = so.schedule_buffers(so.make_beta_schedule("cosine", 500))
    for pm in ("x_start", "noise", "pred_v"):
        got = so.ddpm_step(sb, x, t, mo, c[:, :8], nz, (0.0, 1.0), pm)
        np.testing.assert_allclose(got.numpy(), g["ddpm_" + pm], rtol=1e-6, atol=1e-6, err_msg=pm)
    np.testing.assert_array_equal(so.q_sample(sb, x, t, nz).numpy(), g["q_sample"])
    for eta in (0.0, 0.5):
        got = so.ddim_step(sb, x, t, mo, nz, eta)
        np.testing.assert_allclose(got.numpy(), g[f"ddim_eta{eta}"], rtol=1e-6, atol=1e-6)


def _wv3_model():
    kw = synth.unet_kwargs("wv3")
    sd = synth.make_state_dict(0, **kw)
    cfg = _cfg("wv3")
    return lambda x, t, c, sc: uo.unet_forward(sd, cfg, x, t, c, sc)


def test_ddpm_and_ddim_loops(golden_dir):
    model = _wv3_model()
    g = _load(golden_dir, "ddpm_T6.npz")
    cond = synth.make_batch("wv3", 1, seed=int(g["data_seed"]))["cond"]
    gen = torch.Generator().manual_seed(int(g["noise_seed"]))
    T = int(g["T"])
    noises = [torch.randn(1, 8, 64, 64, generator=gen) for _ in range(T + 1)]
    sb = so.schedule_buffers(so.make_beta_schedule("cosine", T))
    out = so.ddpm_sample_loop(model, sb, cond, 8, noises)
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=5e-5)

    g = _load(golden_dir, "ddim_T100_5.npz")
    gen = torch.Generator().manual_seed(int(g["noise_seed"]))
    noises = [torch.randn(1, 8, 64, 64, generator=gen) for _ in range(6)]
    out = so.ddim_sample_loop(model, so.make_beta_schedule("cosine", 100), cond, 8, noises, "ddim5")
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=5e-5)


def test_dpm_solver(golden_dir):
    g = _load(golden_dir, "dpm.npz")
    sb = so.schedule_buffers(so.make_beta_schedule("cosine", 500))
    ns = so.VPSchedule(sb["betas"])
    ts = torch.tensor(g["ts"])
    np.testing.assert_allclose(ns.log_alpha(ts).numpy(), g["log_alpha"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(ns.std(ts).numpy(), g["std"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(ns.lam(ts).numpy(), g["lam"], rtol=1e-5, atol=1e-6)
    model = _wv3_model()
    cond = synth.make_batch("wv3", 1, seed=int(g["data_seed"]))["cond"]
    gen = torch.Generator().manual_seed(int(g["noise_seed"]))
    x_T = torch.randn(1, 8, 64, 64, generator=gen)
    t_seen = []

    def m(x, t, c, sc):
        t_seen.append(t.clone())
        return model(x, t, c, sc)

    trace = []
    out = so.dpmpp_multistep_sample(m, ns, x_T.clone(), cond, steps=20, order=2, trace=trace)
    np.testing.assert_allclose(torch.stack(t_seen).reshape(-1).numpy(), g["t_in20"], rtol=1e-6)
    assert len(t_seen) == 20  # NFE
    # the x_start -> noise -> x_start round trip divides by alpha_T ~ 1e-4: compare relative to scale
    ref = g["out_o2_s20"]
    assert np.abs(out.numpy() - ref).max() <= 2e-3 * np.abs(ref).max()
    np.testing.assert_allclose(trace[1][1].numpy(), g["x_after_step1"], rtol=1e-4, atol=1e-4)
    out3 = so.dpmpp_multistep_sample(model, ns, x_T.clone(), cond, steps=12, order=3)
    ref = g["out_o3_s12"]
    assert np.abs(out3.numpy() - ref).max() <= 2e-3 * np.abs(ref).max()
    out5 = so.dpmpp_multistep_sample(model, ns, x_T.clone(), cond, steps=5, order=2)
    ref = g["out_o2_s5"]
    assert np.abs(out5.numpy() - ref).max() <= 2e-3 * np.abs(ref).max()


def test_haar_known_answers():
    ca, cd = wo.haar_dwt1(np.array([1.0, 2.0, 3.0, 4.0]))
    np.testing.assert_allclose(ca, [2.12132034, 4.94974747], atol=1e-8)
    np.testing.assert_allclose(cd, [-0.70710678, -0.70710678], atol=1e-8)
    cA, (cH, cV, cD) = wo.haar_dwt2(np.array([[1.0, 2.0], [3.0, 4.0]]))
    assert (cA.item(), cH.item(), cV.item(), cD.item()) == (5.0, -2.0, -1.0, 0.0)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, 3, 8, 6))
    cA, (cH, cV, cD) = wo.haar_dwt2(x)
    np.testing.assert_allclose(wo.haar_idwt2(cA, cH, cV, cD), x, atol=1e-12)
    # separable: 2-D == 1-D along rows then columns
    lo, hi = wo.haar_dwt1(x)
    ll, lh = wo.haar_dwt1(np.swapaxes(lo, -1, -2))
    np.testing.assert_allclose(np.swapaxes(ll, -1, -2), cA, atol=1e-12)


def test_cond_assembly_matches_synth():
    d = synth.make_batch("wv3", 2, seed=3)
    w = wo.wavelet_channels(d["lms_dn"].numpy(), d["pan_dn"].numpy(), 2047.0, "pan")
    np.testing.assert_allclose(w.numpy(), d["wavelets"].numpy(), atol=1e-7)
    c = wo.assemble_cond(d["lms"], d["pan"], w)
    assert c.shape == (2, 20, 64, 64)
    np.testing.assert_allclose(c.numpy(), d["cond"].numpy(), atol=1e-7)
    d = synth.make_batch("cave", 1, seed=3)
    assert d["cond"].shape == (1, 74, 64, 64)


def test_metrics(golden_dir):
    g = _load(golden_dir, "metrics.npz")
    d = synth.make_batch("wv3", 2, seed=int(g["seed"]))
    gen = torch.Generator().manual_seed(int(g["noise_seed"]))
    out = (d["hr"] + 0.03 * torch.randn(d["hr"].shape, generator=gen)).clamp(0, 1)
    for i in range(2):
        m = metrics_oracle.sam_ergas_psnr(d["hr"][i], out[i])
        np.testing.assert_allclose([m["SAM"], m["ERGAS"], m["PSNR"]], g["sam_ergas_psnr"][i], rtol=1e-5)
