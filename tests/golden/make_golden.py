#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Imports models/sr3_dwt.py, diffusion/diffusion_ddpm_pan.py, solver/dpm_solver.py and
utils/_metric_legacy.py from /root/reference (with the `timm.DropPath` stub in _stubs/), feeds them
seeded synthetic inputs from dif_pan_b200.synth and stores inputs-by-seed + reference outputs as
small .npz files.  Randomness inside the reference loops (`torch.randn`) is replaced by a queue of
pre-generated tensors so every implementation can be driven with identical noise.
"""
from __future__ import annotations

import contextlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DDIF_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(1, REF)
sys.path.insert(2, ROOT)

from models.sr3_dwt import UNetSR3 as RefUNet  # noqa: E402
from diffusion.diffusion_ddpm_pan import GaussianDiffusion as RefDiffusion, make_beta_schedule as ref_betas  # noqa: E402
from solver.dpm_solver import NoiseScheduleVP, model_wrapper, DPM_Solver  # noqa: E402
from utils._metric_legacy import analysis_accu  # noqa: E402

from dif_pan_b200 import synth  # noqa: E402

torch.set_grad_enabled(False)
torch.set_num_threads(max(1, os.cpu_count() or 1))


@contextlib.contextmanager
def queued_randn(queue):
    """Replace torch.randn by a FIFO of pre-generated tensors (shape-checked)."""
    real = torch.randn

    def fake(*size, **kw):
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        t = queue.pop(0)
        assert tuple(t.shape) == shape, (t.shape, shape)
        return t.clone()

    torch.randn = fake
    try:
        yield
    finally:
        torch.randn = real


def build_ref_unet(dataset, seed=0):
    kw = synth.unet_kwargs(dataset)
    net = RefUNet(**kw).eval()
    sd = synth.make_state_dict(seed, **kw)
    missing = net.load_state_dict(sd, strict=True)
    assert len(net.state_dict()) == len(sd) == 702, len(sd)
    return net, kw


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrs.items()})
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.0f} KiB")


def gen_unet(dataset, batch, seed):
    net, kw = build_ref_unet(dataset)
    data = synth.make_batch(dataset, batch, seed=seed)
    g = torch.Generator().manual_seed(seed + 7)
    C = kw["in_channel"]
    x = torch.randn(batch, C, 64, 64, generator=g)
    sc = torch.randn(batch, C, 64, 64, generator=g) * 0.3
    t_long = torch.tensor([417, 3][:batch], dtype=torch.long)
    t_float = torch.tensor([948.1, 49.9][:batch], dtype=torch.float32)
    taps = {}
    hooks = []
    for grp in ("downs", "mid", "ups"):
        for i, m in enumerate(getattr(net, grp)):
            hooks.append(m.register_forward_hook(
                lambda _m, _i, o, key=f"{grp}.{i}": taps.__setitem__(key, torch.stack([o.mean(), o.std(), o.abs().max()]))))
    y_long = net(x, t_long, data["cond"])
    tap_keys = sorted(taps)
    tap_stats = torch.stack([taps[k] for k in tap_keys])
    for h in hooks:
        h.remove()
    y_float = net(x, t_float, data["cond"])
    y_sc = net(x, t_long, data["cond"], sc)
    n_params = sum(p.numel() for p in net.parameters())
    save(f"unet_{dataset}.npz", seed=seed, batch=batch, gen_seed=seed + 7, t_long=t_long, t_float=t_float,
         y_long=y_long, y_float=y_float, y_selfcond=y_sc, tap_keys=np.array(tap_keys), tap_stats=tap_stats,
         n_params=n_params, n_entries=len(net.state_dict()))


def gen_schedule():
    betas = ref_betas("cosine", 500, cosine_s=8e-3)
    net, kw = build_ref_unet("wv3")
    dif = RefDiffusion(net, image_size=64, channels=8, pred_mode="x_start", loss_type="l1", device="cpu", clamp_range=(0, 1))
    dif.set_new_noise_schedule(betas=betas)
    bufs = {k: v.clone() for k, v in dif.named_buffers()}
    use = dif.space_timesteps(dif.num_timesteps, "ddim25")
    dif.space_new_betas(use)
    bufs25 = {"ddim25_" + k: v.clone() for k, v in dif.named_buffers()}
    lin = ref_betas("linear", 200)
    save("schedule.npz", betas64=betas, linear200=lin, ddim25_use=np.array(sorted(use)), **bufs, **bufs25)


def gen_loops():
    net, kw = build_ref_unet("wv3")
    data = synth.make_batch("wv3", 1, seed=1236)
    cond = data["cond"]
    g = torch.Generator().manual_seed(99)

    # --- full DDPM loop, short schedule (T=6, cosine), clamp (0,1), x_start ---------------------------------
    T = 6
    dif = RefDiffusion(net, image_size=64, channels=8, pred_mode="x_start", loss_type="l1", device="cpu", clamp_range=(0, 1))
    dif.set_new_noise_schedule(betas=ref_betas("cosine", T, cosine_s=8e-3))
    noises = [torch.randn(1, 8, 64, 64, generator=g) for _ in range(T + 1)]
    with queued_randn([n.clone() for n in noises]):
        out = dif(cond, mode="ddpm_sample")
    save("ddpm_T6.npz", T=T, data_seed=1236, noise_seed=99, out=out, noise0=noises[0])

    # --- DDIM: T=100 cosine respaced to ddim5, eta 0 --------------------------------------------------------
    dif = RefDiffusion(net, image_size=64, channels=8, pred_mode="x_start", loss_type="l1", device="cpu", clamp_range=(0, 1))
    dif.set_new_noise_schedule(betas=ref_betas("cosine", 100, cosine_s=8e-3))
    g = torch.Generator().manual_seed(100)
    noises = [torch.randn(1, 8, 64, 64, generator=g) for _ in range(6)]
    with queued_randn([n.clone() for n in noises]):
        out = dif(cond, mode="ddim_sample", section_counts="ddim5")
    save("ddim_T100_5.npz", T=100, data_seed=1236, noise_seed=100, out=out)

    # --- single ddpm / ddim / q_sample steps on random tensors (no UNet) with pred_mode variants -----------
    g = torch.Generator().manual_seed(101)
    dif.set_new_noise_schedule(betas=ref_betas("cosine", 500, cosine_s=8e-3))
    x = torch.randn(3, 8, 16, 16, generator=g)
    mo = torch.randn(3, 8, 16, 16, generator=g) * 0.2
    c = torch.rand(3, 20, 16, 16, generator=g)
    nz = torch.randn(3, 8, 16, 16, generator=g)
    t = torch.tensor([499, 250, 0])
    steps = {}
    for pm in ("x_start", "noise", "pred_v"):
        dif.pred_mode = pm
        mean, _, logvar, x0 = dif.p_mean_variance(x, t, True, condition_x=c, model_out=mo.clone())
        nzm = (1 - (t == 0).float()).reshape(3, 1, 1, 1)
        steps["ddpm_" + pm] = mean + nzm * (0.5 * logvar).exp() * nz
    dif.pred_mode = "x_start"
    steps["q_sample"] = dif.q_sample(x, t, nz)
    # ddim step with eta 0 and 0.5: reproduce ddim_sample's arithmetic through the reference's helpers
    for eta in (0.0, 0.5):
        class _M:  # tiny denoiser returning the fixed model_out
            self_condition = True
            pred_var = False

            def forward(self, *a, **k):
                return mo.clone()
        d2 = RefDiffusion(_M(), image_size=16, channels=8, pred_mode="x_start", loss_type="l1", device="cpu", clamp_range=(0, 1))
        d2.set_new_noise_schedule(betas=ref_betas("cosine", 500, cosine_s=8e-3))
        with queued_randn([nz.clone()]):
            steps[f"ddim_eta{eta}"] = d2.ddim_sample(x, t, condition_x=c, eta=eta)
    save("steps.npz", seed=101, **steps)

    # --- DPM-Solver++ multistep, T=500 cosine, 20 steps order 2 and 12 steps order 3, B=1 -------------------
    dif.set_new_noise_schedule(betas=ref_betas("cosine", 500, cosine_s=8e-3))
    ns = NoiseScheduleVP("discrete", betas=dif.betas)
    t_seen = []

    def unet(x, t, c):
        t_seen.append(t.clone())
        return net(x, t, c)

    mfn = model_wrapper(unet, ns, model_type="x_start", guidance_type="classifier-free", condition=cond, guidance_scale=1.0)
    sol = DPM_Solver(mfn, ns, algorithm_type="dpmsolver++")
    g = torch.Generator().manual_seed(102)
    x_T = torch.randn(1, 8, 64, 64, generator=g)
    out2, inter2 = sol.sample(x_T.clone(), steps=20, order=2, skip_type="time_uniform", method="multistep", return_intermediate=True)
    t_in20 = torch.stack(t_seen).reshape(-1)
    t_seen.clear()
    out3 = sol.sample(x_T.clone(), steps=12, order=3, skip_type="time_uniform", method="multistep")
    t_seen.clear()
    out1 = sol.sample(x_T.clone(), steps=5, order=2, skip_type="time_uniform", method="multistep")  # lower_order_final path
    # scalar tables of the noise schedule at the 21 time points
    ts = torch.linspace(1.0, 1.0 / 500, 21)
    save("dpm.npz", data_seed=1236, noise_seed=102, out_o2_s20=out2, x_after_step1=inter2[1], x_after_step10=inter2[10],
         t_in20=t_in20, out_o3_s12=out3, out_o2_s5=out1, ts=ts, log_alpha=ns.marginal_log_mean_coeff(ts),
         std=ns.marginal_std(ts), lam=ns.marginal_lambda(ts))


def gen_metrics():
    d = synth.make_batch("wv3", 2, seed=1240)
    g = torch.Generator().manual_seed(5)
    out = (d["hr"] + 0.03 * torch.randn(d["hr"].shape, generator=g)).clamp(0, 1)
    rows = []
    for a, b in zip(d["hr"], out):
        m = analysis_accu(a.permute(1, 2, 0), b.permute(1, 2, 0), 4, choices=4)
        rows.append([float(m["SAM"]), float(m["ERGAS"]), float(m["PSNR"])])
    save("metrics.npz", seed=1240, noise_seed=5, sam_ergas_psnr=np.array(rows))


if __name__ == "__main__":
    gen_schedule()
    gen_metrics()
    gen_unet("wv3", 2, 1235)
    gen_unet("gf2", 1, 1237)
    gen_unet("cave", 1, 1238)
    gen_loops()
