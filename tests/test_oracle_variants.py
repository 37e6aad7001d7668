"""Pin the oracle's solver variants (DPM-Solver / DPM-Solver++, multistep and singlestep), training objective and CC metric against
the fixtures the reference itself produced (tests/golden/make_golden2.py)."""
import os

import numpy as np
import pytest
import torch

from dif_pan_b200 import synth
from oracle import metrics_oracle, sampler_oracle as so

torch.set_grad_enabled(False)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# kept in sync with tests/golden/make_golden2.py::DPM_CASES (the fixture keys encode the case)
DPM_CASES = [
    ("singlestep", 1, 5, "time_uniform", "dpmsolver++", "x_start"),
    ("singlestep", 2, 6, "time_uniform", "dpmsolver++", "x_start"),
    ("singlestep", 2, 7, "time_uniform", "dpmsolver++", "noise"),
    ("singlestep", 3, 9, "time_uniform", "dpmsolver++", "x_start"),
    ("singlestep", 3, 10, "logSNR", "dpmsolver++", "x_start"),
    ("singlestep", 3, 11, "time_quadratic", "dpmsolver++", "v"),
    ("singlestep_fixed", 2, 8, "time_uniform", "dpmsolver++", "x_start"),
    ("singlestep_fixed", 3, 9, "time_uniform", "dpmsolver", "noise"),
    ("singlestep", 2, 6, "time_uniform", "dpmsolver", "x_start"),
    ("singlestep", 3, 8, "logSNR", "dpmsolver", "noise"),
    ("multistep", 2, 12, "time_uniform", "dpmsolver", "x_start"),
    ("multistep", 3, 12, "time_uniform", "dpmsolver", "noise"),
    ("multistep", 2, 6, "logSNR", "dpmsolver", "v"),
]


def case_key(c):
    return "_".join(str(v).replace("+", "p") for v in c)


def dpm_inputs(seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, 4, 16, 16, generator=g), torch.rand(1, 12, 16, 16, generator=g)


def rel_max(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("case", DPM_CASES, ids=[case_key(c) for c in DPM_CASES])
def test_dpm_variants_oracle_matches_reference(case):
    g = np.load(os.path.join(GOLDEN, "dpm_variants.npz"))
    x_T, cond = dpm_inputs(int(g["seed"]))
    method, order, steps, skip, algo, mtype = case
    ns = so.VPSchedule(torch.as_tensor(so.make_beta_schedule("cosine", 500), dtype=torch.float32))
    out = so.dpm_sample(so.analytic_denoiser, ns, x_T.clone(), cond, steps=steps, order=order, skip_type=skip, method=method,
                        algorithm=algo, model_type=mtype)
    # same fp32 expressions in the same order; trajectories that divide by alpha_T ~ 1e-4 amplify rounding, hence relative-to-max
    assert rel_max(out.numpy(), g[case_key(case)]) <= 2e-4, rel_max(out.numpy(), g[case_key(case)])


def loss_inputs(seed):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.rand(3, 8, 16, 16, generator=g) - 0.5
    nz = torch.randn(3, 8, 16, 16, generator=g)
    cond = torch.rand(3, 20, 16, 16, generator=g)
    return x0, nz, cond


def loss_model(x, t, cond=None, self_cond=None):
    y = so.analytic_denoiser(x, t.to(torch.float32), cond)
    return y if self_cond is None else y + 0.25 * self_cond


@pytest.mark.parametrize("pred_mode", ["x_start", "noise", "pred_v"])
@pytest.mark.parametrize("loss_type", ["l1", "l2"])
@pytest.mark.parametrize("sc", [0, 1])
def test_p_losses_oracle_matches_reference(pred_mode, loss_type, sc):
    g = np.load(os.path.join(GOLDEN, "losses.npz"))
    x0, nz, cond = loss_inputs(int(g["seed"]))
    sb = so.schedule_buffers(so.make_beta_schedule("cosine", 500), p2_gamma=0.5 if loss_type == "l2" else 0.0)
    loss, recon = so.p_losses(sb, loss_model, x0, torch.as_tensor(g["t"]), nz, cond, pred_mode, loss_type, bool(sc))
    assert abs(float(loss) - float(g[f"loss_{pred_mode}_{loss_type}_{sc}"])) <= 2e-5 * abs(float(g[f"loss_{pred_mode}_{loss_type}_{sc}"]))
    assert rel_max(recon.numpy(), g[f"recon_{pred_mode}_{loss_type}_{sc}"]) <= 1e-5


def test_cc_oracle_matches_reference():
    g = np.load(os.path.join(GOLDEN, "metrics_cc.npz"))
    d = synth.make_batch("wv3", 2, seed=int(g["seed"]))
    gen = torch.Generator().manual_seed(int(g["noise_seed"]))
    out = (d["hr"] + 0.03 * torch.randn(d["hr"].shape, generator=gen)).clamp(0, 1)
    for i in range(2):
        m = metrics_oracle.sam_ergas_psnr(d["hr"][i], out[i])
        np.testing.assert_allclose([m["SAM"], m["ERGAS"], m["PSNR"], m["CC"]], g["sam_ergas_psnr_cc"][i], rtol=1e-5)


ADAPTIVE_CASES = [(2, "dpmsolver++", "x_start"), (3, "dpmsolver++", "x_start"), (2, "dpmsolver", "x_start")]


@pytest.mark.parametrize("order,algo,mtype", ADAPTIVE_CASES)
def test_dpm_adaptive_oracle_matches_reference(order, algo, mtype):
    g = np.load(os.path.join(GOLDEN, "dpm_variants.npz"))
    x_T, cond = dpm_inputs(int(g["seed"]))
    ns = so.VPSchedule(torch.as_tensor(so.make_beta_schedule("cosine", 500), dtype=torch.float32))
    out, nfe = so.dpm_adaptive(so.analytic_denoiser, ns, x_T.clone(), cond, order=order, algorithm=algo, model_type=mtype)
    key = f"adaptive_{order}_{algo.replace('+', 'p')}_{mtype}"
    assert nfe > 0 and rel_max(out.numpy(), g[key]) <= 2e-4, (nfe, rel_max(out.numpy(), g[key]))


TAYLOR_CASES = [
    ("multistep", 2, 10, "time_uniform", "dpmsolver++", "x_start"),
    ("multistep", 2, 10, "time_uniform", "dpmsolver", "x_start"),
    ("singlestep", 2, 8, "time_uniform", "dpmsolver++", "x_start"),
    ("singlestep", 2, 7, "logSNR", "dpmsolver", "x_start"),
]


@pytest.mark.parametrize("case", TAYLOR_CASES, ids=[case_key(c) for c in TAYLOR_CASES])
def test_dpm_taylor_oracle_matches_reference(case):
    g = np.load(os.path.join(GOLDEN, "dpm_variants.npz"))
    x_T, cond = dpm_inputs(int(g["seed"]))
    method, order, steps, skip, algo, mtype = case
    ns = so.VPSchedule(torch.as_tensor(so.make_beta_schedule("cosine", 500), dtype=torch.float32))
    out = so.dpm_sample(so.analytic_denoiser, ns, x_T.clone(), cond, steps=steps, order=order, skip_type=skip, method=method,
                        algorithm=algo, model_type=mtype, solver_type="taylor")
    assert rel_max(out.numpy(), g["taylor_" + case_key(case)]) <= 2e-4


DTZ_CASES = [("multistep", 2, 10, "time_uniform", "dpmsolver++", "x_start"), ("singlestep", 3, 9, "time_uniform", "dpmsolver", "x_start")]


@pytest.mark.parametrize("case", DTZ_CASES, ids=[case_key(c) for c in DTZ_CASES])
def test_dpm_denoise_to_zero_oracle_matches_reference(case):
    g = np.load(os.path.join(GOLDEN, "dpm_variants.npz"))
    x_T, cond = dpm_inputs(int(g["seed"]))
    method, order, steps, skip, algo, mtype = case
    ns = so.VPSchedule(torch.as_tensor(so.make_beta_schedule("cosine", 500), dtype=torch.float32))
    out = so.dpm_sample(so.analytic_denoiser, ns, x_T.clone(), cond, steps=steps, order=order, skip_type=skip, method=method,
                        algorithm=algo, model_type=mtype, denoise_to_zero=True)
    assert rel_max(out.numpy(), g["dtz_" + case_key(case)]) <= 2e-4
