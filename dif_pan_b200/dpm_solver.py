"""NoiseScheduleVP / model_wrapper / DPM_Solver — drop-in for /root/reference/solver/dpm_solver.py ("dpm.py").

Same class / function signatures.  Schedule scalars (log-alpha interpolation, lambda, sigma, expm1 ...) are a few
floats per step and are computed on the HOST with fp32 torch-CPU arithmetic using the reference's formulas
(dpm.py:126-175, 555-588, 804-912); the per-element work of one solver step — the reference's
x_start -> noise -> x_start round trip (dpm.py:299-300, 446-447) plus the multistep update — is ONE fused CUDA
kernel (`ddif_dpmpp_step_f32`, csrc/sampler.cu).  With a `dif_pan_b200.UNetSR3` denoiser the loop runs in place
on the model's device buffers: per step one CUDA-graph launch + one fused kernel.

Implemented: algorithm_type 'dpmsolver++' and 'dpmsolver'; method 'multistep' (orders 1-3, lower_order_final),
'singlestep' (the "DPM-Solver-fast" order schedule, dpm.py:490-548) and 'singlestep_fixed' (orders 1-3, one fused
`ddif_dpm_single_f32` kernel per denoiser evaluation); skip types time_uniform / time_quadratic / logSNR; model types
x_start / noise / v; guidance 'uncond' or 'classifier-free' with scale 1 (the wiring SURVEY.md §3.3 names).
method='adaptive' (dpm.py:964-1018; one error-norm reduction kernel + a host scalar decision per iteration) is implemented
as well, and solver_type='taylor' for the multistep and the first/second-order singlestep updates (host scalars only).  The
third-order singlestep 'taylor' form and the thresholding correctors raise NotImplementedError (not used by any BASELINE config);
denoise_to_zero is one more evaluation + the data-prediction kernel.

Reference quirk kept out: model_wrapper multiplies `[B]`-shaped alpha_t against `[B,C,H,W]` (dpm.py:299-300),
which only broadcasts for B == 1 or B == W; all entries are equal, so a scalar multiply is the same arithmetic and
works for every batch size.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch

from . import _lib
from .unet import UNetSR3

MODEL_TYPES = {"x_start": 0, "noise": 1, "v": 2}


def interpolate_fn(x: torch.Tensor, xp: torch.Tensor, yp: torch.Tensor) -> torch.Tensor:
    """Piecewise-linear interpolation with linear extrapolation, x:[N,1], xp/yp:[1,K] (dpm.py:1261-1300),
    via searchsorted instead of the reference's sort+gather."""
    xs, xk, yk = x.reshape(-1), xp.reshape(-1), yp.reshape(-1)
    K = xk.shape[0]
    lo = torch.clamp(torch.searchsorted(xk, xs, right=False) - 1, 0, K - 2)
    x0, x1, y0, y1 = xk[lo], xk[lo + 1], yk[lo], yk[lo + 1]
    return (y0 + (xs - x0) * (y1 - y0) / (x1 - x0)).reshape(-1, 1)


class NoiseScheduleVP:
    def __init__(self, schedule="discrete", betas=None, alphas_cumprod=None, continuous_beta_0=0.1, continuous_beta_1=20.0,
                 dtype=torch.float32):
        if schedule not in ["discrete", "linear", "cosine"]:
            raise ValueError("Unsupported noise schedule {}. The schedule needs to be 'discrete' or 'linear' or 'cosine'".format(schedule))
        self.schedule = schedule
        if schedule == "discrete":
            if betas is not None:
                log_alphas = 0.5 * torch.log(1 - betas.detach().cpu()).cumsum(dim=0)
            else:
                assert alphas_cumprod is not None
                log_alphas = 0.5 * torch.log(alphas_cumprod.detach().cpu())
            self.total_N = len(log_alphas)
            self.T = 1.0
            self.t_array = torch.linspace(0.0, 1.0, self.total_N + 1)[1:].reshape((1, -1)).to(dtype=dtype)
            self.log_alpha_array = log_alphas.reshape((1, -1)).to(dtype=dtype)
        else:
            self.total_N = 1000
            self.beta_0, self.beta_1 = continuous_beta_0, continuous_beta_1
            self.cosine_s, self.cosine_beta_max = 0.008, 999.0
            self.cosine_t_max = (math.atan(self.cosine_beta_max * (1.0 + self.cosine_s) / math.pi) * 2.0 * (1.0 + self.cosine_s)
                                 / math.pi - self.cosine_s)
            self.cosine_log_alpha_0 = math.log(math.cos(self.cosine_s / (1.0 + self.cosine_s) * math.pi / 2.0))
            self.T = 0.9946 if schedule == "cosine" else 1.0

    def marginal_log_mean_coeff(self, t):
        t = torch.as_tensor(t, dtype=torch.float32).cpu()
        if self.schedule == "discrete":
            return interpolate_fn(t.reshape((-1, 1)), self.t_array, self.log_alpha_array).reshape((-1))
        if self.schedule == "linear":
            return -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.log(torch.cos((t + self.cosine_s) / (1.0 + self.cosine_s) * math.pi / 2.0)) - self.cosine_log_alpha_0

    def marginal_alpha(self, t):
        return torch.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        lm = self.marginal_log_mean_coeff(t)
        return lm - 0.5 * torch.log(1.0 - torch.exp(2.0 * lm))

    def inverse_lambda(self, lamb):
        lamb = torch.as_tensor(lamb, dtype=torch.float32).cpu()
        if self.schedule == "linear":
            tmp = 2.0 * (self.beta_1 - self.beta_0) * torch.logaddexp(-2.0 * lamb, torch.zeros((1,)))
            delta = self.beta_0 ** 2 + tmp
            return tmp / (torch.sqrt(delta) + self.beta_0) / (self.beta_1 - self.beta_0)
        if self.schedule == "discrete":
            log_alpha = -0.5 * torch.logaddexp(torch.zeros((1,)), -2.0 * lamb)
            t = interpolate_fn(log_alpha.reshape((-1, 1)), torch.flip(self.log_alpha_array, [1]), torch.flip(self.t_array, [1]))
            return t.reshape((-1,))
        log_alpha = -0.5 * torch.logaddexp(-2.0 * lamb, torch.zeros((1,)))
        return torch.arccos(torch.exp(log_alpha + self.cosine_log_alpha_0)) * 2.0 * (1.0 + self.cosine_s) / math.pi - self.cosine_s


class _WrappedModel:
    """What model_wrapper returns: callable like the reference's closure, plus the fields the fused path needs."""

    def __init__(self, model, noise_schedule, model_type, model_kwargs, guidance_type, condition, unconditional_condition,
                 guidance_scale, classifier_fn, classifier_kwargs):
        self.model, self.noise_schedule, self.model_type = model, noise_schedule, model_type
        self.model_kwargs, self.guidance_type, self.condition = model_kwargs, guidance_type, condition
        self.unconditional_condition, self.guidance_scale = unconditional_condition, guidance_scale
        if guidance_type == "classifier" or (guidance_type == "classifier-free" and guidance_scale != 1.0
                                             and unconditional_condition is not None):
            raise NotImplementedError("classifier guidance / classifier-free guidance with scale != 1 are not wired on the CUDA "
                                      "path (the DDIF wiring uses guidance_scale=1., SURVEY.md §3.3)")

    def input_time(self, t_continuous):
        if self.noise_schedule.schedule == "discrete":  # dpm.py:285-288
            return (t_continuous - 1.0 / self.noise_schedule.total_N) * 1000.0
        return t_continuous

    def raw(self, x, t_continuous):
        """The denoiser's raw output at continuous time t (dpm.py:290-295)."""
        t_in = self.input_time(t_continuous)
        if self.guidance_type == "uncond" or self.condition is None:
            return self.model(x, t_in, **self.model_kwargs)
        return self.model(x, t_in, self.condition, **self.model_kwargs)

    def __call__(self, x, t_continuous):
        """Noise prediction, like the reference closure (kept for API compatibility; the solver uses the fused kernel)."""
        out = self.raw(x, t_continuous)
        ns = self.noise_schedule
        t0 = torch.as_tensor(t_continuous).reshape(-1)[:1]
        a, s = float(ns.marginal_alpha(t0)), float(ns.marginal_std(t0))
        if self.model_type == "noise":
            return out
        if self.model_type == "x_start":
            return (x - a * out) / s
        if self.model_type == "v":
            return a * out + s * x
        return -s * out


def model_wrapper(model, noise_schedule, model_type="noise", model_kwargs={}, guidance_type="uncond", condition=None,
                  unconditional_condition=None, guidance_scale=1.0, classifier_fn=None, classifier_kwargs={}):
    assert model_type in ["noise", "x_start", "v", "score"]
    assert guidance_type in ["uncond", "classifier", "classifier-free"]
    return _WrappedModel(model, noise_schedule, model_type, model_kwargs, guidance_type, condition, unconditional_condition,
                         guidance_scale, classifier_fn, classifier_kwargs)


class DPM_Solver:
    def __init__(self, model_fn, noise_schedule, algorithm_type="dpmsolver++", correcting_x0_fn=None, correcting_xt_fn=None,
                 thresholding_max_val=1.0, dynamic_thresholding_ratio=0.995):
        assert algorithm_type in ["dpmsolver", "dpmsolver++"]
        if correcting_x0_fn is not None or correcting_xt_fn is not None:
            raise NotImplementedError("correcting_x0_fn / correcting_xt_fn (torch.quantile thresholding) are not on the CUDA path")
        if not isinstance(model_fn, _WrappedModel) or model_fn.model_type not in MODEL_TYPES:
            raise NotImplementedError("DPM_Solver needs the callable returned by dif_pan_b200.model_wrapper "
                                      "(model_type noise / x_start / v)")
        self.wrapped = model_fn
        self.model = lambda x, t: model_fn(x, t.expand((x.shape[0])))
        self.noise_schedule = noise_schedule
        self.algorithm_type = algorithm_type
        self._solver_type = "dpmsolver"

    def get_time_steps(self, skip_type, t_T, t_0, N, device=None):
        """dpm.py:461-488 (host tensors)."""
        if skip_type == "logSNR":
            lT = self.noise_schedule.marginal_lambda(torch.tensor(t_T))
            l0 = self.noise_schedule.marginal_lambda(torch.tensor(t_0))
            return self.noise_schedule.inverse_lambda(torch.linspace(lT.item(), l0.item(), N + 1))
        if skip_type == "time_uniform":
            return torch.linspace(t_T, t_0, N + 1)
        if skip_type == "time_quadratic":
            return torch.linspace(t_T ** 0.5, t_0 ** 0.5, N + 1).pow(2)
        raise ValueError("Unsupported skip_type {}, need to be 'logSNR' or 'time_uniform' or 'time_quadratic'".format(skip_type))

    def _coefficients(self, t_hist: List[torch.Tensor], t_next: torch.Tensor, order: int) -> dict:
        """fp32 scalars of one multistep update (dpm.py:569-599, 820-860, 876-912), both algorithm types.  For
        algorithm_type='dpmsolver' the scalars carry the signs of the ++ expression the kernel evaluates
        (x = cx*x - ca*m0 - cb*D1 at order 2, x = cx*x - ca*m0 + cb*D1 - cc*D2 at order 3)."""
        ns = self.noise_schedule
        lam = ns.marginal_lambda
        t0 = t_hist[-1]
        h = lam(t_next) - lam(t0)
        pp = self.algorithm_type == "dpmsolver++"
        if pp:
            alpha_t = torch.exp(ns.marginal_log_mean_coeff(t_next))
            phi_1 = torch.expm1(-h)
            scale = alpha_t
            c = dict(cx=ns.marginal_std(t_next) / ns.marginal_std(t0), ca=alpha_t * phi_1)
        else:
            scale = ns.marginal_std(t_next)
            phi_1 = torch.expm1(h)
            c = dict(cx=torch.exp(ns.marginal_log_mean_coeff(t_next) - ns.marginal_log_mean_coeff(t0)), ca=scale * phi_1)
        if order == 2:
            r0 = (lam(t0) - lam(t_hist[-2])) / h
            if self._solver_type == "taylor":   # dpm.py:843-848 / 855-860: x -= ... becomes  + a*(phi_1/h + 1)*D1  resp.  - s*(phi_1/h - 1)*D1
                cb = -(scale * (phi_1 / h + 1.0)) if pp else scale * (phi_1 / h - 1.0)
            else:
                cb = 0.5 * (scale * phi_1)
            c.update(cb=cb, inv_r0=1.0 / r0)
        elif order == 3:
            r0 = (lam(t0) - lam(t_hist[-2])) / h
            r1 = (lam(t_hist[-2]) - lam(t_hist[-3])) / h
            phi_2 = phi_1 / h + 1.0 if pp else phi_1 / h - 1.0
            phi_3 = phi_2 / h - 0.5
            cb = scale * phi_2
            c.update(cb=cb if pp else -cb, cc=scale * phi_3, inv_r0=1.0 / r0, inv_r1=1.0 / r1, k1=r0 / (r0 + r1), k2=1.0 / (r0 + r1))
        return {k: float(v) for k, v in c.items()}

    def get_orders_and_timesteps_for_singlestep_solver(self, steps, order, skip_type, t_T, t_0, device=None):
        """dpm.py:490-548 ("DPM-Solver-fast" order schedule)."""
        if order == 3:
            K = steps // 3 + 1
            if steps % 3 == 0:
                orders = [3] * (K - 2) + [2, 1]
            elif steps % 3 == 1:
                orders = [3] * (K - 1) + [1]
            else:
                orders = [3] * (K - 1) + [2]
        elif order == 2:
            if steps % 2 == 0:
                K = steps // 2
                orders = [2] * K
            else:
                K = steps // 2 + 1
                orders = [2] * (K - 1) + [1]
        elif order == 1:
            K = 1
            orders = [1] * steps
        else:
            raise ValueError("'order' must be '1' or '2' or '3'.")
        if skip_type == "logSNR":
            timesteps_outer = self.get_time_steps(skip_type, t_T, t_0, K)
        else:
            timesteps_outer = self.get_time_steps(skip_type, t_T, t_0, steps)[torch.cumsum(torch.tensor([0] + orders), 0)]
        return timesteps_outer, orders

    def _single_coefficients(self, s, t, order, r1, r2):
        """Stages of one singlestep update s -> t as (eval time, next eval time or None, kernel mode, c0, c1, c2), fp32 scalars in
        the reference's operation order (dpm.py:581-599 first, :623-674 second, :714-797 third; solver_type 'dpmsolver')."""
        ns = self.noise_schedule
        pp = self.algorithm_type == "dpmsolver++"
        lam_s, lam_t = ns.marginal_lambda(s), ns.marginal_lambda(t)
        h = lam_t - lam_s
        la, sd = ns.marginal_log_mean_coeff, ns.marginal_std
        em = (lambda v: torch.expm1(-v)) if pp else torch.expm1

        def lead(u):   # coefficient of x: sigma_u / sigma_s (++) or exp(log_alpha_u - log_alpha_s)
            return sd(u) / sd(s) if pp else torch.exp(la(u) - la(s))

        def amp(u):    # alpha_u (++) or sigma_u
            return torch.exp(la(u)) if pp else sd(u)

        f = lambda v: float(v)
        if order == 1:
            return [(s, None, 0, f(lead(t)), f(amp(t) * em(h)), 0.0)]
        if order == 2:
            r1 = 0.5 if r1 is None else r1
            s1 = ns.inverse_lambda(lam_s + r1 * h)
            phi_11, phi_1 = em(r1 * h), em(h)
            if self._solver_type == "taylor":   # dpm.py:651-656 / 674-679
                c2 = (1.0 / r1) * (amp(t) * (phi_1 / h + 1.0)) if pp else -((1.0 / r1) * (amp(t) * (phi_1 / h - 1.0)))
                c2 = f(c2)
            else:
                c2 = -f((0.5 / r1) * (amp(t) * phi_1))
            return [(s, s1, 0, f(lead(s1)), f(amp(s1) * phi_11), 0.0),
                    (s1, None, 1, f(lead(t)), f(amp(t) * phi_1), c2)]
        if self._solver_type == "taylor":
            raise NotImplementedError("solver_type='taylor' for the third-order singlestep update (three model values, dpm.py:759-768) is not on "
                                      "the CUDA path")
        r1 = 1.0 / 3.0 if r1 is None else r1
        r2 = 2.0 / 3.0 if r2 is None else r2
        s1, s2 = ns.inverse_lambda(lam_s + r1 * h), ns.inverse_lambda(lam_s + r2 * h)
        phi_11, phi_12, phi_1 = em(r1 * h), em(r2 * h), em(h)
        if pp:
            phi_22 = torch.expm1(-r2 * h) / (r2 * h) + 1.0
            phi_2 = phi_1 / h + 1.0
        else:
            phi_22 = torch.expm1(r2 * h) / (r2 * h) - 1.0
            phi_2 = phi_1 / h - 1.0
        c2_s2 = r2 / r1 * (amp(s2) * phi_22)
        c2_t = (1.0 / r2) * (amp(t) * phi_2)
        sgn = 1.0 if pp else -1.0
        return [(s, s1, 0, f(lead(s1)), f(amp(s1) * phi_11), 0.0),
                (s1, s2, 1, f(lead(s2)), f(amp(s2) * phi_12), sgn * f(c2_s2)),
                (s2, None, 1, f(lead(t)), f(amp(t) * phi_1), sgn * f(c2_t))]

    def dpm_solver_adaptive(self, x, order, t_T, t_0, h_init=0.05, atol=0.0078, rtol=0.05, theta=0.9, t_err=1e-5, solver_type="dpmsolver"):
        """Adaptive step-size solver (dpm.py:964-1018): a lower-order and a higher-order singlestep update share their denoiser
        evaluations; the per-sample error norm is ONE reduction kernel (`ddif_dpm_err_f32`) and the accept / step-size decision is a
        host scalar per iteration, as in the reference (`E.max()`, `torch.all(E <= 1.)` synchronise there too)."""
        if order not in (2, 3):
            raise ValueError("For adaptive step size solver, order must be 2 or 3, got {}".format(order))
        if solver_type != "dpmsolver":
            raise NotImplementedError("solver_type='taylor' is not on the CUDA path")
        if not x.is_cuda:
            raise RuntimeError("dif_pan_b200 sampler kernels run on CUDA only (no CPU fallback)")
        ns, wm = self.noise_schedule, self.wrapped
        B, dev = x.shape[0], x.device
        stream = _lib.stream_of(dev)
        fast = isinstance(wm.model, UNetSR3) and wm.condition is not None and not wm.model_kwargs
        predict = 0 if self.algorithm_type == "dpmsolver++" else 1
        mt = MODEL_TYPES[wm.model_type]
        s = t_T * torch.ones((1,))
        lambda_s = ns.marginal_lambda(s)
        lambda_0 = ns.marginal_lambda(t_0 * torch.ones_like(s))
        h = h_init * torch.ones_like(s)
        nfe = 0
        with torch.no_grad():
            if fast:
                rt = wm.model.runtime(B, x.shape[2], x.shape[3])
                rt.set_cond(wm.condition)
                xe, tbuf = rt.x_buf, rt.t_buf
            else:
                xe, tbuf = torch.empty_like(x, dtype=torch.float32).contiguous(), None
            xs = x.to(torch.float32).clone().contiguous()       # state at time s
            x_prev = xs.clone()
            x_lo, x_hi = torch.empty_like(xs), torch.empty_like(xs)
            m_s, m_k = torch.empty_like(xs), torch.empty_like(xs)
            err = torch.empty(B, dtype=torch.float64, device=dev)

            def evaluate(t_ev):
                if fast:
                    tbuf.fill_(float(wm.input_time(t_ev)))
                    rt.step()
                    return rt.out_buf
                return wm.raw(xe, t_ev.reshape(-1)[:1].to(dev).expand(B)).contiguous()

            def stage(out, t_ev, mode, c0, c1, c2, m_cur, x_out):
                _lib.launch("ddif_dpm_single_t", stream, x_base=xs.data_ptr(), x_eval=xe.data_ptr(), model_out=out.data_ptr(), m_cur=m_cur.data_ptr(),
                            m_a=m_s.data_ptr() if mode == 1 else None, x_out=x_out.data_ptr(), time_out=None, n=xs.numel(), batch=B, model_type=mt,
                            predict=predict, mode=mode, alpha_e=float(ns.marginal_alpha(t_ev)), sigma_e=float(ns.marginal_std(t_ev)), c0=c0, c1=c1,
                            c2=c2, t_next_in=0.0)

            while torch.abs((s - t_0)).mean() > t_err:
                t = ns.inverse_lambda(lambda_s + h)
                xe.copy_(xs)
                out = evaluate(s)
                if order == 2:
                    (_, _, _, a0, a1, _), = self._single_coefficients(s, t, 1, None, None)
                    hi = self._single_coefficients(s, t, 2, 0.5, None)
                    stage(out, s, 0, a0, a1, 0.0, m_s, x_lo)                         # x_lower = DPM-Solver-1, model_s
                    stage(out, s, 0, hi[0][3], hi[0][4], 0.0, m_s, xe)               # x_s1
                    out = evaluate(hi[1][0])
                    stage(out, hi[1][0], 1, hi[1][3], hi[1][4], hi[1][5], m_k, x_hi)  # x_higher = DPM-Solver-2
                else:
                    r1, r2 = 1.0 / 3.0, 2.0 / 3.0
                    lo = self._single_coefficients(s, t, 2, r1, None)
                    hi = self._single_coefficients(s, t, 3, r1, r2)
                    stage(out, s, 0, lo[0][3], lo[0][4], 0.0, m_s, xe)               # x_s1 (shared by both orders), model_s
                    out = evaluate(lo[1][0])
                    stage(out, lo[1][0], 1, lo[1][3], lo[1][4], lo[1][5], m_k, x_lo)  # x_lower = DPM-Solver-2 with r1 = 1/3
                    stage(out, hi[1][0], 1, hi[1][3], hi[1][4], hi[1][5], m_k, xe)    # x_s2 from model_s, model_s1
                    out = evaluate(hi[2][0])
                    stage(out, hi[2][0], 1, hi[2][3], hi[2][4], hi[2][5], m_k, x_hi)  # x_higher = DPM-Solver-3
                _lib.launch("ddif_dpm_err_t", stream, x_higher=x_hi.data_ptr(), x_lower=x_lo.data_ptr(), x_prev=x_prev.data_ptr(), out=err.data_ptr(),
                            batch=B, chw=xs[0].numel(), atol=float(atol), rtol=float(rtol))
                E = torch.sqrt(err / float(xs[0].numel())).max().to(torch.float32).cpu()
                if torch.all(E <= 1.0):
                    xs, x_hi = x_hi, xs
                    x_prev, x_lo = x_lo, x_prev
                    s = t
                    lambda_s = ns.marginal_lambda(s)
                h = torch.min(theta * h * torch.float_power(E, -1.0 / order).float(), lambda_0 - lambda_s)
                nfe += order
            self.last_nfe = nfe
            return xs.clone()

    def _denoise_to_zero(self, xb, t_0, rt=None):
        """denoise_to_zero_fn (dpm.py:550-554): one more denoiser evaluation at t_0 and x <- data_prediction_fn(x, t_0) (always the DATA
        prediction, whatever the algorithm type).  `rt`: the UNet runtime when xb is its input buffer (fast path)."""
        ns, wm = self.noise_schedule, self.wrapped
        B, dev = xb.shape[0], xb.device
        t = torch.ones((1,)) * t_0
        if rt is not None:
            rt.t_buf.fill_(float(wm.input_time(t)))
            rt.step()
            out = rt.out_buf
        else:
            out = wm.raw(xb, t.to(dev).expand(B)).contiguous()
        x0 = torch.empty_like(xb)
        _lib.launch("ddif_dpm_single_t", _lib.stream_of(dev), x_base=None, x_eval=xb.data_ptr(), model_out=out.data_ptr(),
                    m_cur=x0.data_ptr(), m_a=None, x_out=None, time_out=None, n=xb.numel(), batch=B, model_type=MODEL_TYPES[wm.model_type],
                    predict=0, mode=2, alpha_e=float(ns.marginal_alpha(t)), sigma_e=float(ns.marginal_std(t)), c0=0.0, c1=0.0, c2=0.0,
                    t_next_in=0.0)
        return x0

    def _sample_singlestep(self, x, steps, order, skip_type, method, t_T, t_0, return_intermediate, denoise_to_zero=False):
        """dpm.py:1222-1240: every outer step s -> t evaluates the denoiser `order` times (at s, s1[, s2]); each evaluation is
        followed by ONE fused kernel (model_wrapper round trip + prediction + the stage's linear update)."""
        ns, wm = self.noise_schedule, self.wrapped
        if method == "singlestep":
            outer, orders = self.get_orders_and_timesteps_for_singlestep_solver(steps, order, skip_type, t_T, t_0)
        else:
            K = steps // order
            orders = [order] * K
            outer = self.get_time_steps(skip_type, t_T, t_0, K)
        B, dev = x.shape[0], x.device
        stream = _lib.stream_of(dev)
        fast = isinstance(wm.model, UNetSR3) and wm.condition is not None and not wm.model_kwargs
        predict = 0 if self.algorithm_type == "dpmsolver++" else 1
        with torch.no_grad():
            if fast:
                rt = wm.model.runtime(B, x.shape[2], x.shape[3])
                rt.set_cond(wm.condition)
                xe, out, tbuf = rt.x_buf, rt.out_buf, rt.t_buf
                xe.copy_(x)
                tbuf.fill_(float(wm.input_time(outer[0])))
            else:
                xe, tbuf = x.clone().contiguous(), None
            xs = xe.clone()                 # state at the outer time s (the stages all start from it)
            m_s = torch.empty_like(xe)      # model_s
            m_k = torch.empty_like(xe)      # model_s1 / model_s2 (scratch)
            inter = []
            for step, o in enumerate(orders):
                s_t, t_t = outer[step], outer[step + 1]
                inner = self.get_time_steps(skip_type, s_t.item(), t_t.item(), o)
                lam_in = ns.marginal_lambda(inner)
                hh = lam_in[-1] - lam_in[0]
                r1 = None if o <= 1 else (lam_in[1] - lam_in[0]) / hh
                r2 = None if o <= 2 else (lam_in[2] - lam_in[0]) / hh
                stages = self._single_coefficients(s_t.reshape(1), t_t.reshape(1), o, r1, r2)
                for k, (t_ev, t_nx, mode, c0, c1, c2) in enumerate(stages):
                    if fast:
                        rt.step()
                    else:
                        out = wm.raw(xe, t_ev.reshape(-1)[:1].to(dev).expand(B)).contiguous()
                    last = k == len(stages) - 1
                    nxt_label = t_nx if not last else (outer[step + 1] if step + 1 < len(orders) else None)
                    _lib.launch(
                        "ddif_dpm_single_t", stream, x_base=xs.data_ptr(), x_eval=xe.data_ptr(), model_out=out.data_ptr(),
                        m_cur=(m_s if k == 0 else m_k).data_ptr(), m_a=m_s.data_ptr() if mode == 1 else None, x_out=xe.data_ptr(),
                        time_out=tbuf.data_ptr() if (fast and nxt_label is not None) else None, n=xe.numel(), batch=B,
                        model_type=MODEL_TYPES[wm.model_type], predict=predict, mode=mode, alpha_e=float(ns.marginal_alpha(t_ev)),
                        sigma_e=float(ns.marginal_std(t_ev)), c0=c0, c1=c1, c2=c2,
                        t_next_in=float(wm.input_time(nxt_label)) if nxt_label is not None else 0.0)
                xs.copy_(xe)
                if return_intermediate:
                    inter.append(xe.clone())
            res = xe.clone()
            if denoise_to_zero:
                res = self._denoise_to_zero(xe, t_0, rt if fast else None)
                if return_intermediate:
                    inter.append(res.clone())
        return (res, inter) if return_intermediate else res

    def sample(self, x, steps=20, t_start=None, t_end=None, order=2, skip_type="time_uniform", method="multistep",
               lower_order_final=True, denoise_to_zero=False, solver_type="dpmsolver", atol=0.0078, rtol=0.05,
               return_intermediate=False):
        """dpm.py:1055-1253, multistep branch (:1179-1221)."""
        ns = self.noise_schedule
        t_0 = 1.0 / ns.total_N if t_end is None else t_end
        t_T = ns.T if t_start is None else t_start
        assert t_0 > 0 and t_T > 0, "Time range needs to be greater than 0. For discrete-time DPMs, it needs to be in [1 / N, 1], where N is the length of betas array"
        if method not in ("multistep", "singlestep", "singlestep_fixed", "adaptive"):
            raise ValueError("Got wrong method {}".format(method))
        if return_intermediate:
            assert method in ["multistep", "singlestep", "singlestep_fixed"], "Cannot use adaptive solver when saving intermediate values"
        if solver_type not in ("dpmsolver", "taylor"):
            raise ValueError("'solver_type' must be either 'dpmsolver' or 'taylor', got {}".format(solver_type))
        self._solver_type = solver_type  # 'taylor' only changes host-side scalars (orders <= 2; the multistep third order has one form)
        if order not in (1, 2, 3):
            raise ValueError("Solver order must be 1 or 2 or 3, got {}".format(order))
        if not x.is_cuda:
            raise RuntimeError("dif_pan_b200 sampler kernels run on CUDA only (no CPU fallback)")
        if method == "adaptive":
            res = self.dpm_solver_adaptive(x, order=order, t_T=t_T, t_0=t_0, atol=atol, rtol=rtol, solver_type=solver_type)
            return self._denoise_to_zero(res, t_0) if denoise_to_zero else res
        if method != "multistep":
            return self._sample_singlestep(x, steps, order, skip_type, method, t_T, t_0, return_intermediate, denoise_to_zero)
        assert steps >= order
        wm = self.wrapped
        ts = self.get_time_steps(skip_type, t_T, t_0, steps)
        assert ts.shape[0] - 1 == steps
        B = x.shape[0]
        dev = x.device
        stream = _lib.stream_of(dev)
        fast = isinstance(wm.model, UNetSR3) and wm.condition is not None and not wm.model_kwargs
        with torch.no_grad():
            if fast:
                rt = wm.model.runtime(B, x.shape[2], x.shape[3])
                rt.set_cond(wm.condition)
                xb, out, tbuf = rt.x_buf, rt.out_buf, rt.t_buf
                xb.copy_(x)
                tbuf.fill_(float(wm.input_time(ts[0])))
            else:
                xb = x.clone().contiguous()
                tbuf = None
            m = [torch.empty_like(xb) for _ in range(3)]
            inter = [xb.clone()] if return_intermediate else []
            for s in range(steps):
                t_cur = ts[s]
                if fast:
                    rt.step()
                else:
                    out = wm.raw(xb, t_cur.to(dev).expand(B)).contiguous()
                nxt = s + 1
                if nxt < order:
                    o = nxt
                elif lower_order_final and steps < 10:
                    o = min(order, steps + 1 - nxt)
                else:
                    o = order
                coef = self._coefficients([ts[i] for i in range(max(0, s - 2), s + 1)], ts[nxt], o)
                last_eval = nxt == steps
                _lib.launch(
                    "ddif_dpmpp_step_t", stream, x=xb.data_ptr(), model_out=out.data_ptr(), m_cur=m[s % 3].data_ptr(),
                    m_prev1=m[(s - 1) % 3].data_ptr(), m_prev2=m[(s - 2) % 3].data_ptr(),
                    time_out=tbuf.data_ptr() if (fast and not last_eval) else None, n=xb.numel(), batch=B, order=o,
                    model_type=MODEL_TYPES[wm.model_type], alpha_t=float(ns.marginal_alpha(t_cur)), sigma_t=float(ns.marginal_std(t_cur)),
                    t_next_in=float(wm.input_time(ts[nxt])), predict=0 if self.algorithm_type == "dpmsolver++" else 1, **coef)
                if return_intermediate:
                    inter.append(xb.clone())
            res = xb.clone()
            if denoise_to_zero:
                res = self._denoise_to_zero(xb, t_0, rt if fast else None)
                if return_intermediate:
                    inter.append(res.clone())
        return (res, inter) if return_intermediate else res
