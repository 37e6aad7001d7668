"""GPU parity of the memory-bound kernels (csrc/elementwise.cu, csrc/sampler.cu) against torch / the CPU oracle."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gpu_util import DEV, nhwc_bf16, pack_w, rel_err, stream, to_nchw_f32  # noqa: E402
import dif_pan_b200 as dp  # noqa: E402
from dif_pan_b200 import _lib, synth  # noqa: E402
from oracle import sampler_oracle as so, wavelet_oracle as wo  # noqa: E402


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def _stats_of(x_nhwc):
    f = x_nhwc.double()
    return torch.stack([f.sum(dim=(1, 2, 3)), (f * f).sum(dim=(1, 2, 3))], dim=1).contiguous()


# ---------------------------------------------------------------------------------------------------------------
def test_in_convert_and_upsample_and_stats():
    B, C, H, W = 2, 8, 16, 24
    x, sc = _rand(B, C, H, W, seed=1), _rand(B, C, H, W, seed=2)
    out = torch.zeros(B, H, W, 16, dtype=torch.bfloat16, device=DEV)
    _lib.launch("ddif_in_convert_t", stream(), x=x.data_ptr(), self_cond=sc.data_ptr(), out=out.data_ptr(), batch=B, c=C, h=H, w=W, c_pad=16)
    ref = torch.cat([sc, x], 1).to(torch.bfloat16)
    assert torch.equal(to_nchw_f32(out), ref.float())
    out5 = torch.full((B, H, W, 8), 7.0, dtype=torch.bfloat16, device=DEV)
    x5 = x[:, :5].contiguous()
    _lib.launch("ddif_in_convert_t", stream(), x=x5.data_ptr(), self_cond=None, out=out5.data_ptr(), batch=B, c=5, h=H, w=W, c_pad=8)
    assert torch.equal(to_nchw_f32(out5)[:, :5], x[:, :5].to(torch.bfloat16).float()) and float(out5[..., 5:].abs().max()) == 0.0
    a = nhwc_bf16(_rand(B, 32, H, W, seed=3))
    up = torch.zeros(B, 2 * H, 2 * W, 32, dtype=torch.bfloat16, device=DEV)
    _lib.launch("ddif_upsample2x_t", stream(), **{"in": a.data_ptr()}, out=up.data_ptr(), batch=B, h=H, w=W, c=32)
    assert torch.equal(to_nchw_f32(up), F.interpolate(to_nchw_f32(a), scale_factor=2, mode="nearest"))
    st = torch.zeros(B, 2, dtype=torch.float64, device=DEV)
    _lib.launch("ddif_stats_t", stream(), **{"in": a.data_ptr()}, stats=st.data_ptr(), batch=B, hw=H * W, c=32)
    assert torch.allclose(st, _stats_of(a), rtol=1e-5, atol=1e-2)


@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("two", [False, True])
def test_gn_apply(act, two):
    B, H, W, c1, c2 = 3, 16, 16, 96, (32 if two else 0)
    a = nhwc_bf16(_rand(B, c1, H, W, seed=4, scale=2.0) + 0.7)
    b = nhwc_bf16(_rand(B, 32, H, W, seed=5) - 0.3) if two else None
    C = c1 + c2
    gamma, beta = 1 + 0.1 * _rand(C, seed=6), 0.1 * _rand(C, seed=7)
    dw = _rand(C, 1, 3, 3, seed=8, scale=0.3)
    dw9 = dw.reshape(C, 9).t().contiguous()
    out = torch.zeros(B, H, W, C, dtype=torch.bfloat16, device=DEV)
    out_dw = torch.zeros_like(out)
    sa, sb = _stats_of(a), (_stats_of(b) if two else None)
    _lib.launch("ddif_gn_apply_t", stream(), src1=a.data_ptr(), c1=c1, src2=b.data_ptr() if two else None, c2=c2, stats1=sa.data_ptr(),
                stats2=sb.data_ptr() if two else None, gamma=gamma.data_ptr(), beta=beta.data_ptr(), out=out.data_ptr(),
                dw_w=dw9.data_ptr(), out_dw=out_dw.data_ptr(), batch=B, h=H, w=W, act=act, eps=1e-5)
    torch.cuda.synchronize()
    xcat = torch.cat([to_nchw_f32(a)] + ([to_nchw_f32(b)] if two else []), 1)
    ref = F.group_norm(xcat, 1, gamma, beta, eps=1e-5)
    if act:
        ref = ref * torch.sigmoid(ref)
    assert rel_err(to_nchw_f32(out), ref) < 4e-3
    assert float((to_nchw_f32(out) - ref).abs().max()) < 0.05
    ref_dw = F.conv2d(ref, dw, None, padding=1, groups=C)
    assert rel_err(to_nchw_f32(out_dw), ref_dw) < 5e-3


@pytest.mark.parametrize("H,W,c1,c2", [(24, 40, 64, 32), (8, 8, 128, 128), (64, 64, 32, 32), (10, 18, 64, 0)])
def test_gn_dw_tiled(H, W, c1, c2):
    """smem-tiled GroupNorm + depthwise 3x3 (csrc/elementwise.cu gn_dw_tile_kernel): partial tiles, two sources."""
    B = 2
    a = nhwc_bf16(_rand(B, c1, H, W, seed=14, scale=2.0) + 0.7)
    b = nhwc_bf16(_rand(B, c2, H, W, seed=15) - 0.3) if c2 else None
    C = c1 + c2
    gamma, beta = 1 + 0.1 * _rand(C, seed=16), 0.1 * _rand(C, seed=17)
    dw = _rand(C, 1, 3, 3, seed=18, scale=0.3)
    dw9 = dw.reshape(C, 9).t().contiguous()
    out = torch.zeros(B, H, W, C, dtype=torch.bfloat16, device=DEV)
    out_dw = torch.zeros_like(out)
    sa, sb = _stats_of(a), (_stats_of(b) if c2 else None)
    _lib.launch("ddif_gn_apply_t", stream(), src1=a.data_ptr(), c1=c1, src2=b.data_ptr() if c2 else None, c2=c2, stats1=sa.data_ptr(),
                stats2=sb.data_ptr() if c2 else None, gamma=gamma.data_ptr(), beta=beta.data_ptr(), out=out.data_ptr(),
                dw_w=dw9.data_ptr(), out_dw=out_dw.data_ptr(), batch=B, h=H, w=W, act=0, eps=1e-5)
    torch.cuda.synchronize()
    xcat = torch.cat([to_nchw_f32(a)] + ([to_nchw_f32(b)] if c2 else []), 1)
    ref = F.group_norm(xcat, 1, gamma, beta, eps=1e-5)
    assert rel_err(to_nchw_f32(out), ref) < 4e-3
    assert rel_err(to_nchw_f32(out_dw), F.conv2d(ref, dw, None, padding=1, groups=C)) < 5e-3


def test_softmax_h_and_attention():
    B, H, W, C = 2, 16, 8, 64
    q = nhwc_bf16(_rand(B, C, H, W, seed=9, scale=2.0))
    out = torch.zeros_like(q)
    _lib.launch("ddif_softmax_h_t", stream(), **{"in": q.data_ptr()}, out=out.data_ptr(), batch=B, h=H, w=W, c=C, scale=0.25)
    ref = to_nchw_f32(q).softmax(dim=-2) * 0.25
    assert rel_err(to_nchw_f32(out), ref) < 4e-3
    # every register-resident height (one pass, H values per thread), the generic two-pass kernel (H = 24), and a pitched input
    # (the first c channels of a wider tensor: q inside the [q | attn_res(x_hat)] output of the FWM q conv)
    for H2, W2, C2, ld in ((64, 16, 64, 0), (32, 8, 96, 0), (8, 8, 256, 0), (24, 8, 64, 0), (24, 8, 64, 96), (64, 16, 64, 96), (16, 8, 128, 192)):
        wide = nhwc_bf16(_rand(B, ld or C2, H2, W2, seed=12, scale=3.0))
        out2 = torch.zeros(B, H2, W2, C2, dtype=torch.bfloat16, device=DEV)
        _lib.launch("ddif_softmax_h_t", stream(), **{"in": wide.data_ptr()}, out=out2.data_ptr(), batch=B, h=H2, w=W2, c=C2, scale=0.5, in_ld=ld)
        ref2 = to_nchw_f32(wide)[:, :C2].softmax(dim=-2) * 0.5
        assert rel_err(to_nchw_f32(out2), ref2) < 4e-3, (H2, W2, C2, ld)
    for ntok_hw, Cc in (((8, 8), 128), ((12, 8), 64)):
        h, w = ntok_hw
        qkv = nhwc_bf16(_rand(B, 3 * Cc, h, w, seed=10))
        o = torch.zeros(B, h, w, Cc, dtype=torch.bfloat16, device=DEV)
        _lib.launch("ddif_attn_t", stream(), qkv=qkv.data_ptr(), out=o.data_ptr(), batch=B, ntok=h * w, c=Cc, heads=8, scale=1 / math.sqrt(Cc))
        t = to_nchw_f32(qkv).view(B, 8, 3 * Cc // 8, h * w)
        hd = Cc // 8
        qq, kk, vv = t[:, :, :hd], t[:, :, hd:2 * hd], t[:, :, 2 * hd:]
        att = torch.softmax(torch.einsum("bncq,bnck->bnqk", qq, kk) / math.sqrt(Cc), -1)
        ref = torch.einsum("bnqk,bnck->bncq", att, vv).reshape(B, Cc, h, w)
        assert rel_err(to_nchw_f32(o), ref) < 5e-3


@pytest.mark.parametrize("hw,B", [((16, 8), 3), ((16, 16), 2), ((32, 32), 2), ((64, 64), 1)])
def test_tcgen05_flash_attention_long_token_counts(hw, B):
    """attn_tc.cu: the tcgen05 / TMEM flash kernel behind DDIF_OP_ATTN for head_dim 16 and ntok % 128 == 0 (128 .. 4096 tokens: the attention
    level of 128x64 patches up to 512x512 whole scenes) against fp32 torch attention (sr3_dwt.py:347-357: scale 1/sqrt(C), not 1/sqrt(d))."""
    h, w = hw
    Cc, heads = 128, 8
    qkv = nhwc_bf16(_rand(B, 3 * Cc, h, w, seed=40 + h, scale=1.5))
    o = torch.zeros(B, h, w, Cc, dtype=torch.bfloat16, device=DEV)
    _lib.launch("ddif_attn_t", stream(), qkv=qkv.data_ptr(), out=o.data_ptr(), batch=B, ntok=h * w, c=Cc, heads=heads, scale=1 / math.sqrt(Cc))
    torch.cuda.synchronize()
    t = to_nchw_f32(qkv).view(B, heads, 3 * Cc // heads, h * w)
    hd = Cc // heads
    qq, kk, vv = t[:, :, :hd], t[:, :, hd:2 * hd], t[:, :, 2 * hd:]
    att = torch.softmax(torch.einsum("bncq,bnck->bnqk", qq, kk) / math.sqrt(Cc), -1)
    ref = torch.einsum("bnqk,bnck->bncq", att, vv).reshape(B, Cc, h, w)
    e = rel_err(to_nchw_f32(o), ref)
    print(f"[attn_tc {h}x{w} B={B}] rel err {e:.4g}")
    assert e < 6e-3, (hw, e)
    # a sharper distribution (large logits): the online-softmax rescale path matters
    qkv2 = nhwc_bf16(_rand(B, 3 * Cc, h, w, seed=41 + h, scale=6.0))
    _lib.launch("ddif_attn_t", stream(), qkv=qkv2.data_ptr(), out=o.data_ptr(), batch=B, ntok=h * w, c=Cc, heads=heads, scale=1 / math.sqrt(Cc))
    t = to_nchw_f32(qkv2).view(B, heads, 3 * Cc // heads, h * w)
    qq, kk, vv = t[:, :, :hd], t[:, :, hd:2 * hd], t[:, :, 2 * hd:]
    att = torch.softmax(torch.einsum("bncq,bnck->bnqk", qq, kk) / math.sqrt(Cc), -1)
    ref = torch.einsum("bnqk,bnck->bncq", att, vv).reshape(B, Cc, h, w)
    e = rel_err(to_nchw_f32(o), ref)
    print(f"[attn_tc sharp {h}x{w}] rel err {e:.4g}")
    assert e < 8e-3, (hw, e)


@pytest.mark.parametrize("H,W,C,ld", [(128, 16, 192, 0), (256, 8, 128, 192), (512, 8, 64, 96), (512, 4, 96, 128)])
def test_softmax_h_tall_images(H, W, C, ld):
    """softmax_h_col_kernel: column split over the threads of a CTA for H = 128 .. 512 (whole-scene mode)."""
    B = 2
    wide = nhwc_bf16(_rand(B, ld or C, H, W, seed=50 + H, scale=3.0))
    out = torch.zeros(B, H, W, C, dtype=torch.bfloat16, device=DEV)
    _lib.launch("ddif_softmax_h_t", stream(), **{"in": wide.data_ptr()}, out=out.data_ptr(), batch=B, h=H, w=W, c=C, scale=0.5, in_ld=ld)
    ref = to_nchw_f32(wide)[:, :C].softmax(dim=-2) * 0.5
    assert rel_err(to_nchw_f32(out), ref) < 4e-3, (H, W, C, ld)


def test_resize_matches_interpolate():
    B, Ct, H, W = 2, 20, 64, 64
    cond = torch.rand(B, Ct, H, W, generator=torch.Generator().manual_seed(11)).to(DEV)
    for oh in (64, 32, 16, 8, 48):
        nh = torch.zeros(B, oh, oh, 16, dtype=torch.bfloat16, device=DEV)
        nc = torch.zeros(B, 11, oh, oh, dtype=torch.float32, device=DEV)
        _lib.launch("ddif_resize_t", stream(), src=cond.data_ptr(), batch=B, c_total=Ct, c0=0, c=9, h=H, w=W, out_h=oh, out_w=oh,
                    dst_nhwc=nh.data_ptr(), c_pad=16, dst_nchw=None)
        _lib.launch("ddif_resize_t", stream(), src=cond.data_ptr(), batch=B, c_total=Ct, c0=9, c=11, h=H, w=W, out_h=oh, out_w=oh,
                    dst_nhwc=None, c_pad=0, dst_nchw=nc.data_ptr())
        ref = F.interpolate(cond, size=(oh, oh), mode="bilinear")
        assert float((nc - ref[:, 9:]).abs().max()) < 1e-6
        assert float((to_nchw_f32(nh)[:, :9] - ref[:, :9]).abs().max()) < 4e-3
        assert float(nh[..., 9:].abs().max()) == 0.0


def test_fwm_context_and_weff():
    B, cd, dim, o, H, W = 2, 11, 96, 64, 16, 16
    c = torch.rand(B, cd, H, W, generator=torch.Generator().manual_seed(12)).to(DEV)
    kv0, kv1, kb = _rand(cd, 1, 3, 3, seed=13, scale=0.3), _rand(2 * dim, cd, 1, 1, seed=14, scale=0.3), _rand(2 * dim, seed=15, scale=0.1)
    wout = _rand(o, dim, seed=16, scale=0.1)
    d = dim // 8
    ctx = torch.zeros(B, 8, d, d, device=DEV)
    _lib.launch("ddif_fwm_context_t", stream(), c_dec=c.data_ptr(), kv0_w=kv0.reshape(cd, 9).contiguous().data_ptr(),
                kv1_w=kv1.reshape(2 * dim, cd).contiguous().data_ptr(), kv1_b=kb.data_ptr(), ctx=ctx.data_ptr(), batch=B, h=H, w=W, cd=cd,
                dim=dim, heads=8)
    kv = F.conv2d(F.conv2d(c, kv0, None, padding=1, groups=cd), kv1, kb)
    k, v = kv.chunk(2, dim=1)
    k = k.softmax(dim=-1).reshape(B, 8, d, H * W)
    ref = torch.einsum("bhdn,bhen->bhde", k, v.reshape(B, 8, d, H * W))
    assert rel_err(ctx, ref) < 1e-4
    # tall image (whole-scene mode): rows split over CTAs, partial contexts added atomically into the zeroed buffer
    H2, W2 = 160, 24
    c2 = torch.rand(B, cd, H2, W2, generator=torch.Generator().manual_seed(21)).to(DEV)
    ctx2 = torch.full((B, 8, d, d), 7.0, device=DEV)  # stale contents must not leak into the result
    _lib.launch("ddif_fwm_context_t", stream(), c_dec=c2.data_ptr(), kv0_w=kv0.reshape(cd, 9).contiguous().data_ptr(),
                kv1_w=kv1.reshape(2 * dim, cd).contiguous().data_ptr(), kv1_b=kb.data_ptr(), ctx=ctx2.data_ptr(), batch=B, h=H2, w=W2, cd=cd,
                dim=dim, heads=8)
    kv2 = F.conv2d(F.conv2d(c2, kv0, None, padding=1, groups=cd), kv1, kb)
    k2, v2 = kv2.chunk(2, dim=1)
    ref2 = torch.einsum("bhdn,bhen->bhde", k2.softmax(dim=-1).reshape(B, 8, d, H2 * W2), v2.reshape(B, 8, d, H2 * W2))
    assert rel_err(ctx2, ref2) < 5e-4   # 3 840-term fp32 sums of values up to ~10^2, partial sums combined with atomics
    weff = torch.zeros(B, o, dim, dtype=torch.bfloat16, device=DEV)
    scale = 1 / math.sqrt(d)
    _lib.launch("ddif_fwm_weff_t", stream(), ctx=ctx.data_ptr(), w_out=wout.data_ptr(), weff=weff.data_ptr(), batch=B, o=o, dim=dim, heads=8,
                o_pad=o, k_pad=dim, scale=scale)
    refw = scale * torch.einsum("ohe,bhde->bohd", wout.view(o, 8, d), ref).reshape(B, o, dim)
    assert rel_err(weff.float(), refw) < 4e-3


def test_time_embed():
    B, ic, nfilm = 4, 32, 200
    w1, b1, w2, b2 = _rand(4 * ic, ic, seed=17, scale=0.2), _rand(4 * ic, seed=18, scale=0.1), _rand(ic, 4 * ic, seed=19, scale=0.1), _rand(ic, seed=20, scale=0.1)
    wf, bf = _rand(nfilm, ic, seed=21, scale=0.2), _rand(nfilm, seed=22, scale=0.1)
    film = torch.zeros(B, nfilm, device=DEV)
    for t in (torch.tensor([0.0, 1.0, 417.0, 499.0]), torch.tensor([998.0, 948.1, 49.9, 3.25])):
        tt = t.to(DEV)
        _lib.launch("ddif_time_embed_t", stream(), time=tt.data_ptr(), w1=w1.data_ptr(), b1=b1.data_ptr(), w2=w2.data_ptr(), b2=b2.data_ptr(),
                    wf=wf.data_ptr(), bf=bf.data_ptr(), film=film.data_ptr(), batch=B, inner=ic, nfilm=nfilm)
        step = torch.arange(ic // 2, dtype=torch.float32, device=DEV) / (ic // 2)
        enc = tt.unsqueeze(1) * torch.exp(-math.log(1e4) * step.unsqueeze(0))
        enc = torch.cat([enc.sin(), enc.cos()], -1)
        h = F.linear(enc, w1, b1)
        te = F.linear(h * torch.sigmoid(h), w2, b2)
        ref = F.linear(te, wf, bf)
        assert float((film - ref).abs().max()) < 2e-4, float((film - ref).abs().max())


# ---- sampler kernels vs the CPU oracle --------------------------------------------------------------------------
def _diffusion(T=500, pred_mode="x_start"):
    class _M:
        self_condition, pred_var = True, False
    d = dp.GaussianDiffusion(_M(), image_size=16, channels=8, pred_mode=pred_mode, loss_type="l1", device=DEV, clamp_range=(0, 1))
    d.set_new_noise_schedule(betas=dp.make_beta_schedule("cosine", T), device=DEV)
    return d


def test_schedule_buffers_match_oracle():
    d = _diffusion()
    sb = so.schedule_buffers(so.make_beta_schedule("cosine", 500))
    for k in so.SCHEDULE_BUFFERS:
        assert torch.equal(getattr(d, k).cpu(), sb[k]), k
    d.space_new_betas(d.space_timesteps(500, "ddim25"))
    sb25 = so.schedule_buffers(so.spaced_betas(sb["alphas_cumprod"], so.space_timesteps(500, "ddim25")))
    for k in so.SCHEDULE_BUFFERS:
        assert torch.equal(getattr(d, k).cpu(), sb25[k]), k


@pytest.mark.parametrize("pred_mode", ["x_start", "noise", "pred_v"])
def test_ddpm_ddim_qsample_steps(pred_mode):
    g = torch.Generator().manual_seed(101)
    x, mo = torch.randn(3, 8, 16, 16, generator=g), torch.randn(3, 8, 16, 16, generator=g) * 0.2
    c, nz = torch.rand(3, 20, 16, 16, generator=g), torch.randn(3, 8, 16, 16, generator=g)
    sb = so.schedule_buffers(so.make_beta_schedule("cosine", 500))
    d = _diffusion(pred_mode=pred_mode)
    for t in (499, 250, 1, 0):
        tt = torch.full((3,), t, dtype=torch.long)
        ref = so.ddpm_step(sb, x, tt, mo, c[:, :8], nz, (0.0, 1.0), pred_mode)
        tout = torch.zeros(3, device=DEV)
        got = d._step("ddpm", x.to(DEV).clone(), mo.to(DEV), c.to(DEV), t, nz.to(DEV), time_out=tout)
        assert float((got.cpu() - ref).abs().max()) <= 1e-6 * max(1.0, float(ref.abs().max())), (pred_mode, t)
        assert float(tout[0]) == max(t - 1, 0)
        for eta in (0.0, 0.5):
            ref = so.ddim_step(sb, x, tt, mo, nz, eta, pred_mode=pred_mode)
            got = d._step("ddim", x.to(DEV).clone(), mo.to(DEV), c.to(DEV), t, nz.to(DEV), eta=eta, clip=False)
            assert float((got.cpu() - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max())), (pred_mode, t, eta)
    tq = torch.tensor([499, 250, 0])
    got = d.q_sample(x.to(DEV), tq.to(DEV), nz.to(DEV))
    assert torch.equal(got.cpu(), so.q_sample(sb, x, tq, nz))


def test_dpmpp_coefficients_and_step():
    sb = so.schedule_buffers(so.make_beta_schedule("cosine", 500))
    ns_o = so.VPSchedule(sb["betas"])
    ns = dp.NoiseScheduleVP("discrete", betas=sb["betas"].to(DEV))
    ts = torch.linspace(1.0, 1 / 500, 21)
    assert torch.equal(ns.marginal_log_mean_coeff(ts), ns_o.log_alpha(ts))
    assert torch.equal(ns.marginal_lambda(ts), ns_o.lam(ts))
    # one fused order-2 step against the oracle's elementwise arithmetic
    g = torch.Generator().manual_seed(3)
    x, out, m1 = (torch.randn(2, 8, 16, 16, generator=g) for _ in range(3))
    t0, t1, t2 = ts[3], ts[4], ts[5]
    a, s = ns_o.alpha(t1), ns_o.std(t1)
    noise = (x - a * out) / s
    m0 = (x - s * noise) / a
    l0, l1, l2 = ns_o.lam(t0), ns_o.lam(t1), ns_o.lam(t2)
    h0, h = l1 - l0, l2 - l1
    D1 = (1.0 / (h0 / h)) * (m0 - m1)
    phi = torch.expm1(-h)
    ref = (ns_o.std(t2) / ns_o.std(t1)) * x - (ns_o.alpha(t2) * phi) * m0 - 0.5 * (ns_o.alpha(t2) * phi) * D1
    wm = dp.model_wrapper(None, ns, model_type="x_start", guidance_type="classifier-free", condition=None, guidance_scale=1.0)
    sol = dp.DPM_Solver(wm, ns)
    coef = sol._coefficients([t0, t1], t2, 2)
    xd, mc = x.to(DEV).clone(), torch.zeros(2, 8, 16, 16, device=DEV)
    out_d, m1_d = out.to(DEV), m1.to(DEV)  # keep the device tensors alive while the kernel runs
    _lib.launch("ddif_dpmpp_step_t", stream(), x=xd.data_ptr(), model_out=out_d.data_ptr(), m_cur=mc.data_ptr(),
                m_prev1=m1_d.data_ptr(), m_prev2=None, time_out=None, n=x.numel(), batch=2, order=2, model_type=0,
                alpha_t=float(a), sigma_t=float(s), t_next_in=0.0, **coef)
    torch.cuda.synchronize()
    assert float((mc.cpu() - m0).abs().max()) <= 1e-5 * float(m0.abs().max())
    assert float((xd.cpu() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


@pytest.mark.parametrize("steps,order", [(20, 2), (12, 3), (5, 2), (7, 3), (6, 1)])
def test_dpm_solver_loop_fp32_model(steps, order):
    """Whole multistep solver (warm-up orders, lower_order_final, buffer rotation) against the CPU oracle with a cheap
    analytic fp32 'denoiser', so that no bf16 noise hides a wrong coefficient."""
    def model(x, t, c, sc=None):
        return 0.5 * torch.tanh(x) + 0.1 * torch.sin(t.reshape(-1, 1, 1, 1) / 100.0) + 0.05 * c[:, :8]

    sb = so.schedule_buffers(so.make_beta_schedule("cosine", 500))
    g = torch.Generator().manual_seed(steps * 10 + order)
    x_T, cond = torch.randn(2, 8, 16, 16, generator=g), torch.rand(2, 20, 16, 16, generator=g)
    ref = so.dpmpp_multistep_sample(model, so.VPSchedule(sb["betas"]), x_T.clone(), cond, steps=steps, order=order)
    ns = dp.NoiseScheduleVP("discrete", betas=sb["betas"].to(DEV))
    wm = dp.model_wrapper(model, ns, model_type="x_start", guidance_type="classifier-free", condition=cond.to(DEV), guidance_scale=1.0)
    got = dp.DPM_Solver(wm, ns).sample(x_T.to(DEV), steps=steps, order=order, skip_type="time_uniform", method="multistep")
    err = float((got.cpu() - ref).abs().max()) / float(ref.abs().max())
    print(f"dpm fp32 loop steps={steps} order={order}: max rel err {err:.3g}")
    assert err < 2e-3  # the reference's own round trip divides by alpha_T ~ 1e-4 (fp32 noise ~1e-3 at the first step)


def test_haar_and_cond_assembly():
    for ds in ("wv3", "cave"):
        d = synth.make_batch(ds, 2, seed=21)
        spec = synth.DATASETS[ds]
        lms_dn, pan_dn = d["lms_dn"].float().to(DEV), d["pan_dn"].float().to(DEV)
        cA, (cH, cV, cD) = dp.haar_dwt2(lms_dn, spec.division)
        rA, (rH, rV, rD) = wo.haar_dwt2(d["lms_dn"].numpy())
        for got, ref in ((cA, rA), (cH, rH), (cV, rV), (cD, rD)):
            assert float((got.cpu().double() - torch.tensor(ref / spec.division)).abs().max()) <= 1e-6
        wav = dp.wavelet_channels(lms_dn, pan_dn, spec.division, spec.wavelet_order)
        assert float((wav.cpu() - d["wavelets"]).abs().max()) <= 1e-6
        cond = dp.assemble_cond(d["lms"].to(DEV), d["pan"].to(DEV), wav)
        assert float((cond.cpu() - d["cond"]).abs().max()) <= 2e-6
    # round trip (the only parity an IDWT has: the reference never calls one) at a large, ragged-ish size
    x = _rand(3, 5, 130, 260, seed=23)
    cA, co = dp.haar_dwt2(x)
    assert float((dp.haar_idwt2(cA, co) - x).abs().max()) <= 1e-6
    lin = dp.haar_dwt2(2 * x + 1)[0] - (2 * cA + 2.0)  # linearity: LL of a constant 1 image is 2
    assert float(lin.abs().max()) <= 1e-5
    empty = dp.haar_dwt2(torch.zeros(0, 4, 8, 8, device=DEV))[0]
    assert empty.shape == (0, 4, 4, 4)
    with pytest.raises(ValueError):
        dp.haar_dwt2(torch.zeros(1, 1, 7, 8, device=DEV))
    with pytest.raises(RuntimeError):
        dp.haar_dwt2(torch.zeros(1, 1, 8, 8))  # CPU tensor: no fallback


def test_device_randn_is_standard_normal():
    n = 1 << 22
    z = dp.device_randn((n,), DEV, seed=7, offset=0)
    z2 = dp.device_randn((n,), DEV, seed=7, offset=0)
    z3 = dp.device_randn((n,), DEV, seed=8, offset=0)
    assert torch.equal(z, z2) and not torch.equal(z, z3)
    assert abs(float(z.mean())) < 3e-3 and abs(float(z.std()) - 1) < 3e-3
    assert abs(float((z ** 4).mean()) - 3.0) < 0.05 and torch.isfinite(z).all()
    assert abs(float((z[:-1] * z[1:]).mean())) < 3e-3


def test_fused_attention_block_matches_torch():
    """ddif_attn_block_f32 path (GN + qkv + 64-token 8-head attention + out + residual + statistics in one kernel) against an fp32
    torch restatement of SelfAttention.forward (sr3_dwt.py:330-360) on the same bf16 inputs / weights."""
    import math
    B, C, heads, hd = 5, 128, 8, 16
    g = torch.Generator().manual_seed(17)
    x = (torch.randn(B, C, 8, 8, generator=g) * 1.5 + 0.3)
    xa = nhwc_bf16(x.to(DEV))                                   # [B, 8, 8, C] bf16
    xf = to_nchw_f32(xa)                                          # the bf16-rounded input as fp32 NCHW
    gamma, beta = (torch.rand(C, generator=g) + 0.5).to(DEV), (torch.randn(C, generator=g) * 0.2).to(DEV)
    wqkv = (torch.randn(3 * C, C, 1, 1, generator=g) / math.sqrt(C)).to(DEV)
    wout = (torch.randn(C, C, 1, 1, generator=g) / math.sqrt(C)).to(DEV)
    bout = (torch.randn(C, generator=g) * 0.1).to(DEV)
    stats_in = torch.stack([xf.double().sum(dim=(1, 2, 3)), (xf.double() ** 2).sum(dim=(1, 2, 3))], dim=1).contiguous()
    wq_p = pack_w(wqkv)                                           # [1][384][128] bf16
    pos = torch.arange(C, device=DEV)
    src = 16 * (2 * (pos // 32) + ((pos % 8) // 2) // 2) + 8 * (((pos % 8) // 2) % 2) + 2 * ((pos % 32) // 8) + pos % 2
    wo_p = wout[:, :, 0, 0][:, src].to(torch.bfloat16).contiguous()
    out = torch.zeros(B, 8, 8, C, dtype=torch.bfloat16, device=DEV)
    stats_out = torch.zeros(B, 2, dtype=torch.float64, device=DEV)
    _lib.launch("ddif_attn_block_t", stream(), x=xa.data_ptr(), stats_in=stats_in.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr(),
                wqkv=wq_p.data_ptr(), wout=wo_p.data_ptr(), bout=bout.data_ptr(), out=out.data_ptr(), stats_out=stats_out.data_ptr(), batch=B,
                ntok=64, c=C, heads=heads, scale=1.0 / math.sqrt(C), eps=1e-5)
    torch.cuda.synchronize()
    # fp32 reference with the same roundings of the stored operands (bf16 weights)
    n = F.group_norm(xf, 1, gamma, beta, eps=1e-5)
    qkv = F.conv2d(n, wqkv.to(torch.bfloat16).float()).view(B, heads, 3 * hd, 64)
    q, k, v = qkv.chunk(3, dim=2)
    att = torch.softmax(torch.einsum("bhdq,bhdk->bhqk", q, k) / math.sqrt(C), dim=-1)
    o = torch.einsum("bhqk,bhdk->bhdq", att, v).reshape(B, C, 8, 8)
    ref = F.conv2d(o, wout.to(torch.bfloat16).float(), bout) + xf
    got = to_nchw_f32(out)
    e = rel_err(got, ref)
    print("fused attention block rel err", e)
    assert e < 6e-3
    assert float((stats_out[:, 0] - ref.double().sum(dim=(1, 2, 3))).abs().max()) < 2e-2 * 64 * C ** 0.5
    assert float((stats_out[:, 1] / (ref.double() ** 2).sum(dim=(1, 2, 3)) - 1).abs().max()) < 1e-2
    with pytest.raises(RuntimeError):
        _lib.launch("ddif_attn_block_t", stream(), x=xa.data_ptr(), stats_in=stats_in.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr(),
                    wqkv=wq_p.data_ptr(), wout=wo_p.data_ptr(), bout=bout.data_ptr(), out=out.data_ptr(), stats_out=None, batch=B, ntok=128,
                    c=C, heads=heads, scale=1.0, eps=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("c1,c2,B", [(128, 128, 5), (128, 64, 3)])
def test_fused_fwm_front_matches_torch(c1, c2, B):
    """DDIF_OP_FWM_FRONT (csrc/fwm_front.cu): prenorm_x + DW3x3 + q 1x1 + softmax over H + W_eff[b] . qs + attn_res + bias at 8 x 8 in one launch,
    against the fp32 torch expression of FastAttnCondInjection's front (sr3_dwt.py:507-573) on the same bf16 inputs."""
    torch.manual_seed(c1 + c2 + B)
    dim, o, H = c1 + c2, 128, 8
    x = (torch.randn(B, H, H, c1) * 1.5 + 0.3).to(torch.bfloat16).to(DEV)
    sk = (torch.randn(B, H, H, c2) * 0.7 - 0.2).to(torch.bfloat16).to(DEV)
    st = lambda v: torch.stack([v.double().sum(dim=(1, 2, 3)), (v.double() ** 2).sum(dim=(1, 2, 3))], dim=1).contiguous()
    s1, s2 = st(x), st(sk)
    gamma, beta = (torch.rand(dim) + 0.5).to(DEV), (torch.randn(dim) * 0.1).to(DEV)
    dw = (torch.randn(dim, 1, 3, 3) * 0.3).to(DEV)
    w1 = (torch.randn(dim, dim) / dim ** 0.5).to(torch.bfloat16).to(DEV)
    b1 = (torch.randn(dim) * 0.1).to(DEV)
    weff = (torch.randn(B, o, dim) * 0.2).to(torch.bfloat16).to(DEV)
    wres = (torch.randn(o, dim) / dim ** 0.5).to(torch.bfloat16).to(DEV)
    bias = torch.randn(o).to(DEV)
    out = torch.zeros(B, H, H, o, dtype=torch.bfloat16, device=DEV)
    dw9 = dw.reshape(dim, 9).t().contiguous()
    _lib.launch("ddif_fwm_front_t", stream(), x=x.data_ptr(), skip=sk.data_ptr(), c1=c1, c2=c2, stats1=s1.data_ptr(), stats2=s2.data_ptr(),
                gamma=gamma.data_ptr(), beta=beta.data_ptr(), eps=1e-5, dw_w=dw9.data_ptr(), w1=w1.data_ptr(), w1_ld=dim, b1=b1.data_ptr(),
                weff=weff.data_ptr(), weff_ld=dim, weff_rows=o, wres=wres.data_ptr(), wres_ld=dim, bias=bias.data_ptr(), out=out.data_ptr(),
                out_ld=o, batch=B, h=H, w=H, o=o)
    torch.cuda.synchronize()
    xc = torch.cat([x, sk], dim=3).float().permute(0, 3, 1, 2)                       # NCHW
    xh = F.group_norm(xc, 1, gamma, beta, 1e-5)
    q = F.conv2d(F.conv2d(xh, dw, padding=1, groups=dim), w1.float()[:, :, None, None], b1)
    qs = torch.softmax(q, dim=-2)
    ref = torch.einsum("boc,bchw->bohw", weff.float(), qs) + F.conv2d(xh, wres.float()[:, :, None, None]) + bias[None, :, None, None]
    err = rel_err(out.float().permute(0, 3, 1, 2), ref)
    print(f"[fwm front] dim {dim}: rel err {err:.3e}")
    assert err < 1e-2, err
