"""GPU parity of the tcgen05 implicit-GEMM conv kernel (csrc/gemm_tc.cu) against torch fp32 convolution of the
same bf16-rounded operands and against the CUDA-core direct conv checker.  Tolerance: the output is stored in bf16
(relative rounding 2^-9), accumulation is fp32 -> |err| <= 2^-7 |ref| + 2e-3 * max|ref|."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gpu_util import DEV, gemm, nhwc_bf16, pack_w, rel_err, stream, to_nchw_f32  # noqa: E402
from dif_pan_b200 import _lib  # noqa: E402


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def _close(got, ref, what=""):
    tol = (ref.abs() * 2 ** -7 + 2e-3 * ref.abs().max()).double()
    err = (got.double() - ref.double()).abs()
    bad = int((err > tol).sum())
    assert bad == 0, f"{what}: {bad}/{err.numel()} elements off, max err {float(err.max()):.4g}, rel {rel_err(got, ref):.3g}"


CASES = [
    # name,        B,  H,  W, Cin, Cout, k, stride
    ("c32_sw64",   2, 16, 16,  32,  32, 3, 1),
    ("c64_sw128",  2, 32, 32,  64,  64, 3, 1),
    ("c16_sw32",   1, 64, 64,  16,  32, 3, 1),
    ("c96_k32",    1, 64, 64,  96,  64, 1, 1),
    ("c128_n384",  3,  8,  8, 128, 384, 1, 1),
    ("c128_3x3_8", 5,  8,  8, 128, 128, 3, 1),
    ("c256_n128",  2,  8,  8, 256, 128, 3, 1),
    ("s2_64",      2, 32, 32,  64,  64, 3, 2),
    ("s2_32_big",  3, 64, 64,  32,  32, 3, 2),
    ("n512",       1,  8,  8,  16, 512, 3, 1),
    ("rect",       1, 64, 128, 32,  64, 3, 1),
    ("n96",        1, 64, 64,  64,  96, 1, 1),
    ("n192",       2, 16, 16,  32, 192, 3, 1),
    ("n48",        2, 16, 16,  32,  48, 3, 1),
    ("many_tiles", 40, 32, 32, 32,  32, 3, 1),   # > 148 CTAs: every persistent CTA walks several tiles
]


@pytest.mark.parametrize("name,B,H,W,Cin,Cout,k,stride", CASES, ids=[c[0] for c in CASES])
def test_conv_matches_torch(name, B, H, W, Cin, Cout, k, stride):
    x = _rand(B, Cin, H, W, seed=1)
    w = _rand(Cout, Cin, k, k, seed=2, scale=1.0 / math.sqrt(Cin * k * k))
    bias = _rand(Cout, seed=3)
    xa = nhwc_bf16(x)
    wp = pack_w(w)
    out, _ = gemm([xa], [wp], Cout, taps=[k * k], stride=stride, bias=bias)
    ref = F.conv2d(to_nchw_f32(xa), w.to(torch.bfloat16).float(), bias, stride=stride, padding=k // 2)
    _close(to_nchw_f32(out), ref, name)


FUSED = [
    # name,          B,  H,  W, Cin, Cout
    ("f32_32",       3, 64, 64,  32,  32),   # span 64, resident weights, many tiles per CTA
    ("f64_64",       2, 32, 32,  64,  64),   # span 128, resident weights
    ("f64_128",      2, 32, 32,  64, 128),   # streamed weights
    ("f128_64",      2, 16, 16, 128,  64),   # two K slabs, streamed weights
    ("f96_64",       1, 32, 32,  96,  64),   # three K slabs of 32
    ("f16_32",       2, 64, 64,  16,  32),   # span 32
    ("f_ragged",     2, 24, 40,  32,  48),   # partial tiles in x and y
    ("f256_128",     1, 16, 16, 256, 128),
]


@pytest.mark.parametrize("name,B,H,W,Cin,Cout", FUSED, ids=[c[0] for c in FUSED])
def test_fused_gn_swish_conv3x3(name, B, H, W, Cin, Cout):
    """GroupNorm(1 group)+Swish fused into the 3x3 conv loader vs torch, and the plain fused kernel vs the TMA kernel."""
    x = _rand(B, Cin, H, W, seed=21, scale=1.5) + 0.4
    w = _rand(Cout, Cin, 3, 3, seed=22, scale=1.0 / math.sqrt(Cin * 9))
    bias = _rand(Cout, seed=23)
    gamma, beta = 1 + 0.1 * _rand(Cin, seed=24), 0.1 * _rand(Cin, seed=25)
    xa, wp = nhwc_bf16(x), pack_w(w)
    xf = to_nchw_f32(xa)
    stats = torch.stack([xf.double().sum(dim=(1, 2, 3)), (xf.double() ** 2).sum(dim=(1, 2, 3))], dim=1).contiguous()
    out, _ = gemm([xa], [wp], Cout, taps=[9], bias=bias, gn=(stats, gamma, beta, 1))
    h = F.group_norm(xf, 1, gamma, beta, eps=1e-5)
    h = (h * torch.sigmoid(h)).to(torch.bfloat16).float()  # the kernel feeds the MMA bf16 operands
    ref = F.conv2d(h, w.to(torch.bfloat16).float(), bias, padding=1)
    got = to_nchw_f32(out)
    assert rel_err(got, ref) < 6e-3, rel_err(got, ref)
    # a bf16 flip of the normalised operand moves single outputs by ~2^-8 |h| |w|: bound the tail loosely
    assert float((got - ref).abs().max()) < 0.06 * float(ref.abs().max())
    plain, _ = gemm([xa], [wp], Cout, taps=[9], bias=bias)
    tma, _ = gemm([xa], [wp], Cout, taps=[9], bias=bias, force_tma=1)
    _close(to_nchw_f32(plain), F.conv2d(xf, w.to(torch.bfloat16).float(), bias, padding=1), name + " plain")
    _close(to_nchw_f32(plain), to_nchw_f32(tma), name + " fused-vs-tma")


CONCAT = [
    # name,        B,  H,  W, c1, c2, Cout, act
    ("cat64_32",   2, 32, 32, 64, 32,  96, 0),   # kslab 32, three slabs, N = 96 in one CTA
    ("cat32_32",   3, 64, 64, 32, 32,  64, 0),
    ("cat64_64",   2, 32, 32, 64, 64, 128, 0),   # kslab 64, N split over two CTAs
    ("cat128_64",  2, 16, 16, 128, 64, 192, 1),  # three slabs of 64, N split
    ("cat_ragged", 1, 24, 40, 32, 16,  48, 1),   # kslab 16, partial tiles
    ("cat_many32", 40, 64, 64, 32, 32,  64, 0),  # every persistent CTA walks several (tile, slab) units
    ("cat_many64", 48, 32, 32, 64, 64, 128, 1),  # ... with the N split and three stages
    ("cat_many192", 80, 16, 16, 128, 64, 192, 0),
]


@pytest.mark.parametrize("name,B,H,W,c1,c2,Cout,act", CONCAT, ids=[c[0] for c in CONCAT])
def test_concat_gn_conv3x3(name, B, H, W, c1, c2, Cout, act):
    """Two K segments = channel concat of two tensors, GroupNorm over the concatenation fused into the loader
    (FWM q = Conv1x1(DW3x3(GN(cat(x, skip)))) composed into one dense 3x3; sr3_dwt.py:212,507-541)."""
    x1 = _rand(B, c1, H, W, seed=41, scale=1.5) + 0.4
    x2 = _rand(B, c2, H, W, seed=42, scale=0.7) - 0.2
    C = c1 + c2
    w = _rand(Cout, C, 3, 3, seed=43, scale=1.0 / math.sqrt(C * 9))
    bias = _rand(Cout, seed=44)
    gamma, beta = 1 + 0.1 * _rand(C, seed=45), 0.1 * _rand(C, seed=46)
    a1, a2 = nhwc_bf16(x1), nhwc_bf16(x2)
    f1, f2 = to_nchw_f32(a1), to_nchw_f32(a2)
    st = lambda f: torch.stack([f.double().sum(dim=(1, 2, 3)), (f.double() ** 2).sum(dim=(1, 2, 3))], dim=1).contiguous()
    s1, s2 = st(f1), st(f2)
    w1, w2 = pack_w(w[:, :c1].contiguous()), pack_w(w[:, c1:].contiguous())
    out, stats = gemm([a1, a2], [w1, w2], Cout, taps=[9, 9], bias=bias, gn=(s1, gamma, beta, act, s2), want_stats=True)
    h = F.group_norm(torch.cat([f1, f2], 1), 1, gamma, beta, eps=1e-5)
    if act:
        h = h * torch.sigmoid(h)
    ref = F.conv2d(h.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), bias, padding=1)
    got = to_nchw_f32(out)
    assert rel_err(got, ref) < 6e-3, rel_err(got, ref)
    assert float((got - ref).abs().max()) < 0.06 * float(ref.abs().max())
    s_ref = torch.stack([ref.double().sum(dim=(1, 2, 3)), (ref.double() ** 2).sum(dim=(1, 2, 3))], dim=1)
    assert torch.allclose(stats, s_ref, rtol=5e-3, atol=1.0), (stats, s_ref)
    # no GroupNorm: plain concat conv
    plain, _ = gemm([a1, a2], [w1, w2], Cout, taps=[9, 9], bias=bias)
    _close(to_nchw_f32(plain), F.conv2d(torch.cat([f1, f2], 1), w.to(torch.bfloat16).float(), bias, padding=1), name + " plain")


DWQ = [
    # name, B, H, W, c1, c2 (0 = one segment), o (attn_res channels, 0 = none), act
    ("dw64_32",    3, 64, 64, 32, 32, 32, 0),   # kslab 32: 64x64 level (dim 64, o 32)
    ("dw96_32",    2, 64, 64, 64, 32, 32, 0),   # segments of different width (kslab 32)
    ("dw128_64",   2, 32, 32, 64, 64, 64, 0),   # kslab 64: two columns per thread
    ("dw48_16",    2, 24, 40, 32, 16, 16, 1),   # kslab 16, partial tiles, Swish
    ("dw_single",  2, 16, 16, 64, 0, 0, 0),     # one segment, no attn_res part
    ("dw_many",   40, 64, 64, 32, 32, 32, 0),   # persistent CTAs walk several tiles (both A buffers, both rings)
    ("dw_many128", 48, 32, 32, 64, 64, 64, 0),
]


@pytest.mark.parametrize("name,B,H,W,c1,c2,o,act", DWQ, ids=[c[0] for c in DWQ])
def test_depthwise_q_path(name, B, H, W, c1, c2, o, act):
    """Depthwise mode: [q | r] = [Conv1x1(DW3x3(x_hat)) | attn_res(x_hat)], x_hat = GN(cat(x, skip)), with the depthwise 3x3
    computed inside the kernel and ONE tensor-core tap (sr3_dwt.py:507-517,541,573)."""
    x1 = _rand(B, c1, H, W, seed=51, scale=1.5) + 0.4
    C = c1 + c2
    wdw = _rand(C, 1, 3, 3, seed=53, scale=0.4)
    wq = _rand(C, C, 1, 1, seed=54, scale=1.0 / math.sqrt(C))
    wr = _rand(o, C, 1, 1, seed=55, scale=1.0 / math.sqrt(C)) if o else None
    bias = _rand(C + o, seed=56)
    gamma, beta = 1 + 0.1 * _rand(C, seed=57), 0.1 * _rand(C, seed=58)
    st = lambda f: torch.stack([f.double().sum(dim=(1, 2, 3)), (f.double() ** 2).sum(dim=(1, 2, 3))], dim=1).contiguous()
    a1 = nhwc_bf16(x1)
    f1 = to_nchw_f32(a1)
    srcs, feats, stats = [a1], [f1], [st(f1)]
    if c2:
        a2 = nhwc_bf16(_rand(B, c2, H, W, seed=52, scale=0.7) - 0.2)
        srcs.append(a2); feats.append(to_nchw_f32(a2)); stats.append(st(feats[1]))
    w11 = torch.cat([wq, wr], 0) if o else wq
    ws = [pack_w(w11[:, :c1].contiguous())] + ([pack_w(w11[:, c1:].contiguous())] if c2 else [])
    dw_flat = wdw.reshape(C, 9).t().contiguous()  # [9][C]
    out, _ = gemm(srcs, ws, C + o, taps=[9] * len(srcs), w_s=[1] * len(srcs), bias=bias,
                  gn=(stats[0], gamma, beta, act, stats[1] if c2 else None), dw=(dw_flat, C))
    h = F.group_norm(torch.cat(feats, 1), 1, gamma, beta, eps=1e-5)
    if act:
        h = h * torch.sigmoid(h)
    hb = h.to(torch.bfloat16).float()
    d = F.conv2d(hb, wdw, None, padding=1, groups=C).to(torch.bfloat16).float()
    ref = F.conv2d(d, wq.to(torch.bfloat16).float(), bias[:C])
    if o:
        ref = torch.cat([ref, F.conv2d(hb, wr.to(torch.bfloat16).float(), bias[C:])], 1)
    got = to_nchw_f32(out)
    assert rel_err(got[:, :C], ref[:, :C]) < 6e-3, rel_err(got[:, :C], ref[:, :C])
    if o:
        assert rel_err(got[:, C:], ref[:, C:]) < 6e-3, rel_err(got[:, C:], ref[:, C:])
    assert float((got - ref).abs().max()) < 0.06 * float(ref.abs().max())


def test_upsample_kernel_conv_and_epilogue():
    B, H, W, C = 2, 16, 16, 64
    x = _rand(B, C, H, W, seed=31)
    w, bias = _rand(C, C, 3, 3, seed=32, scale=0.05), _rand(C, seed=33)
    xa, wp = nhwc_bf16(x), pack_w(w)
    # Upsample = nearest x2 (its own kernel) -> Conv3x3 (sr3_dwt.py:266-273); a_up (a loader-side fold) is reserved and rejected
    up = torch.empty(B, 2 * H, 2 * W, C, dtype=torch.bfloat16, device=DEV)
    _lib.launch("ddif_upsample2x_t", stream(), **{"in": xa.data_ptr()}, out=up.data_ptr(), batch=B, h=H, w=W, c=C)
    out, stats = gemm([up], [wp], C, taps=[9], bias=bias, want_stats=True)
    ref = F.conv2d(F.interpolate(to_nchw_f32(xa), scale_factor=2, mode="nearest"), w.to(torch.bfloat16).float(), bias, padding=1)
    _close(to_nchw_f32(out), ref, "upsample + conv")
    s_ref = torch.stack([ref.double().sum(dim=(1, 2, 3)), (ref.double() ** 2).sum(dim=(1, 2, 3))], dim=1)
    assert torch.allclose(stats, s_ref, rtol=2e-3, atol=0.5)
    with pytest.raises(RuntimeError):
        gemm([xa], [wp], C, taps=[9], bias=bias, a_up=1)
    # residual + FiLM + fp32 NCHW store with N = 8 (the final conv)
    res = nhwc_bf16(_rand(B, C, H, W, seed=34))
    film = _rand(B, C, seed=35)
    out, _ = gemm([xa], [wp], C, taps=[9], bias=bias, film=film, residual=res)
    ref = F.conv2d(to_nchw_f32(xa), w.to(torch.bfloat16).float(), bias, padding=1) + film[:, :, None, None] + to_nchw_f32(res)
    _close(to_nchw_f32(out), ref, "fused epilogue")
    out8, _ = gemm([xa], [pack_w(w[:8])], 8, taps=[9], bias=bias[:8], want_nchw=True)
    assert rel_err(out8, F.conv2d(to_nchw_f32(xa), w[:8].to(torch.bfloat16).float(), bias[:8], padding=1)) < 1e-4


def test_direct_conv_checker_agrees():
    B, H, W, Cin, Cout = 2, 16, 16, 32, 48
    x, w, bias = _rand(B, Cin, H, W, seed=4), _rand(Cout, Cin, 3, 3, seed=5, scale=0.06), _rand(Cout, seed=6)
    xa, wp = nhwc_bf16(x), pack_w(w)
    out = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16, device=DEV)
    _lib.launch("ddif_conv_direct_t", stream(), **{"in": xa.data_ptr()}, in_ld=Cin, cin=Cin, in_h=H, in_w=W, w=wp.data_ptr(), w_k=Cin,
                taps=9, stride=1, bias=bias.data_ptr(), out=out.data_ptr(), out_ld=Cout, batch=B, out_h=H, out_w=W, n_valid=Cout,
                n_pad=48, act=0)
    tc, _ = gemm([xa], [wp], Cout, taps=[9], bias=bias)
    torch.cuda.synchronize()
    _close(to_nchw_f32(out), to_nchw_f32(tc), "direct vs tcgen05")


def test_epilogue_all_options():
    """bias + FiLM + CSM modulation + residual + SiLU + GroupNorm statistics, and the fp32 NCHW store."""
    B, H, W, Cin, Cout = 3, 16, 16, 64, 32
    x, w = _rand(B, Cin, H, W, seed=7), _rand(Cout, Cin, 1, 1, seed=8, scale=0.12)
    bias, film = _rand(Cout, seed=9), _rand(B, 100, seed=10)
    mod = nhwc_bf16(_rand(B, 2 * Cout, H, W, seed=11, scale=0.5))
    res = nhwc_bf16(_rand(B, Cout, H, W, seed=12))
    xa, wp = nhwc_bf16(x), pack_w(w)
    film_view = film[:, 20:]  # offset view with row pitch 100: pass base pointer of the slice and ld=100

    class _F:  # tiny shim so gemm() sees data_ptr()/shape of the strided view
        def __init__(self, t, ld):
            self.t, self.shape = t, (t.shape[0], ld)

        def data_ptr(self):
            return self.t.data_ptr()

    conv = F.conv2d(to_nchw_f32(xa), w.to(torch.bfloat16).float(), bias) + film[:, 20:20 + Cout, None, None]
    sc, sh = to_nchw_f32(mod)[:, :Cout], to_nchw_f32(mod)[:, Cout:]
    pre = conv * (1 + sc) + sh + to_nchw_f32(res)
    for act in (0, 1):
        ref = F.silu(pre) if act else pre
        out, stats = gemm([xa], [wp], Cout, taps=[1], bias=bias, film=_F(film_view, 100), mod=mod, residual=res, act=act, want_stats=True)
        _close(to_nchw_f32(out), ref, f"epilogue act={act}")
        s_ref = torch.stack([ref.double().sum(dim=(1, 2, 3)), (ref.double() ** 2).sum(dim=(1, 2, 3))], dim=1)
        assert torch.allclose(stats, s_ref, rtol=2e-3, atol=0.5), (stats, s_ref)
    out, _ = gemm([xa], [pack_w(w[:8])], 8, taps=[1], bias=bias[:8], want_nchw=True)
    ref8 = F.conv2d(to_nchw_f32(xa), w[:8].to(torch.bfloat16).float(), bias[:8])
    assert rel_err(out, ref8) < 1e-4  # fp32 store: only accumulation-order error


@pytest.mark.parametrize("H", [8, 16, 32])
def test_dual_segment_per_sample_weights(H):
    """attn_out(W_eff[b] q) + attn_res(x_hat): two K segments, first with per-sample weights (64- and 128-row tiles)."""
    B, dim, o = 3, 96, 64
    q, xh = _rand(B, dim, H, H, seed=13), _rand(B, dim, H, H, seed=14)
    weff = (_rand(B, o, dim, seed=15, scale=0.1)).to(torch.bfloat16).contiguous()
    wres = _rand(o, dim, 1, 1, seed=16, scale=0.1)
    bias = _rand(o, seed=17)
    qa, xa = nhwc_bf16(q), nhwc_bf16(xh)
    out, _ = gemm([qa, xa], [weff, pack_w(wres)], o, taps=[1, 1], bias=bias, per_sample=[1, 0])
    ref = torch.einsum("bok,bkhw->bohw", weff.float(), to_nchw_f32(qa)) + F.conv2d(to_nchw_f32(xa), wres.to(torch.bfloat16).float(), bias)
    _close(to_nchw_f32(out), ref, f"dual H={H}")


def test_bad_arguments_return_errors():
    x = nhwc_bf16(_rand(1, 32, 16, 16))
    w = pack_w(_rand(32, 32, 3, 3))
    with pytest.raises(RuntimeError):
        gemm([x], [w], 32, taps=[4])           # taps must be 1 or 9
    with pytest.raises(RuntimeError):
        gemm([x], [w], 32, taps=[9], stride=3)  # stride 1|2
    with pytest.raises(RuntimeError):
        gemm([nhwc_bf16(_rand(1, 24, 16, 16))], [w], 32, taps=[9])  # Cin % 16

@pytest.mark.parametrize("shape", [(3, 64, 64, 64, 32), (2, 64, 64, 96, 32), (3, 32, 32, 128, 64), (2, 32, 32, 96, 64), (5, 16, 16, 128, 64),
                                   (2, 16, 16, 256, 128), (2, 64, 16, 64, 32), (2, 16, 64, 64, 48)])
def test_softmax_h_fused_into_per_sample_gemm(shape):
    """FWM `q.softmax(dim=-2)` + per-sample attn_out + bias + attn_res residual (sr3_dwt.py:541-573) in ONE kernel (cs_gemm_tc_kernel):
    q and r live in one [B, H, W, dim + o] tensor like the q conv writes them; checked against the fp32 torch expression on the same
    bf16 inputs and against the two-kernel path (DDIF_OP_SOFTMAX_H + plain GEMM)."""
    B, H, W, dim, o = shape
    g = torch.Generator().manual_seed(sum(shape))
    qr = (torch.randn(B, H, W, dim + o, generator=g) * 2.0).to(torch.bfloat16).to(DEV)
    weff = (torch.randn(B, (o + 15) // 16 * 16, dim, generator=g) * 0.2).to(torch.bfloat16).to(DEV)
    weff[:, o:] = 0
    bias = torch.randn(o, generator=g).to(DEV)
    out, _ = gemm([qr], [weff], o, taps=[1], bias=bias, per_sample=[1], w_s=[B], a_c=[dim], residual=qr, residual_off=dim, softmax_h=True)
    q, r = qr[..., :dim].float(), qr[..., dim:].float()
    ref = torch.einsum("bhwc,boc->bhwo", torch.softmax(q, dim=1), weff[:, :o].float()) + bias + r
    assert rel_err(out[..., :o].float(), ref) < 6e-3, rel_err(out[..., :o].float(), ref)
    # two-kernel path on the same inputs
    qs = torch.zeros(B, H, W, dim, dtype=torch.bfloat16, device=DEV)
    _lib.launch("ddif_softmax_h_t", stream(), **{"in": qr.data_ptr()}, out=qs.data_ptr(), batch=B, h=H, w=W, c=dim, scale=1.0, in_ld=dim + o)
    out2, _ = gemm([qs], [weff], o, taps=[1], bias=bias, per_sample=[1], w_s=[B], residual=qr, residual_off=dim)
    assert rel_err(out[..., :o].float(), out2[..., :o].float()) < 4e-3
    # no residual, shared weights
    out3, _ = gemm([qr], [weff[:1].contiguous()], o, taps=[1], bias=bias, a_c=[dim], softmax_h=True)
    ref3 = torch.einsum("bhwc,oc->bhwo", torch.softmax(q, dim=1), weff[0, :o].float()) + bias
    assert rel_err(out3[..., :o].float(), ref3) < 6e-3


def test_softmax_h_gemm_rejects_unsupported_shapes():
    qr = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16, device=DEV)
    w = torch.zeros(1, 32, 64, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError):
        gemm([qr], [w], 32, taps=[1], softmax_h=True)  # 8 lines: a tile would need 16 columns
