"""CPU checks of the host logic: module/state_dict compatibility, weight packing, schedule construction,
workspace liveness packing, and the C-ABI surface (library loads and exports every declared symbol)."""
import ctypes
import os
import subprocess

import pytest
import torch

from dif_pan_b200 import _lib, synth
from dif_pan_b200.plan import Buf, PlanBuilder, pack_lifetimes
from dif_pan_b200.unet import Schedule, UNetSR3, pack_state_dict


@pytest.mark.parametrize("dataset", ["wv3", "gf2", "cave"])
def test_state_dict_compat(dataset):
    kw = synth.unet_kwargs(dataset)
    net = UNetSR3(**kw)
    sd = synth.make_state_dict(0, **kw)
    mine = net.state_dict()
    assert len(mine) == 702 and set(mine) == set(sd)
    for k in sd:
        assert tuple(mine[k].shape) == tuple(sd[k].shape), k
    net.load_state_dict(sd, strict=True)
    # CSM last conv is zero-initialised like the reference (sr3_dwt.py:386-387)
    fresh = UNetSR3(**kw)
    assert float(fresh.downs[1].cond_inj.body[3].weight.detach().abs().sum()) == 0.0


def test_unsupported_options_raise():
    with pytest.raises(NotImplementedError):
        UNetSR3(pred_var=True, norm_groups=1)
    with pytest.raises(NotImplementedError):
        UNetSR3(norm_groups=32)
    net = UNetSR3(**synth.unet_kwargs("wv3")).eval()
    with pytest.raises(RuntimeError):  # no CPU fallback
        net(torch.zeros(1, 8, 64, 64), torch.zeros(1), torch.zeros(1, 20, 64, 64))


def _schedule(dataset, B, H, W):
    kw = synth.unet_kwargs(dataset)
    net = UNetSR3(**kw)
    net.load_state_dict(synth.make_state_dict(0, **kw))
    P = pack_state_dict({k: v.detach() for k, v in net.state_dict().items()}, net)
    addr = {k: 0x10000000 + 4096 * i for i, (k, v) in enumerate(P.items()) if isinstance(v, torch.Tensor)}
    io = dict(x=1 << 40, sc=2 << 40, t=3 << 40, out=4 << 40, cond=5 << 40)
    s = Schedule(net, addr, B, H, W, io)
    s.build_cond()
    s.build_forward(P["_film_offsets"], int(P["film.w"].shape[0]))
    return net, P, s


def test_packed_weights_layout():
    net, P, _ = _schedule("wv3", 2, 64, 64)
    assert P["downs.0.w"].shape == (9, 32, 16) and P["downs.0.w"].dtype == torch.bfloat16
    assert P["final.w"].shape == (9, 16, 32)
    assert P["film.w"].shape == (2272, 32)  # sum of FiLM outputs
    assert P["downs.1.cond_inj.body0.w"].shape == (9, 128, 16)
    assert P["ups.0.cond_inj.q0"].shape == (9, 256)
    w = net.state_dict()["ups.3.cond_inj.ffn.2.weight"]
    w3 = net.state_dict()["ups.3.cond_inj.ffn.3.weight"][:, :, 0, 0]
    ref = torch.einsum("om,mckl->ockl", w3, w).permute(2, 3, 0, 1).reshape(9, w3.shape[0], w.shape[1])
    got = P["ups.3.cond_inj.ffn23.w"][:, : w3.shape[0], : w.shape[1]].float()
    assert torch.allclose(got, ref, atol=2e-3, rtol=2e-2)


@pytest.mark.parametrize("dataset,B,H,W", [("wv3", 2, 64, 64), ("cave", 1, 64, 64), ("gf2", 1, 128, 64)])
def test_schedule_structure(dataset, B, H, W):
    net, P, s = _schedule(dataset, B, H, W)
    kinds = [op.struct for op in s.fwd.ops]
    # 12 enc x (x_conv, 2 conv) + 16 dec x (q1 | qconv, attn, ffn0, ffn23, 2 conv) + mid 2x2 + 8 attn x 2 + conv0 + 3 down + 3 up + final
    # at 64 tokens (64x64 inputs) each SelfAttention block (norm, qkv, core, out) is ONE fused launch (csrc/attn_block.cu)
    fused_attn = (H // 8) * (W // 8) == 64
    # at an 8x8 lowest level the front of the four FWM blocks there (prenorm+dw, q 1x1, softmax_h, attn_out+res) is ONE launch (csrc/fwm_front.cu)
    front = kinds.count("ddif_fwm_front_t")
    assert front == (4 if (H // 8, W // 8) == (8, 8) else 0)
    assert kinds.count("ddif_gemm_t") == 12 * 3 + 16 * 6 + 4 + (0 if fused_attn else 8 * 2) + 1 + 3 + 3 + 1 - 2 * front
    assert kinds.count("ddif_attn_t") == (0 if fused_attn else 8)
    assert kinds.count("ddif_attn_block_t") == (8 if fused_attn else 0)
    # FWM softmax over H is fused into the attn_out GEMM's loader (cs_gemm_tc_kernel) for the merged-q decoder blocks whose level has 16, 32 or
    # 64 lines and a width that is a multiple of 128 / lines; the others keep the stand-alone softmax kernel (8x8 level, 128-line level, dim + o > 192)
    cs = [op for op in s.fwd.ops if op.struct == "ddif_gemm_t" and op.fields.get("a_softmax_h")]
    assert len(cs) == (7 if (H, W) == (128, 64) else 11)
    assert kinds.count("ddif_softmax_h_t") == 16 - len(cs) - front
    for op in cs:
        f = op.fields
        assert f["out_h"] in (16, 32, 64) and f["out_w"] % (128 // f["out_h"]) == 0 and f["taps"][0] == 1 and f["w_per_sample"][0] == 1
        assert f["a_c"][0] % 64 == 0 and f["a_c"][0] == f["w_k"][0] and f["a_c"][0] <= f["a_ld"][0] and f["residual"] is not None
    # GroupNorm+Swish of every 3x3 conv is fused into the conv's loader (all 30 resblocks x 2 + final conv);
    # stand-alone normalisation launches left: 8 attention norms + 5 FWM prenorm(+dw) -- the other 11 decoder blocks get
    # r = attn_res(x_hat) as extra output channels of the q conv (dim + o <= 192), so x_hat is never materialised
    assert kinds.count("ddif_gn_apply_t") == (0 if fused_attn else 8) + 5 - front
    merged = [op for op in s.fwd.ops if op.label.endswith(".qconv") and op.fields["n_valid"] > op.fields["w_k"][0]]
    assert len(merged) == 11
    assert kinds.count("ddif_upsample2x_t") == 3  # nearest x2 as its own kernel + halo conv (faster than the fused LDG loader)
    fused = [op for op in s.fwd.ops if op.struct == "ddif_gemm_t" and op.fields.get("gn_stats") is not None]
    # 30 resblocks x 2 + final conv, + the FWM q path (prenorm -> DW3x3 -> 1x1 composed into one 3x3 conv with the
    # GroupNorm in its loader) of every decoder block with H >= 16, W >= 8, dim <= 192
    qconv = [op for op in fused if op.label.endswith(".qconv")]
    assert len(qconv) == (13 if (H, W) == (128, 64) else 12)
    assert len(fused) - len(qconv) == 30 * 2 + 1
    assert len(s.mod) == 12 and len(s.weff) == 16
    flops = sum(op.flops for op in s.fwd.ops) + sum(op.flops for op in s.cnd.ops)
    assert flops > 0
    for pb in (s.fwd, s.cnd, s.cache):
        total = pb.layout()
        bufs = pb.bufs
        for b in bufs:
            assert b.offset >= 0 and b.offset % 1024 == 0 and b.offset + b.nbytes <= total
        for i, a in enumerate(bufs):
            for b in bufs[i + 1:]:
                live = a.persistent or b.persistent or not (a.last < b.first or b.last < a.first)
                overlap = a.offset < b.offset + max(b.nbytes, 1) and b.offset < a.offset + max(a.nbytes, 1)
                assert not (live and overlap), (a.name, b.name)
    # liveness packing must actually reuse memory
    assert s.fwd.arena_bytes < 0.5 * sum(b.nbytes for b in s.fwd.bufs)


def test_executed_flops_match_reference_count():
    """Executed conv FLOPs per step + cond-only FLOPs track the reference's 8.378 GFLOP (WV3) within the
    known deltas (ffn.2/ffn.3 composed: -conv1x1; FWM context folded into weights)."""
    _, _, s = _schedule("wv3", 1, 64, 64)
    step = sum(op.flops for op in s.fwd.ops)
    cond = sum(op.flops for op in s.cnd.ops)
    # executed vs reference-equivalent work: the FWM q path runs its depthwise 3x3 inside the conv kernel (CUDA cores) and ONE
    # tensor-core tap where dim >= 96; the dim-64 blocks at 64x64 (and the dim-192 block at 16x16) use the composed dense 3x3, which
    # executes 9x the 1x1's FLOPs (+1.0 GFLOP per patch-forward) but is the faster kernel there (unet.py)
    assert 8.2e9 < step < 8.9e9
    ref = sum(op.ref_flops for op in s.fwd.ops)
    assert 7.0e9 < ref < 7.6e9


def test_pack_lifetimes_basic():
    a, b, c = Buf("a", 4096), Buf("b", 4096), Buf("c", 2048)
    a.first, a.last = 0, 1
    b.first, b.last = 2, 3
    c.first, c.last = 1, 2
    total = pack_lifetimes([a, b, c])
    assert a.offset == b.offset and c.offset >= 4096 and total == 4096 + 2048


def test_bad_sizes_rejected():
    kw = synth.unet_kwargs("wv3")
    net = UNetSR3(**kw)
    with pytest.raises(ValueError):
        Schedule(net, {}, 1, 60, 64, dict(x=0, sc=0, t=0, out=0, cond=0))
    with pytest.raises(ValueError):
        Schedule(net, {}, 1, 32, 32, dict(x=0, sc=0, t=0, out=0, cond=0))


def test_cabi_exports_and_struct_sizes():
    lib = _lib.load()
    assert lib.ddif_version() >= 100
    syms = subprocess.run(["nm", "-D", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in _lib.EXPORTS:
        assert f" T {name}" in syms, name
        getattr(lib, name)
    for name, st in _lib.STRUCTS.items():
        assert ctypes.sizeof(st) % 8 == 0 and all(ctypes.sizeof(t) % 8 == 0 for _, t in st._fields_), name
    assert lib.ddif_error_string(-2).decode().startswith("ddif")
    # argument validation works without a GPU (no kernel is launched)
    st = _lib.make("ddif_haar_t", planes=1, h=3, w=4, divisor=1.0)
    assert lib.ddif_haar_dwt2_f32(ctypes.byref(st), None) == -2
    p = lib.ddif_plan_create()
    assert lib.ddif_plan_size(ctypes.c_void_p(p)) == 0
    lib.ddif_plan_destroy(ctypes.c_void_p(p))


def test_dpm_solver_host_logic_singlestep_orders_and_times():
    """Host side of the singlestep solvers (no GPU): the "DPM-Solver-fast" order schedule of dpm_solver.py:490-548 (cases from its
    docstring) and the outer time grid; stage coefficients are finite and consistent between the orders that share evaluations."""
    import math
    from dif_pan_b200.dpm_solver import DPM_Solver, NoiseScheduleVP, model_wrapper
    from dif_pan_b200.diffusion import make_beta_schedule
    ns = NoiseScheduleVP("discrete", betas=torch.as_tensor(make_beta_schedule("cosine", 500), dtype=torch.float32))
    sol = DPM_Solver(model_wrapper(lambda x, t, c: x, ns, model_type="x_start", guidance_type="classifier-free", condition=torch.zeros(1)), ns)
    t0, tT = 1.0 / ns.total_N, ns.T
    cases = {(9, 3): [3, 3, 2, 1], (10, 3): [3, 3, 3, 1], (11, 3): [3, 3, 3, 2], (6, 2): [2, 2, 2], (7, 2): [2, 2, 2, 1], (4, 1): [1, 1, 1, 1]}
    for (steps, order), want in cases.items():
        outer, orders = sol.get_orders_and_timesteps_for_singlestep_solver(steps, order, "time_uniform", tT, t0)
        assert orders == want and sum(orders) == steps
        grid = torch.linspace(tT, t0, steps + 1)
        assert torch.equal(outer, grid[torch.cumsum(torch.tensor([0] + orders), 0)])
        assert float(outer[0]) == tT and abs(float(outer[-1]) - t0) < 1e-9
    outer, orders = sol.get_orders_and_timesteps_for_singlestep_solver(10, 3, "logSNR", tT, t0)
    assert len(outer) == len(orders) + 1 == 5 and all(float(outer[i]) > float(outer[i + 1]) for i in range(4))
    with pytest.raises(ValueError):
        sol.get_orders_and_timesteps_for_singlestep_solver(10, 4, "time_uniform", tT, t0)
    s, t = torch.tensor([0.8]), torch.tensor([0.6])
    st2 = sol._single_coefficients(s, t, 2, 1.0 / 3.0, None)
    st3 = sol._single_coefficients(s, t, 3, 1.0 / 3.0, 2.0 / 3.0)
    assert len(st2) == 2 and len(st3) == 3
    assert st2[0][3:5] == st3[0][3:5] and float(st2[0][1]) == float(st3[0][1])      # shared first stage (adaptive order 3 relies on it)
    assert all(math.isfinite(v) for stg in st2 + st3 for v in stg[3:])
    assert st3[-1][1] is None and st3[1][1] is not None                              # last stage ends the step
    with pytest.raises(ValueError):
        sol.sample(torch.zeros(1, 1, 4, 4), method="bogus")
    sol._solver_type = "taylor"
    assert len(sol._single_coefficients(s, t, 2, 0.5, None)) == 2
    with pytest.raises(NotImplementedError):
        sol._single_coefficients(s, t, 3, 1.0 / 3.0, 2.0 / 3.0)   # the three-value third-order 'taylor' form is not on the CUDA path
    sol._solver_type = "dpmsolver"
    with pytest.raises(RuntimeError):
        sol.sample(torch.zeros(1, 1, 4, 4), steps=4, order=2, method="singlestep")   # CPU tensor: no fallback


def test_every_op_kind_has_a_struct_and_named_wrappers_exist():
    """Header <-> binding consistency: every DDIF_OP_* kind is reachable from Python through exactly one parameter struct (two kinds
    share ddif_haar_t), and the named convenience wrappers the header declares are exported by the library."""
    kinds = set(_lib.KINDS)
    mapped = set(_lib.KIND_OF_STRUCT.values()) | {"DDIF_OP_HAAR_DWT2", "DDIF_OP_HAAR_IDWT2"}
    assert kinds == mapped, kinds ^ mapped
    assert set(_lib.KIND_OF_STRUCT) | {"ddif_haar_t"} == set(_lib.STRUCTS)
    lib = _lib.load()
    for name in _lib.EXPORTS:
        assert getattr(lib, name) is not None
    assert len(_lib.EXPORTS) >= 32 and len(kinds) == 34   # round 2 added DDIF_OP_WGRAD, DDIF_OP_COLSUM, DDIF_OP_FWM_FRONT


def test_plan_side_branch_argument_checks():
    """ddif_plan_set_side_branch (include/ddif_b200.h) validates its op ranges on the host: no CUDA call is made until the graph is built."""
    import ctypes
    lib = _lib.load()
    plan = ctypes.c_void_p(lib.ddif_plan_create())
    try:
        ms = _lib.make("ddif_memset_t", ptr=None, bytes=0)
        for _ in range(5):
            assert lib.ddif_plan_add(plan, _lib.KINDS["DDIF_OP_MEMSET"], ctypes.byref(ms)) >= 0
        assert lib.ddif_plan_set_side_branch(plan, 1, 2, 4) == 0          # ops [1, 2) beside ops 2, 3; joined before op 4
        assert lib.ddif_plan_set_side_branch(plan, -1, 0, 0) == 0         # clear
        for bad in ((2, 2, 4), (1, 3, 2), (1, 2, 5), (3, 2, 4)):         # empty range, join inside the branch, join past the end, reversed
            assert lib.ddif_plan_set_side_branch(plan, *bad) < 0, bad
        assert lib.ddif_plan_set_side_branch(None, 1, 2, 4) < 0
    finally:
        lib.ddif_plan_destroy(plan)
