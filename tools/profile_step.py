#!/usr/bin/env python
"""Per-op timing of one UNet denoise step (CUDA events around every launch of the recorded plan) + whole-step timing
(eager op-by-op vs CUDA graph).  Writes gpurun_out/step_profile_B<batch>.json and prints a per-category table."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dif_pan_b200 as dp  # noqa: E402
from dif_pan_b200 import synth  # noqa: E402


def category(label, struct):
    if struct == "ddif_gemm_t":
        for key in ("x_conv", "qconv", "q1", "attn_out", "ffn0", "ffn23", "gn+conv", "block1.conv", "block2.conv", "qkv", ".out", "final", "downs.0",
                    "up+conv"):
            if key in label:
                return "gemm:" + key
        return "gemm:resample"
    return struct.replace("ddif_", "").replace("_t", "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--dataset", default="wv3")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--size", type=int, default=64, help="H = W of the input (512 with --batch 1 = whole-scene mode of test_fn)")
    ap.add_argument("--top", type=int, default=12)
    ap.add_argument("--out", default="gpurun_out")
    ap.add_argument("--no-cs-gemm", action="store_true", help="A/B: stand-alone softmax_h kernel + plain attn_out GEMM instead of the fused cs_gemm_tc_kernel")
    ap.add_argument("--no-fwm-front", action="store_true", help="A/B: four launches for the FWM front at 8x8 instead of the fused fwm_front64_kernel")
    ap.add_argument("--no-side-branch", action="store_true", help="A/B: time embedding in line instead of as a side branch of the step graph")
    ap.add_argument("--ncu", action="store_true", help="run cond build + ONE eager step between cudaProfilerStart/Stop and exit")
    a = ap.parse_args()
    torch.set_grad_enabled(False)
    dev = "cuda:0"
    kw = synth.unet_kwargs(a.dataset)
    net = dp.UNetSR3(**kw)
    net.load_state_dict(synth.make_state_dict(0, **kw))
    net = net.to(dev).eval()
    B = a.batch
    if a.no_cs_gemm or a.no_side_branch or a.no_fwm_front:
        from dif_pan_b200.unet import _Runtime
        net.pack_weights()
        net._rt = _Runtime(net, B, a.size, a.size, use_cs_gemm=not a.no_cs_gemm, use_side_branch=not a.no_side_branch,
                           use_fwm_front=not a.no_fwm_front)
    rt = net.runtime(B, a.size, a.size)
    cond = synth.make_batch(a.dataset, min(B, 16), size=a.size, seed=1)["cond"]
    cond = cond.repeat((B + cond.shape[0] - 1) // cond.shape[0], 1, 1, 1)[:B].contiguous().to(dev)
    t0 = time.time()
    rt.set_cond(cond)
    torch.cuda.synchronize()
    print(f"arena: ws {rt.ws.numel() / 2**20:.0f} MiB, cond cache {rt.cache.numel() / 2**20:.0f} MiB; cond plan {len(rt.sch.cnd)} ops")
    dp.diffusion.device_randn_(rt.x_buf, 1, 0)
    rt.t_buf.fill_(250.0)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        s, e = ev(), ev()
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps

    if a.ncu:
        rt.use_graph = False
        rt.step()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        rt.step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    rt.use_graph = False
    t_eager = timed(rt.step, a.reps)
    rt.use_graph = True
    t_graph = timed(rt.step, a.reps)
    t_cond = timed(lambda: rt.set_cond(cond, force=True), 2)
    rows = rt.sch.fwd.profile(rt.stream)
    rows = rt.sch.fwd.profile(rt.stream)
    cats = {}
    for label, struct, ms, fl, by in rows:  # noqa
        c = cats.setdefault(category(label, struct), dict(ms=0.0, flops=0.0, bytes=0.0, n=0))
        c["ms"] += ms; c["flops"] += fl; c["bytes"] += by; c["n"] += 1
    tot = sum(c["ms"] for c in cats.values())
    print(f"B={B}: step eager {t_eager:.3f} ms, graph {t_graph:.3f} ms, sum of per-op events {tot:.3f} ms, cond cache build {t_cond:.3f} ms")
    print(f"{'category':<22}{'n':>4}{'ms':>9}{'%':>7}{'TFLOP/s':>10}{'GB/s':>9}")
    for k, c in sorted(cats.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"{k:<22}{c['n']:>4}{c['ms']:>9.3f}{100 * c['ms'] / tot:>7.1f}{c['flops'] / c['ms'] / 1e9 if c['ms'] else 0:>10.1f}{c['bytes'] / c['ms'] / 1e6 if c['ms'] else 0:>9.0f}")
    gm = [r for r in rows if r[1] == "ddif_gemm_t"]
    gms, gfl = sum(r[2] for r in gm), sum(r[3] for r in gm)
    print(f"gemm total: {gms:.3f} ms, {gfl / gms / 1e9:.1f} TFLOP/s executed")
    os.makedirs(a.out, exist_ok=True)
    json.dump(dict(batch=B, eager_ms=t_eager, graph_ms=t_graph, cond_ms=t_cond, ops=[dict(label=r[0], struct=r[1], ms=r[2], flops=r[3], bytes=r[4]) for r in rows]),
              open(os.path.join(a.out, f"step_profile_B{B}_{a.size}.json"), "w"))
    worst = sorted(rows, key=lambda r: -r[2])[:a.top]
    for r in worst:
        print(f"  {r[0] or r[1]:<40}{r[2]:>8.3f} ms {r[3] / r[2] / 1e9:>8.1f} TF/s {r[4] / r[2] / 1e6:>8.0f} GB/s")


if __name__ == "__main__":
    main()
