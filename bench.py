#!/usr/bin/env python
"""Benchmark of the DDIF denoising hot path on B200 (see the contract in DESIGN.md §Measurement).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU (oracle port)

Workload (BASELINE.json configs[1]): WV3-shaped synthetic 64x64 8-band patches, batch 256 PER GPU, full DDPM
sampling loop (cosine schedule, T = 500, clip to (0,1), x_start prediction, self-conditioning on the current image),
bf16 UNet internals / fp32 sampler state.  One "step" = one complete sampling of the batch: cond cache build +
500 x (UNet forward + fused posterior step) + final clip(sample + lms).  `value` = patches/s over all ranks with
inputs resident in HBM; `e2e` = the same through the public API with the conditioning coming from pinned host
memory and the fused images copied back to the host every step.  Multi-GPU: one process per GPU (torchrun), patches
sharded, no data-path collective, final NCCL all_gather of the finished patches inside the timed region.

The headline line is WEAK scaling (256 patches per GPU; the driver computes efficiency from the per-N values).  BASELINE configs[1]
as written -- 256 patches in TOTAL, 256 / N per GPU -- is measured in the same run and reported under `strong` (or as the headline
with --scaling strong).  Extra keys measured OUTSIDE the timed region on rank 0 at N = 1: `cpu_baseline` (the reference algorithm on
the host cores), `gpu_eager_baseline` (the same reference algorithm as eager PyTorch -- cuDNN / cuBLAS -- on this GPU in fp32 and
under bf16 autocast: the real bar, since the reference ships no Blackwell kernel) and `other_configs`
(BASELINE configs[0], [2], [3] through the public API).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "WV3 64x64 patches/s (full DDPM-500 sampling)"
UNIT = "patches/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="patches per GPU")
    ap.add_argument("--timesteps", type=int, default=500)
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: --batch patches per GPU (headline); strong: --batch patches in total")
    ap.add_argument("--no-extras", action="store_true", help="skip strong / gpu_eager_baseline / other_configs (tuning runs)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], 0, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx or None, power_w_max=max(pw) if pw else None,
                    samples=len(sm), reasons=sorted(reasons))


def cpu_reference_run(cpu_batch: int, timesteps: int, steps: int, warmup: int):
    """The reference algorithm (oracle port: torch CPU fp32 restatement pinned to the reference's golden vectors) on all
    host cores: each step = ONE denoise step (UNet forward + DDPM posterior) of `cpu_batch` WV3 patches; patches/s is
    extrapolated to the full T-step sampling (a full run takes ~1 minute per patch)."""
    import torch
    from dif_pan_b200 import synth
    from oracle import sampler_oracle as so, unet_oracle as uo

    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = synth.unet_kwargs("wv3")
    sd = synth.make_state_dict(0, **kw)
    kw2 = dict(kw)
    kw2.pop("dropout")
    cfg = uo.UNetCfg(**kw2)
    cond = synth.make_batch("wv3", cpu_batch, seed=1)["cond"]
    sb = so.schedule_buffers(so.make_beta_schedule("cosine", timesteps))
    g = torch.Generator().manual_seed(0)
    x = torch.randn(cpu_batch, 8, 64, 64, generator=g)
    times = []
    for i in range(warmup + steps):
        t = torch.full((cpu_batch,), timesteps - 1 - i, dtype=torch.long)
        nz = torch.randn(cpu_batch, 8, 64, 64, generator=g)
        t0 = time.perf_counter()
        out = uo.unet_forward(sd, cfg, x, t, cond, x)
        x = so.ddpm_step(sb, x, t, out, cond[:, :8], nz)
        times.append(time.perf_counter() - t0)
    dt = sum(times[warmup:]) / steps
    value = cpu_batch / (timesteps * dt)
    return dict(value=value, unit=UNIT, cores=cores, kind="port", ms_per_denoise_step=dt * 1e3,
                sample=f"{steps} denoise steps (oracle UNet forward + DDPM posterior) at batch {cpu_batch}, fp32, {cores} threads; "
                       f"patches/s extrapolated to T={timesteps}")


def cpu_config0_run():
    """BASELINE configs[0] exactly as written -- WV3 batch 1, DPM-Solver++ 2M with 20 steps, fp32 on the CPU -- with the oracle port on all host
    cores: ms per complete sampling (20 UNet forwards + the solver updates), one warm-up sampling, one timed."""
    import torch
    from dif_pan_b200 import synth
    from oracle import sampler_oracle as so, unet_oracle as uo

    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = synth.unet_kwargs("wv3")
    sd = synth.make_state_dict(0, **kw)
    kw2 = dict(kw)
    kw2.pop("dropout")
    cfg = uo.UNetCfg(**kw2)
    cond = synth.make_batch("wv3", 1, seed=1)["cond"]
    ns = so.VPSchedule(torch.tensor(so.make_beta_schedule("cosine", 500), dtype=torch.float32))
    model = lambda x, t, c, sc: uo.unet_forward(sd, cfg, x, t, c, x)
    x_T = torch.randn(1, 8, 64, 64, generator=torch.Generator().manual_seed(0))
    so.dpmpp_multistep_sample(model, ns, x_T.clone(), cond, steps=2, order=2)
    t0 = time.perf_counter()
    so.dpmpp_multistep_sample(model, ns, x_T.clone(), cond, steps=20, order=2)
    return dict(cpu_ms_per_sampling=(time.perf_counter() - t0) * 1e3, cores=cores, kind="port")


def gpu_eager_run(dev, batches=(1, 32, 256), nsteps=20):
    """The reference algorithm (oracle port = functional restatement of models/sr3_dwt.py + the DDPM posterior, pinned to the reference's
    golden vectors) as EAGER PyTorch on this GPU: cuDNN convolutions (channels-last, cudnn.benchmark), native group_norm, cuBLAS bmm -- what
    the reference itself would run here, in fp32 (torch defaults: TF32 allowed for cuDNN convolutions) and under torch.autocast(bfloat16).
    ms per denoise step (UNet forward + posterior) over `nsteps` steps after 3 warm-up steps, CUDA events."""
    import torch
    from dif_pan_b200 import synth
    from oracle import sampler_oracle as so, unet_oracle as uo

    torch.backends.cudnn.benchmark = True
    kw = synth.unet_kwargs("wv3")
    cl = lambda v: v.to(dev).contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v.to(dev)
    sd = {k: cl(v) for k, v in synth.make_state_dict(0, **kw).items()}
    kw2 = dict(kw)
    kw2.pop("dropout")
    cfg = uo.UNetCfg(**kw2)
    sb = {k: v.to(dev) for k, v in so.schedule_buffers(so.make_beta_schedule("cosine", 500)).items()}
    rows = {}
    for B in batches:
        base = synth.make_batch("wv3", min(B, 16), seed=1)["cond"]
        cond = cl(base.repeat((B + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:B])
        g = torch.Generator(device=dev).manual_seed(0)
        for mode in ("fp32", "bf16_autocast"):
            x = cl(torch.randn(B, 8, 64, 64, generator=g, device=dev))
            nz = torch.randn(B, 8, 64, 64, generator=g, device=dev)

            def step(i, x):
                t = torch.full((B,), 499 - i, dtype=torch.long, device=dev)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode != "fp32"):
                    out = uo.unet_forward(sd, cfg, x, t, cond, x)
                return so.ddpm_step(sb, x, t, out.float(), cond[:, :8], nz)

            for i in range(3):
                x = step(i, x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(nsteps):
                x = step(3 + i, x)
            e1.record()
            torch.cuda.synchronize()
            rows[f"B{B}_{mode}_ms_per_step"] = e0.elapsed_time(e1) / nsteps
        del cond
        torch.cuda.empty_cache()
    return rows


def other_configs_run(dev):
    """BASELINE configs[0], [2], [3] through the public API (parity cases, not bench lines): ms per sampling, best of 3 after a warm-up."""
    import torch
    import dif_pan_b200 as dp
    from dif_pan_b200 import synth

    def timed(fn, warm=1, reps=3):
        for _ in range(warm):
            fn()
        best = 1e30
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    def net_for(ds):
        kw = synth.unet_kwargs(ds)
        net = dp.UNetSR3(**kw)
        net.load_state_dict(synth.make_state_dict(0, **kw))
        return net.to(dev).eval()

    out = {}
    net = net_for("wv3")
    cond = synth.make_batch("wv3", 1, seed=1)["cond"].to(dev)
    x_T = torch.randn(1, 8, 64, 64, device=dev)
    ms = timed(lambda: dp.sample_cond(net, cond, 8, "dpm20", x_T=x_T), warm=2)
    out["configs[0] WV3 B=1 DPM-Solver++ 2M-20"] = dict(ms_per_sampling=ms, patches_per_s=1e3 / ms, ms_per_denoise_step=ms / 20)
    try:  # the same configuration on the host CPU (BASELINE configs[0] is the reference's CPU-runnable case)
        c0 = cpu_config0_run()
        out["configs[0] WV3 B=1 DPM-Solver++ 2M-20"].update(cpu_reference_ms_per_sampling=c0["cpu_ms_per_sampling"], cpu_cores=c0["cores"],
                                                            cpu_kind=c0["kind"], speedup_vs_cpu=c0["cpu_ms_per_sampling"] / ms)
    except Exception as e:
        out["configs[0] WV3 B=1 DPM-Solver++ 2M-20"]["cpu_error"] = repr(e)[:200]
    torch.set_grad_enabled(False)
    del net
    net = net_for("gf2")
    d = synth.make_batch("gf2", 1, size=512, seed=2)
    lms, pan = d["lms_dn"].float().to(dev), d["pan_dn"].float().to(dev)
    ms = timed(lambda: dp.fuse_scene(net, lms, pan, synth.DATASETS["gf2"].division, sampler="dpm25", patch=64, tile_batch=64))
    out["configs[2] GF2 512x512 scene -> 64 tiles, DPM-Solver++ 2M-25, stitched"] = dict(ms_per_scene=ms, patches_per_s=64e3 / ms, ms_per_denoise_step=ms / 25)
    del net
    net = net_for("cave")
    cond = synth.make_batch("cave", 8, seed=3)["cond"].repeat(16, 1, 1, 1).contiguous().to(dev)
    ms = timed(lambda: dp.sample_cond(net, cond, 31, "ddim25"))
    out["configs[3] CAVE B=128 DDIM-25"] = dict(ms_per_sampling=ms, patches_per_s=128e3 / ms, ms_per_denoise_step=ms / 25)
    del net, cond
    torch.cuda.empty_cache()
    return out


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(a.cpu_batch, a.timesteps, max(a.steps, 1), a.warmup)
        line = dict(impl="reference", metric=METRIC, value=r["value"], unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                    ms_per_step=r["ms_per_denoise_step"] * a.timesteps * (a.batch / a.cpu_batch), higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload="BASELINE configs[1]: WV3 64x64x8 patches, batch %d per GPU, full DDPM loop cosine T=%d, clip (0,1), "
                                         "x_start, self-cond" % (a.batch, a.timesteps), batch_per_gpu=a.batch, timesteps=a.timesteps,
                                sample="reference algorithm (oracle port, torch CPU fp32) on the host cores: one denoise step of %d patches per "
                                       "bench step, extrapolated to the full workload" % a.cpu_batch),
                    cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"]),
                    e2e=dict(value=r["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    import dif_pan_b200 as dp
    from dif_pan_b200 import synth
    from dif_pan_b200.sharding import gather_patches, shard_range

    torch.set_grad_enabled(False)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T = a.timesteps
    kw = synth.unet_kwargs("wv3")
    net = dp.UNetSR3(**kw)
    net.load_state_dict(synth.make_state_dict(0, **kw))
    net = net.to(dev).eval()
    dif = dp.GaussianDiffusion(net, image_size=64, channels=8, pred_mode="x_start", loss_type="l1", device=dev, clamp_range=(0, 1))
    dif.set_new_noise_schedule(betas=dp.make_beta_schedule("cosine", T), device=dev)
    dif = dif.to(dev)

    class Workload:
        """`total` patches over all ranks, this rank's contiguous shard [lo, hi) of them (sharding.shard_range)."""

        def __init__(self, total):
            self.total = total
            self.lo, self.hi = shard_range(total, rank, world)
            self.B = B = self.hi - self.lo
            # synthetic conditioning: 16 distinct image-like samples tiled to the shard (rank-dependent seed), pinned on the host
            base = synth.make_batch("wv3", min(B, 16), seed=1000 + rank)["cond"]
            self.cond_host = base.repeat((B + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:B].contiguous().pin_memory()
            self.out_host = torch.empty(B, 8, 64, 64, dtype=torch.float32).pin_memory()
            self.cond_dev = self.cond_host.to(dev)
            self.rt = net.runtime(B, 64, 64)

        def one_sampling(self, e2e: bool, seed: int):
            dif.seed = seed
            dif.noise_shard = (self.lo, self.total)  # in-kernel Philox noise indexed by the GLOBAL patch index: independent per patch on every rank
            if e2e:
                c = self.cond_host.to(dev, non_blocking=True)  # fresh device tensor -> cond cache rebuilt, like a new batch
            else:
                c = self.cond_dev
                self.rt.cond_key = None  # force the cond-cache build: it is part of every sampling
            s = dif(c, mode="ddpm_sample")
            sr = dp.fuse_output(s, c)
            if world > 1:
                sr = gather_patches(sr, self.total)[self.lo:self.hi]  # final NCCL all_gather of the finished patches
            if e2e:
                self.out_host.copy_(sr, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            return sr

        def timed(self, e2e: bool, k: int):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(k):
                self.one_sampling(e2e, 100 + i)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                dist.barrier()
            return float(ms)

        def measure(self, steps, warmup):
            for i in range(warmup):
                self.one_sampling(i % 2 == 1, i)
            torch.cuda.synchronize()
            return self.timed(False, steps), self.timed(True, steps)

    strong_head = a.scaling == "strong"
    if strong_head and a.batch % world:
        raise SystemExit("--scaling strong needs --batch divisible by the number of GPUs")
    wl = Workload(a.batch if strong_head else a.batch * world)
    B = wl.B
    rt = wl.rt
    for i in range(a.warmup):
        wl.one_sampling(i % 2 == 1, i)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = wl.timed(False, a.steps)
    ms_e2e = wl.timed(True, a.steps)
    clocks = sampler.stop() if rank == 0 else None

    # roofline of the dominant kernel (conv3x3_halo_tc_kernel: every 3x3 stride-1 conv of the UNet): CUDA events around each
    # launch of one eager replay of the denoise-step plan, on the stream the kernels are launched on
    roof = None
    if rank == 0:
        pk = peaks()
        rows = rt.sch.fwd.profile(rt.stream)
        rows = rt.sch.fwd.profile(rt.stream)
        var = rt.sch.fwd.variants()
        ops = rt.sch.fwd.ops
        tot_ms = sum(r[2] for r in rows)
        names = {0: "conv_igemm_tc_kernel", 2: "conv3x3_halo_tc_kernel", 3: "cs_gemm_tc_kernel"}
        per = {}
        for i, r in enumerate(rows):
            if var[i] >= 0:
                d = per.setdefault(names[var[i]], dict(ms=0.0, ref_flops=0.0, executed_flops=0.0, bytes=0.0, launches=0))
                d["ms"] += r[2]; d["ref_flops"] += ops[i].ref_flops; d["executed_flops"] += r[3]; d["bytes"] += r[4]; d["launches"] += 1
        # Physical floor of every halo-conv launch: max(algorithmic bytes / HBM peak, MMA issue time).  One M128 x N x K16 tcgen05.mma cannot
        # issue faster than max(44.8, N / 2) cycles (profiles/r01_microbench_umma_rate.txt), so with this network's N = 32 / 64 the tensor pipe
        # tops out at 36 % / 71 % of its peak whatever the kernel does; N is taken unsplit (a lower bound of the issue time).
        sm_hz = 1.965e9
        floor_ms = {"hbm": 0.0, "mma_issue": 0.0, "max": 0.0}
        for i, r in enumerate(rows):
            if var[i] != 2:
                continue
            f = ops[i].fields
            tiles = f["batch"] * ((f["out_h"] + 15) // 16) * ((f["out_w"] + 7) // 8)
            if f.get("dw_w"):
                k16 = sum(f["a_c"]) / 16.0 * 2
            else:
                k16 = sum(t * c for t, c in zip(f["taps"], f["a_c"])) / 16.0
            mma = tiles * k16 * max(44.8, f["n_pad"] / 2.0) / (148 * sm_hz) * 1e3
            hbm = r[4] / (pk["hbm_gbs"] * 1e9) * 1e3
            floor_ms["hbm"] += hbm; floor_ms["mma_issue"] += mma; floor_ms["max"] += max(hbm, mma)
        dom = max(per, key=lambda k: per[k]["ms"])
        d = per[dom]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(dom + "_dram_bytes_per_launch")
        ach = d["ref_flops"] / (d["ms"] * 1e-3) / 1e12
        gemm_ms = sum(v["ms"] for v in per.values())
        step_graph_ms = ms_dev / a.steps / T                      # one denoise step inside the sampling loop (CUDA graph + sampler kernel)
        in_graph = d["ms"] * step_graph_ms / tot_ms               # the kernel's share of the step applied to the in-graph step time
        ref_step = 8.378e9 * B
        roof = dict(bound="tensor", kernel=dom, achieved=ach, peak=pk["bf16_tflops_sustained"], unit="TFLOP/s",
                    frac=ach / pk["bf16_tflops_sustained"], traffic=traffic, peak_source=pk["source"] + ":bf16_tflops_sustained",
                    timing="eager replay of the step plan, one CUDA-event pair per launch (serialised: no overlap between launches)",
                    in_graph=dict(avg_launch_ms=in_graph / d["launches"], achieved=d["ref_flops"] / (in_graph * 1e-3) / 1e12,
                                  frac=d["ref_flops"] / (in_graph * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                                  how="share of the eager per-launch events x measured in-graph step time"),
                    launches_per_denoise_step=d["launches"], avg_launch_ms=d["ms"] / d["launches"], share_of_step=d["ms"] / tot_ms,
                    floor=dict(hbm_ms=floor_ms["hbm"], mma_issue_ms=floor_ms["mma_issue"], per_launch_max_ms=floor_ms["max"],
                               measured_ms=per.get("conv3x3_halo_tc_kernel", d)["ms"],
                               frac_of_floor=floor_ms["max"] / per.get("conv3x3_halo_tc_kernel", d)["ms"],
                               how="sum over the halo conv launches of max(algorithmic bytes / HBM peak, tiles x K16 steps x max(44.8, N/2) cycles "
                                   "/ 148 SMs at 1.965 GHz): the tcgen05.mma issue-rate floor at this network's small N"),
                    algorithmic_flops_per_launch=d["ref_flops"] / d["launches"], executed_flops_per_launch=d["executed_flops"] / d["launches"],
                    algorithmic_bytes_per_launch=d["bytes"] / d["launches"],
                    hbm=dict(achieved=d["bytes"] / (d["ms"] * 1e-3) / 1e9, peak=pk["hbm_gbs"], unit="GB/s",
                             frac=d["bytes"] / (d["ms"] * 1e-3) / 1e9 / pk["hbm_gbs"]),
                    all_conv_kernels=dict(share_of_step=gemm_ms / tot_ms, launches=sum(v["launches"] for v in per.values()),
                                          achieved=sum(v["ref_flops"] for v in per.values()) / (gemm_ms * 1e-3) / 1e12),
                    whole_step=dict(reference_flops=ref_step, ms_graph=step_graph_ms, achieved=ref_step / (step_graph_ms * 1e-3) / 1e12,
                                    frac=ref_step / (step_graph_ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"]),
                    unet_step_ms_events=tot_ms, launches_per_unet_forward=len(rows))
    launches_fwd, launches_cnd = rt.launches_per_step(), len(rt.sch.cnd)

    # BASELINE configs[1] as written: 256 patches in TOTAL, sharded 256 / N per GPU (strong scaling).  At N = 1 it IS the headline run.
    strong = None
    if not a.no_extras and not strong_head:
        if world == 1:
            strong = dict(patches_total=a.batch, patches_per_gpu=a.batch, value=a.batch * a.steps / (ms_dev * 1e-3), unit=UNIT,
                          e2e=a.batch * a.steps / (ms_e2e * 1e-3), ms_per_denoise_step=ms_dev / a.steps / T, note="N = 1: identical to the headline run")
        elif a.batch % world == 0:
            n1_value = a.batch * a.steps / (ms_dev * 1e-3)   # one GPU sampling 256 patches = this run's per-GPU (weak) figure
            ws = Workload(a.batch)
            sd_ms, se_ms = ws.measure(a.steps, a.warmup)
            if rank == 0:
                v = a.batch * a.steps / (sd_ms * 1e-3)
                strong = dict(patches_total=a.batch, patches_per_gpu=ws.B, value=v, unit=UNIT, e2e=a.batch * a.steps / (se_ms * 1e-3),
                              ms_per_denoise_step=sd_ms / a.steps / T, speedup_vs_n1=v / n1_value, efficiency_vs_n1=v / n1_value / world,
                              n1_value=n1_value, n1_how="per-GPU throughput of the weak run above (one GPU sampling %d patches)" % a.batch,
                              launches_per_denoise_step=ws.rt.launches_per_step() + 1)
    cpu = eager = others = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_reference_run(a.cpu_batch, T, 10, 2)
        cpu = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"], ms_per_denoise_step=r["ms_per_denoise_step"])
    if rank == 0 and world == 1 and not a.no_extras:
        step_ms = ms_dev / a.steps / T
        try:
            eager = gpu_eager_run(dev)
            eager["what"] = ("reference algorithm (oracle port) as eager PyTorch on this GPU: cuDNN / cuBLAS, channels-last, cudnn.benchmark; "
                             "fp32 (torch default TF32 policy) and torch.autocast(bfloat16); ms per denoise step (UNet forward + DDPM posterior), 20 steps")
            eager["speedup_vs_eager_bf16_B%d" % a.batch] = eager.get("B%d_bf16_autocast_ms_per_step" % a.batch, float("nan")) / step_ms
            eager["speedup_vs_eager_fp32_B%d" % a.batch] = eager.get("B%d_fp32_ms_per_step" % a.batch, float("nan")) / step_ms
        except Exception as e:  # the bench line must survive a failure of the side measurement
            eager = dict(error=repr(e)[:300])
        torch.cuda.empty_cache()
        try:
            others = other_configs_run(dev)
        except Exception as e:
            others = dict(error=repr(e)[:300])
    if rank == 0:
        total = wl.total * a.steps
        launches = a.steps * (T * (launches_fwd + 1) + launches_cnd + 2)
        line = dict(metric=METRIC, value=total / (ms_dev * 1e-3), unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup,
                    ms_per_step=ms_dev / a.steps, higher_is_better=True, scaling=a.scaling, vs_baseline=None, dtype="bf16", data="synthetic",
                    config=dict(workload="BASELINE configs[1]: WV3 64x64x8 patches, batch %d per GPU, full DDPM loop cosine T=%d, clip (0,1), "
                                         "x_start, self-cond" % (B, T), batch_per_gpu=B, timesteps=T, patches_total=wl.total,
                                l2_policy="inputs larger than L2 (per-step activation working set ~GBs >> 126 MB)",
                                noise="in-kernel Philox4x32-10, counter = (seed, step, global patch element)",
                                parity_note="Haar DWT oracle is pinned only by the two PyWavelets documentation known answers (pywt absent)"),
                    unet_ms_per_denoise_step=ms_dev / a.steps / T,
                    e2e=dict(value=total / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=wl.cond_host.numel() * 4, d2h_bytes_per_step=wl.out_host.numel() * 4,
                             ms_per_step=ms_e2e / a.steps),
                    gpu_launches=launches, clocks=clocks, roofline=roof, cpu_baseline=cpu, strong=strong, gpu_eager_baseline=eager,
                    other_configs=others)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
