#!/usr/bin/env python
"""BASELINE configs[4]: one DDIF training step (p_losses -> backward -> gradient all-reduce -> clip 0.003 -> AdamW -> EMA) at batch 32 per GPU.

  python tools/train_bench.py [--batch 32] [--steps 10] [--warmup 3] [--pred-mode x_start|noise]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py ...      (one rank per GPU, NCCL)

Prints one JSON line (rank 0): ms per step (max over ranks, CUDA events), patches/s over all ranks, and the split forward+loss / backward
(+ overlapped all-reduce) / reduce-wait + clip + AdamW + EMA.  First slice of the training path: convolutions on this repo's kernels,
the ops between them through torch autograd (dif_pan_b200/training.py) -- this is a functional measurement, not a tuned one."""
import argparse
import copy
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dif_pan_b200 as dp  # noqa: E402
from dif_pan_b200 import synth  # noqa: E402
from dif_pan_b200.ddp import GradientAllReducer  # noqa: E402
from dif_pan_b200.optim import EmaUpdater, FusedAdamW, grad_clip  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pred-mode", default="x_start", choices=["x_start", "noise"])
    ap.add_argument("--graph", action="store_true", help="replay the forward + backward from CUDA graphs (training.GraphedLossStep)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    kw = synth.unet_kwargs("wv3")
    net = dp.UNetSR3(**kw)
    net.load_state_dict(synth.make_state_dict(0, **kw))
    net = net.to(dev).train()
    ema_net = copy.deepcopy(net).eval()
    loss_type = "l1" if a.pred_mode == "x_start" else "l2"
    mk = lambda m: dp.GaussianDiffusion(m, image_size=64, channels=8, pred_mode=a.pred_mode, loss_type=loss_type, device=dev, clamp_range=(0, 1))
    dif, dif_ema = mk(net), mk(ema_net)
    dif.set_new_noise_schedule(betas=dp.make_beta_schedule("cosine", 500), device=dev)
    red = GradientAllReducer(net.parameters())
    opt = FusedAdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)
    ema = EmaUpdater(dif, dif_ema, decay=0.995, start_iter=0)
    d = synth.make_batch("wv3", min(a.batch, 16), seed=100 + rank)
    rep = lambda t: t.repeat((a.batch + t.shape[0] - 1) // t.shape[0], 1, 1, 1)[:a.batch].contiguous().to(dev)
    x0, cond = rep(d["hr"] - d["lms"]), rep(d["cond"])
    ev = lambda: torch.cuda.Event(enable_timing=True)
    split = [0.0, 0.0, 0.0]
    graphed = None
    if a.graph:
        from dif_pan_b200.training import GraphedLossStep
        red.hooks_enabled = False  # hooks do not run at graph replay: finish() reduces all buckets after the backward graph
        graphed = GraphedLossStep(dif, x0, cond)

    def step(it, timed):
        e = [ev() for _ in range(4)]
        red.zero_grad()
        e[0].record()
        if graphed is not None:
            loss = graphed.run(x0, cond)
            e[1].record()
        else:
            with torch.enable_grad():
                loss, _ = dif(x0, cond=cond)
            e[1].record()
            loss.backward()
        e[2].record()
        red.finish()
        grad_clip(red.params, mode="norm", value=0.003)
        opt.step()
        ema.update(it + 1)
        e[3].record()
        if timed:
            torch.cuda.synchronize()
            for k in range(3):
                split[k] += e[k].elapsed_time(e[k + 1])
        return loss

    for i in range(a.warmup):
        step(i, False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = ev(), ev()
    t0.record()
    for i in range(a.steps):
        loss = step(a.warmup + i, True)
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / a.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps(dict(config="BASELINE configs[4]: DDIF training step, WV3 64x64, batch %d per GPU, pred_mode %s + %s, self-cond p=0.5, "
                                     "AdamW lr 1e-4 wd 1e-4, grad clip 0.003, EMA 0.995" % (a.batch, a.pred_mode, loss_type), n_gpus=world, batch_per_gpu=a.batch,
                              ms_per_step=float(ms), patches_per_s=a.batch * world / float(ms) * 1e3, loss=float(loss),
                              split_ms=dict(forward_and_loss=split[0] / a.steps, backward_with_overlapped_allreduce=split[1] / a.steps,
                                            reduce_wait_clip_adamw_ema=split[2] / a.steps),
                              grad_bytes_allreduced=red.nbytes if world > 1 else 0, buckets=len(red.buckets), cuda_graph=bool(a.graph),
                              note="first slice: dense convolutions (fwd / dgrad / wgrad) on the repo's CUDA kernels, ops between them via torch autograd; "
                                   "host-bound at this batch size")))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
