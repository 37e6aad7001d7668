"""Data-parallel training plumbing ("training gradient all-reduce"): one process per GPU, replicated weights, every rank
back-propagates its own mini-batch, gradients are AVERAGED over the ranks before `clip_grad_norm_` / the optimiser step (the reference is
single-process: /root/reference/diffusion_engine.py:219-248; clipping must see the reduced gradients).

`GradientAllReducer` packs the gradients into a few flat fp32 buckets (the parameters' `.grad` are VIEWS into the buckets, so nothing is copied),
fills them in reverse registration order -- the order in which autograd finishes them -- and launches each bucket's all-reduce (NCCL over
NVLink / NVSwitch on GPUs, gloo in the CPU tests) asynchronously from a post-accumulate hook as soon as its last gradient is final, so the
41.6 MB of fp32 gradients travel while the rest of the backward pass still computes.  `finish()` waits for the outstanding buckets and divides by
the world size.  No collective exists on the inference path.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class GradientAllReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 8 << 20, group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.buckets: List[torch.Tensor] = []
        self._bucket_of = {}
        self._pending: List[int] = []
        self._expected: List[int] = []
        self._work = []
        # False: the hooks do nothing and finish() reduces every bucket after backward (no overlap) -- for CUDA-graph replay of the
        # backward pass (training.GraphedLossStep), where Python hooks do not run
        self.hooks_enabled = True
        cur, cur_bytes = [], 0
        groups = []
        for p in reversed(self.params):                       # gradients become final roughly back to front
            if p.dtype != torch.float32:
                raise ValueError("GradientAllReducer expects fp32 master parameters")
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                groups.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            groups.append(cur)
        for bi, ps in enumerate(groups):
            flat = torch.zeros(sum(p.numel() for p in ps), dtype=torch.float32, device=ps[0].device)
            off = 0
            for p in ps:
                p.grad = flat[off:off + p.numel()].view_as(p)  # autograd accumulates in place into this view
                off += p.numel()
                self._bucket_of[p] = bi
                p.register_post_accumulate_grad_hook(self._hook)
            self.buckets.append(flat)
            self._expected.append(len(ps))
        self._pending = list(self._expected)

    # -- per-iteration protocol: zero_grad() -> backward() (hooks fire) -> finish() -> clip / optimiser step -----------------------------
    def zero_grad(self) -> None:
        for flat in self.buckets:
            flat.zero_()
        self._pending = list(self._expected)
        self._work = []

    def _hook(self, p: torch.nn.Parameter) -> None:
        if not self.hooks_enabled:
            return
        bi = self._bucket_of[p]
        if p.grad.data_ptr() < self.buckets[bi].data_ptr() or p.grad.data_ptr() >= self.buckets[bi].data_ptr() + self.buckets[bi].numel() * 4:
            raise RuntimeError("a gradient was re-allocated outside its bucket (zero_grad(set_to_none=True)?): use GradientAllReducer.zero_grad()")
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and self.world > 1:
            self._work.append(dist.all_reduce(self.buckets[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> None:
        """Wait for the bucket all-reduces and turn the sums into means.  Parameters that received no gradient this iteration keep
        zeros (their bucket is reduced here, synchronously, so that every rank issues the same collectives)."""
        if self.world > 1:
            for bi, left in enumerate(self._pending):
                if left > 0:
                    self._work.append(dist.all_reduce(self.buckets[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
                    self._pending[bi] = 0
            for w in self._work:
                w.wait()
            for flat in self.buckets:
                flat.div_(self.world)
        self._work = []

    @property
    def nbytes(self) -> int:
        return sum(b.numel() * 4 for b in self.buckets)
