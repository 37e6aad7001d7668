"""Training-loop surround as fused multi-tensor kernels: `EmaUpdater` (drop-in for
/root/reference/utils/optim_utils.py:24-85), `grad_clip` (/root/reference/utils/misc.py:25-36) and `FusedAdamW` (the `torch.optim.AdamW`
the reference engine steps, diffusion_engine.py:202-241).  The reference touches ~350 parameter tensors with 3-10 ATen kernels each per
iteration; here every operation is ONE launch over a device-side table of (pointer, size) entries (`ddif_multi_tensor_f32`).
The UNet's backward is not implemented yet, so these run on whatever `.grad` the caller provides (tests use synthetic gradients);
checkpoints need nothing special: `UNetSR3.state_dict()` has the reference's 702 names / shapes.  CUDA only, fp32 parameters.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch

from . import _lib

_CHUNK = 65536


class _TensorList:
    """Device-side table for a fixed list of up-to-4-tensor groups (all fp32, contiguous, same numel within a group)."""

    def __init__(self, groups: Sequence[Sequence[Optional[torch.Tensor]]], written: Optional[Sequence[torch.Tensor]] = None):
        """`written`: the tensors (normally nn.Parameters) whose storage the p0 column aliases.  The kernels write through raw device
        pointers, which autograd's version counters cannot see; `run` bumps the counters of these tensors after every op that
        writes p0, so that caches keyed on (data_ptr, _version) -- UNetSR3's packed bf16 weights and its captured CUDA graph --
        notice the update (an EMA / AdamW step followed by sampling is the reference's validation flow)."""
        groups = [list(g) + [None] * (4 - len(g)) for g in groups]
        self.written = list(written) if written is not None else []
        if not groups:
            raise ValueError("empty tensor list")
        dev = groups[0][0].device
        if dev.type != "cuda":
            raise RuntimeError("dif_pan_b200.optim runs on CUDA only (no CPU fallback)")
        ptrs, sizes, chunks = [], [], []
        for i, g in enumerate(groups):
            n = g[0].numel()
            for t in g:
                if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n or t.device != dev):
                    raise ValueError("multi-tensor groups need contiguous fp32 CUDA tensors of equal size on one device")
            ptrs.append([0 if t is None else t.data_ptr() for t in g])
            sizes.append(n)
            chunks += [[i, s] for s in range(0, n, _CHUNK)]
        self.keep = groups
        self.dev = dev
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self.sizes = torch.tensor(sizes, dtype=torch.int64, device=dev)
        self.chunks = torch.tensor(chunks, dtype=torch.int64, device=dev).reshape(-1, 2)
        self.key = tuple(tuple(p) for p in ptrs)

    def run(self, op: int, out: Optional[torch.Tensor] = None, **s) -> None:
        _lib.launch("ddif_multi_tensor_t", _lib.stream_of(self.dev), ptrs=self.ptrs.data_ptr(), sizes=self.sizes.data_ptr(),
                    chunks=self.chunks.data_ptr(), out=out.data_ptr() if out is not None else None, nchunks=self.chunks.shape[0], chunk=_CHUNK,
                    op=op, **{k: float(v) for k, v in s.items()})
        if op != 2:  # every op but SUMSQ writes p0
            if self.written:
                torch._C._increment_version(self.written)


def _table(cache: dict, groups, written=None) -> _TensorList:
    """Rebuild the device table only when a tensor of the list moved (parameters / optimizer state normally never do)."""
    key = tuple(tuple([0 if t is None else t.data_ptr() for t in g] + [0] * (4 - len(g))) for g in groups)
    tl = cache.get("tl")
    if tl is None or tl.key != key:
        tl = _TensorList(groups, written)
        cache["tl"] = tl
    return tl


class EmaUpdater:
    """Same constructor / methods as the reference (`model` and `ema_model` expose `.model.parameters()`, i.e. GaussianDiffusion objects)."""

    def __init__(self, model, ema_model, decay=0.9999, start_iter=0) -> None:
        self.model, self.ema_model, self.decay, self.start_iter = model, ema_model, decay, start_iter
        self.iteration = start_iter
        self._cache: dict = {}

    @torch.no_grad()
    def update(self, iteration):
        self.iteration = iteration
        pairs = list(zip(self.model.model.parameters(), self.ema_model.model.parameters()))
        groups = [(pe.data, p.data) for p, pe in pairs]
        tl = _table(self._cache, groups, written=[pe for _, pe in pairs])
        if iteration > self.start_iter:
            tl.run(0, s0=self.decay, s1=1 - self.decay)   # p_ema = p_ema * decay + p * (1 - decay)
        else:
            tl.run(1)                                      # p_ema = p

    def load_ema_params(self):
        self.model.load_state_dict(self.ema_model.state_dict())

    def load_model_params(self):
        self.ema_model.load_state_dict(self.model.state_dict())

    @property
    def on_fly_model_state_dict(self):
        return (self.model.module.model if hasattr(self.model, "module") else self.model.model).state_dict()

    @property
    def ema_model_state_dict(self):
        return (self.ema_model.module.model if hasattr(self.model, "module") else self.ema_model.model).state_dict()


@torch.no_grad()
def grad_clip(params: Iterable[torch.nn.Parameter], mode: str = "value", value: float = None, **kwargs):
    """utils/misc.py:25-36.  mode 'norm' returns the total gradient 2-norm like `clip_grad_norm_`."""
    assert mode in ["value", "norm"], "mode should be @value or @norm"
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    tl = _TensorList([(g,) for g in grads])
    if mode == "value":
        tl.run(4, s0=value)
        return None
    acc = torch.zeros(1, dtype=torch.float64, device=grads[0].device)
    tl.run(2, out=acc)
    total = torch.sqrt(acc).to(torch.float32).reshape(())
    coef = torch.clamp(value / (total + 1e-6), max=1.0)      # clip_grad_norm_'s own fp32 expression (torch/nn/utils/clip_grad.py)
    tl.run(3, s0=float(coef))
    return total


class FusedAdamW(torch.optim.Optimizer):
    """`torch.optim.AdamW(params, lr, betas, eps, weight_decay)` semantics (no amsgrad / maximize), ONE launch per parameter group and
    step.  A real `torch.optim.Optimizer`: `param_groups` (read on every step, so `MultiStepLR` and the schedulers of
    /root/reference/utils/lr_scheduler.py attach unchanged, diffusion_engine.py:207-210,241), `state_dict()` / `load_state_dict()` with
    AdamW's own state layout (`step`, `exp_avg`, `exp_avg_sq` per parameter), so optimizer checkpoints round-trip with the reference's."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._cache: dict = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            buckets = {}  # parameters of a group normally share one step count -> one launch; stragglers get their own
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p.data, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p.data, memory_format=torch.contiguous_format)
                st["step"] += 1
                buckets.setdefault(int(st["step"]), []).append(p)
            b1, b2 = group["betas"]
            for k, ps in buckets.items():
                groups = [(p.data, p.grad, self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"]) for p in ps]
                tl = _table(self._cache.setdefault((gi, len(ps)), {}), groups, written=ps)
                tl.run(5, s0=group["lr"], s1=b1, s2=b2, s3=group["eps"], s4=group["weight_decay"], s5=1 - b1 ** k, s6=1 - b2 ** k)
        return loss
