"""Host-side schedule builder: turns a sequence of symbolic ops over named activation buffers into
(a) a liveness-packed device workspace and (b) a recorded C plan (`ddif_plan_*`, include/ddif_b200.h).

Pure Python / no CUDA needed until `finalize()`, so the schedule logic is unit-tested on CPU.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from . import _lib

ALIGN = 1024


@dataclass(eq=False)
class Buf:
    """A device buffer inside the plan workspace.  `persistent` buffers live for the whole plan."""

    name: str
    nbytes: int
    persistent: bool = False
    first: int = -1
    last: int = -1
    offset: int = -1

    def touch(self, op_index: int) -> None:
        if self.first < 0:
            self.first = op_index
        self.last = max(self.last, op_index)


@dataclass
class SymOp:
    struct: str
    fields: dict
    kind: Optional[str] = None
    label: str = ""
    flops: float = 0.0
    bytes: float = 0.0
    ref_flops: float = -1.0  # FLOPs of the reference ops this launch replaces (== flops unless the launch restructures them)


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


def pack_lifetimes(bufs: Sequence[Buf]) -> int:
    """Greedy offset assignment: buffers whose [first,last] op ranges overlap never share addresses.
    Returns the arena size in bytes."""
    placed: List[Buf] = []
    total = 0
    order = sorted(bufs, key=lambda b: (-b.nbytes, b.first))
    for b in order:
        size = _round_up(max(b.nbytes, 1), ALIGN)
        conflicts = sorted(
            ((p.offset, p.offset + _round_up(max(p.nbytes, 1), ALIGN)) for p in placed
             if b.persistent or p.persistent or not (p.last < b.first or b.last < p.first)),
            key=lambda t: t[0])
        off = 0
        for lo, hi in conflicts:
            if off + size <= lo:
                break
            off = max(off, hi)
        b.offset = off
        placed.append(b)
        total = max(total, off + size)
    return total


class PlanBuilder:
    """Collects symbolic ops; `finalize(base_ptr)` resolves buffers to addresses and records the C plan."""

    def __init__(self) -> None:
        self.ops: List[SymOp] = []
        self.bufs: List[Buf] = []
        self.arena_bytes = 0
        self.handle = None
        self._keep = []  # torch tensors referenced by raw pointer

    # -- buffers ------------------------------------------------------------------------------------------
    def buf(self, name: str, nbytes: int, persistent: bool = False) -> Buf:
        b = Buf(name, int(nbytes), persistent)
        if persistent:
            b.first, b.last = 0, 1 << 60
        self.bufs.append(b)
        return b

    def keep(self, tensor):
        """Hold a reference to an external torch tensor and return its device address."""
        self._keep.append(tensor)
        return tensor.data_ptr()

    # -- ops ----------------------------------------------------------------------------------------------
    def add(self, struct: str, kind: Optional[str] = None, label: str = "", flops: float = 0.0, traffic: float = 0.0, ref_flops: float = -1.0,
            **fields) -> int:
        idx = len(self.ops)

        def visit(v):
            if isinstance(v, Buf):
                v.touch(idx)
            elif isinstance(v, tuple) and len(v) == 2 and isinstance(v[0], Buf):
                v[0].touch(idx)
            elif isinstance(v, (list, tuple)):
                for e in v:
                    visit(e)

        for v in fields.values():
            visit(v)
        self.ops.append(SymOp(struct, fields, kind, label, flops, traffic, flops if ref_flops < 0 else ref_flops))
        return idx

    def layout(self) -> int:
        self.arena_bytes = pack_lifetimes(self.bufs)
        return self.arena_bytes

    def _resolve(self, v, base: int):
        if isinstance(v, Buf):
            return base + v.offset
        if isinstance(v, tuple) and len(v) == 2 and isinstance(v[0], Buf):
            return base + v[0].offset + int(v[1])
        if isinstance(v, (list, tuple)):
            return [self._resolve(e, base) for e in v]
        return v

    def finalize(self, base_ptr: int, device=None) -> None:
        """Record all ops into a C plan (GEMM tensor maps are encoded here; needs the CUDA driver).  `device`: the plan's device
        (made current while recording: tensor maps and kernel attributes belong to the current device)."""
        lib = _lib.load()
        with _lib.guard(_lib.Stream(0, device)):
            self._finalize(lib, base_ptr)

    def _finalize(self, lib, base_ptr: int) -> None:
        if self.arena_bytes == 0 and self.bufs:
            self.layout()
        self.handle = ctypes.c_void_p(lib.ddif_plan_create())
        for i, op in enumerate(self.ops):
            st = _lib.make(op.struct, **{k: self._resolve(v, base_ptr) for k, v in op.fields.items()})
            kind = _lib.KINDS[op.kind or _lib.KIND_OF_STRUCT[op.struct]]
            rc = lib.ddif_plan_add(self.handle, kind, ctypes.byref(st))
            if rc < 0:
                raise RuntimeError(f"ddif_plan_add failed for op {i} ({op.label or op.struct}): "
                                   f"{lib.ddif_error_string(rc).decode()} (code {rc})")

    def address(self, b: Buf, base_ptr: int) -> int:
        return base_ptr + b.offset

    # -- execution ----------------------------------------------------------------------------------------
    def run(self, stream: int, first: int = 0, last: int = -1) -> None:
        with _lib.guard(stream):
            _lib.check(_lib.load().ddif_plan_run(self.handle, first, last, ctypes.c_void_p(stream)), "ddif_plan_run")

    def set_side_branch(self, first: int, last: int, join_before: int) -> None:
        """Ops [first, last) become a side branch of the captured graph, joined before op `join_before` (ddif_plan_set_side_branch)."""
        _lib.check(_lib.load().ddif_plan_set_side_branch(self.handle, first, last, join_before), "ddif_plan_set_side_branch")

    def graph_build(self, stream: int) -> None:
        with _lib.guard(stream):
            _lib.check(_lib.load().ddif_plan_graph_build(self.handle, ctypes.c_void_p(stream)), "ddif_plan_graph_build")

    def graph_launch(self, stream: int) -> None:
        with _lib.guard(stream):
            _lib.check(_lib.load().ddif_plan_graph_launch(self.handle, ctypes.c_void_p(stream)), "ddif_plan_graph_launch")

    def profile(self, stream: int) -> List[Tuple[str, str, float, float, float]]:
        """[(label, struct, ms, flops, bytes)] per op, timed with CUDA events around each launch."""
        n = len(self.ops)
        ms = (ctypes.c_float * n)()
        kinds = (ctypes.c_int * n)()
        with _lib.guard(stream):
            _lib.check(_lib.load().ddif_plan_profile(self.handle, ctypes.c_void_p(stream), ms, kinds, n), "ddif_plan_profile")
        return [(op.label, op.struct, float(ms[i]), op.flops, op.bytes) for i, op in enumerate(self.ops)]

    def variants(self) -> List[int]:
        """Kernel variant of every op (GEMM ops: 0 generic TMA, 2 halo 3x3, 3 column-softmax GEMM; others -1)."""
        lib = _lib.load()
        return [int(lib.ddif_plan_op_variant(self.handle, i)) for i in range(len(self.ops))]

    def __len__(self) -> int:
        return len(self.ops)

    def close(self) -> None:
        if self.handle is not None:
            _lib.load().ddif_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
