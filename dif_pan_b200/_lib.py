"""ctypes binding of libddif_b200.so (the C ABI in include/ddif_b200.h).

The parameter structs are mirrored by PARSING the header, so the Python side cannot drift from the C side.
There is no CPU fallback: if the shared library is missing or an op fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
HEADER = os.path.join(_ROOT, "include", "ddif_b200.h")
LIB_PATH = os.environ.get("DDIF_LIB", os.path.join(_HERE, "libddif_b200.so"))  # DDIF_LIB: tuning variants (tools/)


def _parse_header(path: str):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    kinds = {}
    m = re.search(r"enum ddif_op_kind \{(.*?)\};", src, flags=re.S)
    for name, val in re.findall(r"(DDIF_OP_\w+)\s*=\s*(\d+)", m.group(1)):
        kinds[name] = int(val)
    structs: Dict[str, type] = {}
    for body, name in re.findall(r"typedef struct \{(.*?)\}\s*(\w+);", src, flags=re.S):
        fields = []
        for stmt in body.split(";"):
            stmt = " ".join(stmt.split())
            if not stmt:
                continue
            mm = re.match(r"^(?:const )?(\w+)\s*(\*?)\s*(.*)$", stmt)
            base, star, rest = mm.group(1), mm.group(2), mm.group(3)
            for decl in rest.split(","):
                decl = decl.strip()
                arr = re.match(r"^(\w+)\[(\d+)\]$", decl)
                fname, count = (arr.group(1), int(arr.group(2))) if arr else (decl, 1)
                if star:
                    ct = ctypes.c_void_p
                elif base == "int64_t":
                    ct = ctypes.c_int64
                elif base == "double":
                    ct = ctypes.c_double
                else:
                    raise ValueError(f"unsupported field type in {name}: {stmt}")
                fields.append((fname, ct * count if count > 1 else ct))
        structs[name] = type(name, (ctypes.Structure,), {"_fields_": fields})
    protos = re.findall(r"^\s*(?:int|const char\*|ddif_plan_t\*|void)\s+(ddif_\w+)\(", src, flags=re.M)
    return kinds, structs, protos


KINDS, STRUCTS, EXPORTS = _parse_header(HEADER)

KIND_OF_STRUCT = {
    "ddif_gemm_t": "DDIF_OP_GEMM", "ddif_in_convert_t": "DDIF_OP_IN_CONVERT", "ddif_time_embed_t": "DDIF_OP_TIME_EMBED",
    "ddif_gn_apply_t": "DDIF_OP_GN_APPLY", "ddif_softmax_h_t": "DDIF_OP_SOFTMAX_H", "ddif_attn_t": "DDIF_OP_ATTN",
    "ddif_upsample2x_t": "DDIF_OP_UPSAMPLE2X", "ddif_conv_direct_t": "DDIF_OP_CONV_DIRECT", "ddif_stats_t": "DDIF_OP_STATS",
    "ddif_memset_t": "DDIF_OP_MEMSET", "ddif_resize_t": "DDIF_OP_RESIZE", "ddif_fwm_context_t": "DDIF_OP_FWM_CONTEXT",
    "ddif_fwm_weff_t": "DDIF_OP_FWM_WEFF", "ddif_ddpm_step_t": "DDIF_OP_DDPM_STEP", "ddif_ddim_step_t": "DDIF_OP_DDIM_STEP",
    "ddif_dpmpp_step_t": "DDIF_OP_DPMPP_STEP", "ddif_q_sample_t": "DDIF_OP_Q_SAMPLE", "ddif_cond_assemble_t": "DDIF_OP_COND_ASSEMBLE",
    "ddif_randn_t": "DDIF_OP_RANDN", "ddif_axpby_clip_t": "DDIF_OP_AXPBY_CLIP", "ddif_dpm_single_t": "DDIF_OP_DPM_SINGLE",
    "ddif_loss_t": "DDIF_OP_LOSS", "ddif_dpm_err_t": "DDIF_OP_DPM_ERR", "ddif_attn_block_t": "DDIF_OP_ATTN_BLOCK", "ddif_multi_tensor_t": "DDIF_OP_MULTI_TENSOR", "ddif_axpby_t": "DDIF_OP_AXPBY", "ddif_metrics_t": "DDIF_OP_METRICS", "ddif_tile_t": "DDIF_OP_TILE",
    "ddif_wavelet_cond_t": "DDIF_OP_WAVELET_COND", "ddif_wgrad_t": "DDIF_OP_WGRAD", "ddif_colsum_t": "DDIF_OP_COLSUM",
    "ddif_fwm_front_t": "DDIF_OP_FWM_FRONT",
}

_lib = None


class Stream(int):
    """A cudaStream_t handle that remembers its device: every launch through this module makes that device current for the
    call (kernel launches, cudaFuncSetAttribute and graph capture act on the CUDA *current* device, which need not be the device
    of the tensors when the caller passes device='cuda:N' like the reference does)."""

    def __new__(cls, handle: int, device=None):
        obj = int.__new__(cls, handle)
        obj.device = device
        return obj


def stream_of(device) -> "Stream":
    """Current torch stream of `device` as a device-carrying handle."""
    import torch
    device = torch.device(device)
    return Stream(torch.cuda.current_stream(device).cuda_stream, device)


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NOGUARD = _NoGuard()


def guard(stream):
    """Context manager making the stream's device current (no-op for plain ints and when it already is)."""
    dev = getattr(stream, "device", None)
    if dev is None:
        return _NOGUARD
    import torch
    if dev.index is None or torch.cuda.current_device() == dev.index:
        return _NOGUARD
    return torch.cuda.device(dev)


def load() -> ctypes.CDLL:
    """Load the CUDA library; raise loudly if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). dif_pan_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.ddif_version.restype = ctypes.c_int
    lib.ddif_error_string.restype = ctypes.c_char_p
    lib.ddif_error_string.argtypes = [ctypes.c_int]
    lib.ddif_launch.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.ddif_plan_create.restype = ctypes.c_void_p
    lib.ddif_plan_destroy.argtypes = [ctypes.c_void_p]
    lib.ddif_plan_destroy.restype = None
    lib.ddif_plan_add.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    lib.ddif_plan_size.argtypes = [ctypes.c_void_p]
    lib.ddif_plan_launches.argtypes = [ctypes.c_void_p]
    lib.ddif_plan_op_variant.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.ddif_plan_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.ddif_plan_graph_build.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.ddif_plan_graph_launch.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.ddif_plan_set_side_branch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.ddif_plan_profile.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    for name in ("ddif_haar_dwt2_f32", "ddif_haar_idwt2_f32", "ddif_cond_assemble_f32", "ddif_ddpm_step_f32",
                 "ddif_ddim_step_f32", "ddif_dpmpp_step_f32", "ddif_q_sample_f32", "ddif_conv_igemm_bf16", "ddif_dpm_single_f32", "ddif_loss_f32", "ddif_dpm_err_f32", "ddif_multi_tensor_f32",
                 "ddif_metrics_f32", "ddif_tile_f32", "ddif_wavelet_cond_f32", "ddif_wgrad_bf16", "ddif_colsum_bf16"):
        getattr(lib, name).argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    _lib = lib
    return lib


def check(rc: int, what: str = "ddif") -> int:
    if rc < 0 or (rc > 0 and what != "index"):
        msg = load().ddif_error_string(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")
    return rc


def make(struct_name: str, **fields):
    """Instantiate a parameter struct; tensors may be passed as data_ptr() ints, None = NULL."""
    st = STRUCTS[struct_name]()
    valid = {f[0] for f in st._fields_}
    for k, v in fields.items():
        if k not in valid:
            raise KeyError(f"{struct_name} has no field {k}")
        if isinstance(v, (list, tuple)):
            arr = getattr(st, k)
            for i, e in enumerate(v):
                arr[i] = 0 if e is None else e
        else:
            setattr(st, k, v)
    return st


def launch(struct_name: str, stream: int, kind: str = None, **fields) -> None:
    """Launch one op immediately on the given CUDA stream handle (`kind` only for structs shared by two ops)."""
    st = make(struct_name, **fields)
    k = KINDS[kind or KIND_OF_STRUCT[struct_name]]
    with guard(stream):
        check(load().ddif_launch(k, ctypes.byref(st), ctypes.c_void_p(stream)), struct_name)


def launch_kind(kind_name: str, st, stream: int) -> None:
    with guard(stream):
        check(load().ddif_launch(KINDS[kind_name], ctypes.byref(st), ctypes.c_void_p(stream)), kind_name)
