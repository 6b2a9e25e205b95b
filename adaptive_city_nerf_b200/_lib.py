"""ctypes binding of libacn_b200.so (C ABI in include/acn_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing or a
tensor is not on a CUDA device, the call raises.  PyTorch is used for device memory, streams
and autograd plumbing only.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libacn_b200.so"
DEBUG_LIB_PATH = _PKG / "libacn_b200_debug.so"

F32, F16 = 0, 1
INTERP = {"Nearest": 0, "Linear": 1, "Smoothstep": 2}

_lock = threading.Lock()
_lib = None
_ctx = {}
_dlib = None
_dctx = {}

c_i64, c_int, c_f, c_p = C.c_int64, C.c_int, C.c_float, C.c_void_p


class FieldWeights(C.Structure):
    _fields_ = [("p", c_p * 14)]


class AdamTensor(C.Structure):
    """acn_adam_tensor (include/acn_b200.h)"""
    _fields_ = [("p", c_p), ("g", c_p), ("m", c_p), ("v", c_p), ("n", c_i64), ("lr", C.c_double),
                ("weight_decay", C.c_double), ("step", c_p), ("bias", c_p)]


ADAM_MAX_TENSORS = 192
LOSS_PARTIALS = 1024
COLOR_SPACE = {"linear": 0, "srgb": 1, "identity": 2}


HEADER = _PKG.parent / "include" / "acn_b200.h"
DEBUG_HEADER = _PKG.parent / "include" / "acn_b200_debug.h"


def _parse_header(path: Path):
    """argtypes for every prototype in include/acn_b200.h -- the header is the single source of
    truth for the ABI, so the binding cannot drift from it."""
    import re
    text = re.sub(r"/\*.*?\*/", "", path.read_text(), flags=re.S)
    sigs = {}
    for m in re.finditer(r"(?:int|const char\s*\*)\s+(acn_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        types = []
        if args not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                if "*" in a or a.startswith("acn_stream"):
                    types.append(c_p)
                elif a.startswith("int64_t"):
                    types.append(c_i64)
                elif a.startswith("float"):
                    types.append(c_f)
                elif a.startswith("double"):
                    types.append(C.c_double)
                elif a.startswith("uint32_t"):
                    types.append(C.c_uint32)
                elif a.startswith("int"):
                    types.append(c_int)
                else:
                    raise RuntimeError(f"acn_b200.h: cannot map parameter '{a}' of {name}")
        sigs[name] = types
    return sigs


SIGNATURES = _parse_header(HEADER)


class _Profile:
    """Optional per-kernel CUDA-event timing + launch counting (bench.py turns it on).  Events are
    recorded on torch's current stream, which is the stream every kernel is launched on."""
    enabled = False
    events: dict = {}
    launches = 0

    @classmethod
    def start(cls):
        cls.enabled, cls.events, cls.launches = True, {}, 0

    @classmethod
    def stop(cls):
        """-> {kernel: (n_launches, total_ms)}; call after torch.cuda.synchronize()."""
        cls.enabled = False
        return {k: (len(v), sum(s.elapsed_time(e) for s, e in v)) for k, v in cls.events.items()}


_NOT_KERNELS = {"acn_version", "acn_last_error", "acn_create", "acn_destroy", "acn_device_info", "acn_debug_field_trace"}


class _Bound:
    """Namespace of the bound entry points; kernel launches go through the profiling shim."""


def _wrap(name, fn):
    if name in _NOT_KERNELS:
        return fn

    def call(*args):
        if not _Profile.enabled:
            return fn(*args)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*args)
        e.record()
        _Profile.events.setdefault(name, []).append((s, e))
        _Profile.launches += 1
        return rc

    call.argtypes = fn.argtypes
    return call


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not LIB_PATH.exists():
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a).  adaptive_city_nerf_b200 has no CPU or PyTorch fallback.")
                l = C.CDLL(str(LIB_PATH))
                ns = _Bound()
                for name, argtypes in SIGNATURES.items():
                    fn = getattr(l, name)
                    fn.argtypes = argtypes
                    fn.restype = C.c_char_p if name == "acn_last_error" else c_int
                    setattr(ns, name, _wrap(name, fn))
                _lib = ns
    return _lib


def debug_lib():
    """libacn_b200_debug.so: the probes and timelines of include/acn_b200_debug.h (plus a private copy of the product
    entry points, so a traced kernel can be driven through the normal ABI).  tools/ and descriptor self-tests only."""
    global _dlib
    if _dlib is None:
        with _lock:
            if _dlib is None:
                if not DEBUG_LIB_PATH.exists():
                    raise RuntimeError(f"{DEBUG_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
                l = C.CDLL(str(DEBUG_LIB_PATH))
                ns = _Bound()
                sigs = dict(SIGNATURES)
                sigs.update(_parse_header(DEBUG_HEADER))
                for name, argtypes in sigs.items():
                    fn = getattr(l, name)
                    fn.argtypes = argtypes
                    fn.restype = C.c_char_p if name == "acn_last_error" else c_int
                    setattr(ns, name, fn)
                _dlib = ns
    return _dlib


def debug_ctx(device: torch.device) -> c_p:
    """Per-device context of the debug library (its own: the two libraries share no state)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    h = _dctx.get(idx)
    if h is None:
        out = c_p()
        debug_check(debug_lib().acn_create(idx, C.byref(out)))
        _dctx[idx] = h = out
    return h


def debug_check(rc: int) -> None:
    if rc != 0:
        msg = debug_lib().acn_last_error()
        raise RuntimeError(f"libacn_b200_debug error {rc}: {msg.decode() if msg else '?'}")


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().acn_last_error()
        raise RuntimeError(f"libacn_b200 error {rc}: {msg.decode() if msg else '?'}")


def ctx(device: torch.device) -> c_p:
    """Opaque per-device context (created on first use)."""
    if device.type != "cuda":
        raise RuntimeError(f"adaptive_city_nerf_b200 kernels need CUDA tensors (got device '{device}'); there is no CPU path")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    h = _ctx.get(idx)
    if h is None:
        with _lock:
            h = _ctx.get(idx)
            if h is None:
                out = c_p()
                check(lib().acn_create(idx, C.byref(out)))
                _ctx[idx] = h = out
    return h


def stream(device: torch.device) -> c_p:
    return c_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else c_p(t.data_ptr())


def dev_f32(t: torch.Tensor, what: str = "tensor") -> torch.Tensor:
    """Contiguous fp32 CUDA view/copy of `t`; refuses CPU tensors loudly."""
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live on a CUDA device (got {t.device}); there is no CPU path")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def pack_weights(ws) -> FieldWeights:
    st = FieldWeights()
    for i, w in enumerate(ws):
        st.p[i] = w.data_ptr() if w is not None else None
    return st
