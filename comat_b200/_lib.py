"""ctypes loader for libcomat_b200.so — the only way the package reaches the GPU.

No CPU fallback: if the library is missing or a call fails this raises; nothing silently routes around it.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcomat_b200.so")
_lib = None
LAUNCH_COUNT = 0          # number of C-ABI compute calls issued (bench.py reports it as gpu_launches evidence)


class ComatError(RuntimeError):
    pass


class AttnmapPlan(C.Structure):
    _fields_ = [("n_maps", C.c_int32), ("n_groups", C.c_int32), ("n_samples", C.c_int32), ("n_words", C.c_int32),
                ("n_pairs", C.c_int32), ("n_work", C.c_int32), ("tokens", C.c_int32), ("max_heads", C.c_int32),
                ("max_maps_per_group", C.c_int32), ("_pad", C.c_int32), ("pred_floats", C.c_int64),
                ("map_ptr", C.c_void_p), ("grp", C.c_void_p), ("smp", C.c_void_p), ("pair", C.c_void_p),
                ("word_ntok", C.c_void_p), ("work", C.c_void_p), ("masks", C.c_void_p)]


def _declare(lib):
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    lib.comat_version.restype = C.c_int
    lib.comat_strerror.restype = C.c_char_p
    lib.comat_strerror.argtypes = [C.c_int]
    lib.comat_last_cuda_error.restype = C.c_int
    lib.comat_attnmap_loss_state_floats.restype = sz
    lib.comat_attnmap_loss_state_floats.argtypes = [C.POINTER(AttnmapPlan)]
    lib.comat_attnmap_loss_fwd.argtypes = [C.POINTER(AttnmapPlan), vp, vp, sz, vp, vp]
    lib.comat_attnmap_loss_bwd.argtypes = [C.POINTER(AttnmapPlan), vp, vp, vp, vp]
    lib.comat_mask_resize_any.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    for name, args in _EXTRA_SIGS.items():
        if hasattr(lib, name):
            getattr(lib, name).argtypes = args
    return lib


_EXTRA_SIGS = {}


def register_signature(name, argtypes):
    _EXTRA_SIGS[name] = argtypes
    if _lib is not None and hasattr(_lib, name):
        getattr(_lib, name).argtypes = argtypes


def lib():
    """Load the C-ABI library, (re)building it first when the sources changed (build.build is a no-op when the digest of
    csrc/ + include/ matches the stamp; it takes a file lock so concurrent ranks do not race).  Raises when unavailable."""
    global _lib
    if _lib is None:
        from . import build as _b
        try:
            _b.build()
        except Exception:
            if not os.path.exists(LIB_PATH):        # no nvcc on this box and no prebuilt library: nothing to run on
                raise
        _lib = _declare(C.CDLL(LIB_PATH))
    return _lib


def check(status: int, what: str = ""):
    if status != 0:
        L = lib()
        msg = L.comat_strerror(status).decode()
        cuda = L.comat_last_cuda_error()
        raise ComatError(f"comat_b200 {what}: {msg} (status {status}, cudaError {cuda})")


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ComatError("comat_b200 ops run on CUDA tensors only (no CPU fallback); got a %s tensor" % t.device)


def count_launch(n=1):
    global LAUNCH_COUNT
    LAUNCH_COUNT += n
