"""ctypes binding of libln3b200.so -- the reference-side stub a maintainer would add.

The reference (pure PyTorch) has no FFI; this file is the whole "binding": argument structs
mirroring include/ln3b200.h field by field, and one checked call helper.  There is no CPU
fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libln3b200.so"

LN3_OK = 0
ACT_NONE, ACT_GELU_ERF, ACT_GELU_TANH, ACT_SILU = 0, 1, 2, 3
OUT_BF16, OUT_F32, OUT_RESID_F32 = 0, 1, 2

_lib = None


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("W", C.c_void_p), ("bias", C.c_void_p), ("out", C.c_void_p),
        ("out2", C.c_void_p), ("gate", C.c_void_p),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("lda", C.c_longlong), ("ldw", C.c_longlong), ("ldo", C.c_longlong),
        ("ldo2", C.c_longlong), ("gate_ld", C.c_longlong),
        ("gate_rows", C.c_int), ("act", C.c_int), ("out_kind", C.c_int),
    ]


def lib() -> C.CDLL:
    """Load libln3b200.so (once).  Raises if it has not been built -- no silent fallback."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -m ln3diff_b200.build` (or "
                "__graft_entry__.build()); the hot path has no CPU fallback")
        L = C.CDLL(str(LIB_PATH))
        L.ln3_abi_version.restype = C.c_int
        L.ln3_last_error.restype = C.c_char_p
        L.ln3_launch_count.restype = C.c_ulonglong
        _lib = L
    return _lib


def check(rc: int, what: str = "ln3") -> None:
    if rc != LN3_OK:
        msg = lib().ln3_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(lib().ln3_launch_count())


def current_stream() -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
