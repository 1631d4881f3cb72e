"""ctypes binding of libsyngular_b200.so (the C ABI declared in include/syngular_b200.h).

There is NO fallback: if the shared library is missing or fails to load, importing the product raises.
PyTorch is used for device memory and streams only; every arithmetic operation of the hot path is a call
into the library.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsyngular_b200.so")


class SynError(RuntimeError):
    pass


class Index(ctypes.Structure):
    _fields_ = [("outer", ctypes.c_int64), ("inner", ctypes.c_int64), ("div", ctypes.c_int32), ("_pad", ctypes.c_int32)]


class GemmDesc(ctypes.Structure):
    _fields_ = [
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32), ("batch", ctypes.c_int32),
        ("a_m", Index), ("a_k", Index), ("a_b", Index),
        ("b_k", Index), ("b_n", Index), ("b_b", Index),
        ("c_m", Index), ("c_n", Index), ("c_b", Index),
        ("alpha", ctypes.c_double), ("beta", ctypes.c_double),
        ("mask_rows", ctypes.c_int32), ("mask_cols", ctypes.c_int32),
    ]


_BIG = 2 ** 31 - 1


def ix(spec):
    """int stride -> single-level index; (outer, inner, div) -> two-level index."""
    if isinstance(spec, Index):
        return spec
    if isinstance(spec, (tuple, list)):
        outer, inner, div = spec
        return Index(int(outer), int(inner), int(div), 0)
    return Index(0, int(spec), _BIG, 0)


def _load():
    if not os.path.exists(LIB_PATH):
        raise SynError(
            "libsyngular_b200.so not found at %s -- build it with `python -m syngular_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.syn_last_error.restype = ctypes.c_char_p
    lib.syn_launch_count.restype = ctypes.c_longlong
    return lib


lib = _load()

_vp = ctypes.c_void_p


def check(rc, what):
    if rc != 0:
        raise SynError("%s failed (rc=%d): %s" % (what, rc, lib.syn_last_error().decode()))


def stream_ptr():
    return _vp(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return _vp(t.data_ptr())


def require_cuda_f64(*tensors):
    for t in tensors:
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64):
            raise SynError("expected a CUDA float64 tensor, got %r" % (type(t) if not isinstance(t, torch.Tensor) else (t.device, t.dtype),))
