"""ctypes binding of librbunet.so (the C ABI declared in include/rbunet.h).

The product path has NO CPU fallback: if the shared library is missing or an entry point fails,
this module raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "librbunet.so")


class GemmOperand(Structure):
    _fields_ = [("x", c_void_p), ("x_ld", c_int64), ("C", c_int), ("w", c_void_p), ("taps", c_int),
                ("dil", c_int), ("gather", c_int)]


class ConvGemmArgs(Structure):
    _fields_ = [("N", c_int), ("H", c_int), ("W", c_int), ("nseg", c_int), ("seg", GemmOperand * 2),
                ("Ncols", c_int), ("y", c_void_p), ("y_ld", c_int64), ("scatter", c_int), ("Cout", c_int),
                ("bias", c_void_p), ("addend", c_void_p), ("addend_ld", c_int64)]


_SIGS = {
    "rbu_version": (c_int, []),
    "rbu_last_error": (c_char_p, []),
    "rbu_device_check": (c_int, []),
    "rbu_sm_count": (c_int, []),
    "rbu_conv_gemm": (c_int, [POINTER(ConvGemmArgs), c_void_p]),
    "rbu_pack_weight": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rbu_conv_direct_ref": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                    c_int, c_int, c_void_p, c_void_p]),
    "rbu_loss_workspace_bytes": (c_size_t, [c_int, c_int64]),
    "rbu_loss_forward": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_float, c_float, c_float, c_float, c_void_p,
                                 c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rbu_loss_backward": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_float, c_float, c_float,
                                  c_void_p, c_void_p]),
    "rbu_confusion_counts": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_float, c_void_p, c_void_p]),
}

_lib = None


def exported_symbols():
    """Names every build of the library must export (checked by the CPU tests against include/rbunet.h)."""
    return sorted(_SIGS)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the CUDA hot path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().rbu_last_error()
        raise RuntimeError(f"librbunet {what} failed (code {rc}): {msg.decode() if msg else ''}")


def call(name: str, *args):
    """Call an int-returning entry point and raise RuntimeError with rbu_last_error() on failure."""
    check(getattr(lib(), name)(*args), name)


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
