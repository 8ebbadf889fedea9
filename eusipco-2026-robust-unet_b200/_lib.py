"""ctypes binding of librbunet.so, generated from the C ABI declared in include/rbunet.h.

The product path has NO CPU fallback: if the shared library is missing or an entry point fails,
this module raises.
"""
import ctypes
import os
import re
from ctypes import Structure, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_ulonglong, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "librbunet.so")
HEADER_PATH = os.path.join(os.path.dirname(PKG_DIR), "include", "rbunet.h")


class GemmOperand(Structure):
    _fields_ = [("x", c_void_p), ("x_ld", c_int64), ("C", c_int), ("w", c_void_p), ("taps", c_int),
                ("dil", c_int), ("gather", c_int)]


class ConvGemmArgs(Structure):
    _fields_ = [("N", c_int), ("H", c_int), ("W", c_int), ("nseg", c_int), ("seg", GemmOperand * 2),
                ("Ncols", c_int), ("y", c_void_p), ("y_ld", c_int64), ("scatter", c_int), ("Cout", c_int),
                ("bias", c_void_p), ("addend", c_void_p), ("addend_ld", c_int64), ("scale", c_void_p), ("relu", c_int),
                ("stats", c_void_p), ("tile_stats", c_void_p)]


class WgradArgs(Structure):
    _fields_ = [("N", c_int), ("H", c_int), ("W", c_int), ("a", c_void_p), ("a_ld", c_int64), ("Ca", c_int),
                ("b", c_void_p), ("b_ld", c_int64), ("Cb", c_int), ("taps", c_int), ("dil", c_int),
                ("gather", c_int), ("out", c_void_p), ("accumulate", c_int)]


_SCALARS = {"int": c_int, "int64_t": c_int64, "size_t": c_size_t, "float": c_float, "double": c_double}


def _ctype(decl: str):
    decl = decl.strip()
    if decl in ("void", ""):
        return None
    if "*" in decl:
        return c_char_p if decl.replace(" ", "") == "constchar*" else c_void_p
    words = decl.replace("const", "").split()
    if words[:3] == ["unsigned", "long", "long"]:
        return c_ulonglong
    if words[:2] == ["long", "long"]:
        return c_int64
    base = decl.replace("const", "").split()[0]
    return _SCALARS[base]


def parse_header(path: str = HEADER_PATH):
    """{name: (restype, [argtypes])} for every `rbu_*` prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)     # struct bodies are not prototypes
    sigs = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(rbu_\w+)\s*\(([^;{}]*?)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3)
        argtypes = []
        for a in args.split(","):
            a = a.strip()
            if a in ("void", ""):
                continue
            # drop the parameter name
            a = re.sub(r"\b\w+\s*(\[\w*\])?$", "", a).strip() if not a.endswith("*") else a
            argtypes.append(_ctype(a))
        sigs[name] = (_ctype(ret), argtypes)
    return sigs


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the CUDA hot path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in parse_header().items():
            fn = getattr(handle, name)   # AttributeError if the build lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().rbu_last_error()
        raise RuntimeError(f"librbunet {what} failed (code {rc}): {msg.decode() if msg else ''}")


class Profiler:
    """Optional per-call device timing (CUDA events on the launching stream) used by bench.py for the roofline
    numbers: every C-ABI call made while a profiler is installed is bracketed by two events and tagged with its
    algorithmic work (`flops` for the tensor-core GEMMs, `bytes` for the bandwidth-bound kernels)."""

    def __init__(self):
        self.records = []      # (entry point, tag, flops, bytes, start event, stop event)

    def detail(self):
        """Per-call rows (label, ms, flops, bytes), slowest first."""
        import torch
        torch.cuda.synchronize()
        rows = [(lab or tag or name, e0.elapsed_time(e1), flops, nbytes) for name, tag, flops, nbytes, e0, e1, lab in self.records]
        return sorted(rows, key=lambda r: -r[1])

    def summary(self):
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, tag, flops, nbytes, e0, e1, _ in self.records:
            d = out.setdefault(tag or name, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["calls"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["bytes"] += nbytes
        return out


PROFILER = None


def _nz(a):
    """True if a ctypes pointer argument is non-NULL."""
    return bool(getattr(a, "value", a))


# Algorithmic bytes of the bandwidth-bound entry points (distinct tensor bytes read + written, each once, SURVEY.md
# §8d) as a function of the positional arguments declared in include/rbunet.h -- used only by the profiler.
ALGO_BYTES = {
    "rbu_bn_stats": lambda a: a[2] * a[3] * a[4] * 2 if (a[5] or a[6]) else 0,
    "rbu_affine_act": lambda a: 2 * a[4] * a[6] * 2,
    "rbu_sa_reduce": lambda a: a[2] * a[4] * 2 + a[2] * 12,
    "rbu_sa_gate": lambda a: a[1] * a[2] * a[3] * 12,
    "rbu_rb_out": lambda a: 3 * a[6] * a[8] * 2 + a[6] * 4,
    "rbu_maxpool2x2": lambda a: a[4] * a[5] * a[6] * a[7] * 2 * 5,
    "rbu_ag_psi": lambda a: 2 * a[4] * a[5] * 2 + a[4] * 4,
    "rbu_ag_apply": lambda a: 2 * a[4] * a[5] * 2 + a[4] * 8,
    "rbu_stem_im2col": lambda a: a[1] * a[3] * a[4] * (a[2] * 4 + a[5] * 2),
    "rbu_head_forward": lambda a: a[2] * a[3] * 2 + a[2] * 4,
    "rbu_head_backward": lambda a: a[6] * 8 + 2 * a[6] * a[7] * 2,
    "rbu_rb_bwd1": lambda a: (4 + (1 if _nz(a[8]) else 0)) * a[10] * a[11] * a[12] * 2 + a[10] * a[11] * 4,
    "rbu_sa_bwd": lambda a: a[3] * a[4] * a[5] * 24,
    "rbu_rb_bwd2": lambda a: 2 * a[4] * a[5] * a[6] * 2 + a[4] * a[5] * 16,
    "rbu_rb_bwd3": lambda a: (3 + (2 if _nz(a[6]) else 0)) * a[10] * a[11] * a[12] * 2 + a[10] * a[11] * 16,
    "rbu_bn_bwd": lambda a: 5 * a[6] * a[7] * a[8] * 2,
    "rbu_ag_bwd": lambda a: (3 * a[16] + 6 * a[17]) * a[14] * a[15] * 2 + a[14] * a[15] * 16,
    "rbu_maxpool2x2_bwd": lambda a: a[6] * a[7] * a[8] * a[9] * 2 * (9 + 4 * a[10]),
    "rbu_chan_sum": lambda a: a[2] * a[3] * 2,
    "rbu_loss_forward": lambda a: a[2] * a[3] * 8,
    "rbu_loss_backward": lambda a: a[2] * 12,
    "rbu_confusion_counts": lambda a: a[2] * a[3] * 8,
}


def call(name: str, *args, tag=None, flops=0.0, nbytes=0.0, label=None):
    """Call an int-returning entry point and raise RuntimeError with rbu_last_error() on failure."""
    prof = PROFILER
    if prof is None:
        check(getattr(lib(), name)(*args), name)
        return
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(lib(), name)(*args), name)
    e1.record()
    if not nbytes and name in ALGO_BYTES:
        nbytes = float(ALGO_BYTES[name](args))
    prof.records.append((name, tag, flops, nbytes, e0, e1, label))


def launch_count() -> int:
    return int(lib().rbu_launch_count())


def stream_ptr():
    """torch's current stream of the CURRENT device.  One device per process is the supported deployment (torchrun);
    Engine.forward/backward and the functional ops enter `torch.cuda.device(tensor.device)` first, so a model living
    on a device other than the current one still launches on its own device's stream."""
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
