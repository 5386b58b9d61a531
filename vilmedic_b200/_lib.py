"""ctypes binding of libvlmb200.so (the C ABI declared in include/vlm_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, we raise.  The product path never
routes through PyTorch eager math or the CPU oracle.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# VLM_LIB=<path> selects an experimental build variant (vilmedic_b200/build.py, VLM_BUILD_TAG); default: the in-tree library.
LIB_PATH = os.environ.get("VLM_LIB") or os.path.join(_HERE, "libvlmb200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "vlm_b200.h")

_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_float = ctypes.c_float
c_u64 = ctypes.c_uint64


class VlmError(RuntimeError):
    pass


def header_symbols():
    """Every function name declared in include/vlm_b200.h (used by the CPU test that checks the exports)."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vlm_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VlmError(
                "libvlmb200.so not found at %s — run `python -m vilmedic_b200.build` (or __graft_entry__.build()). "
                "There is no CPU / eager fallback." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.vlm_last_error.restype = ctypes.c_char_p
        for name in header_symbols():
            fn = getattr(_lib, name)
            if name not in ("vlm_last_error",):
                fn.restype = c_int
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().vlm_last_error().decode("utf-8", "replace")
        raise VlmError("%s failed (rc=%d): %s" % (what, rc, msg))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
