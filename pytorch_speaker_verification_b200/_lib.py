"""ctypes binding of libsvb200.so (the C ABI declared in include/svb200.h).

There is no CPU or eager fallback: if the library is missing or a call fails, we raise.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVB_LIB_PATH") or os.path.join(_HERE, "libsvb200.so")   # (override: dev A/B builds)
CSRC = os.path.join(_HERE, "csrc")

_lib = None

ERRORS = {-1: "bad argument", -2: "CUDA error", -3: "driver entry point / tensor map error",
          -4: "misaligned pointer or stride", -5: "unsupported shape"}


class SvbError(RuntimeError):
    pass


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libsvb200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-8000:])
    if r.returncode != 0:
        raise SvbError("building libsvb200.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SvbError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.svb_last_error.restype = ctypes.c_char_p
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().svb_last_error()
        raise SvbError(f"{what} failed: {ERRORS.get(code, code)}: {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
