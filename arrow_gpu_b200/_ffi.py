"""ctypes binding of libagpu.so (include/agpu.h) — the only way this package reaches the GPU.

There is deliberately no fallback: if the CUDA library is missing or no device is present,
`lib()` / `GpuDevice()` raise.  Nothing under oracle/ is ever imported from here.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libagpu.so")
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "agpu.h")

# ids — keep in sync with include/agpu.h (tests/test_abi.py checks them against the header)
BOOL, I8, I16, I32, U8, U16, U32, F32, DATE32 = range(9)
ADD, SUB, MUL, DIV, REM, MIN, MAX, AND, OR, XOR, POW = range(11)
NEG, ABS, NOT, SQRT, CBRT, EXP, EXP2, LOG, LOG2, SIN, COS, ACOS, SINH = range(13)
GT, GTEQ, LT, LTEQ, EQ = range(5)
SHL, SHR = range(2)

_p, _sz, _i, _u32p = C.c_void_p, C.c_size_t, C.c_int, C.c_void_p

SIGNATURES = {
    "agpu_abi_version": (C.c_int, []),
    "agpu_device_count": (_i, [C.POINTER(C.c_int)]),
    "agpu_device_create": (_i, [_i, C.POINTER(_p)]),
    "agpu_device_destroy": (_i, [_p]),
    "agpu_device_stream": (_p, [_p]),
    "agpu_device_ordinal": (_i, [_p]),
    "agpu_launch_count": (C.c_uint64, [_p]),
    "agpu_error_string": (C.c_char_p, [_i]),
    "agpu_alloc": (_i, [_p, _sz, C.POINTER(_p)]),
    "agpu_free": (_i, [_p, _p]),
    "agpu_buffer_record_use": (_i, [_p, _p]),
    "agpu_trim": (_i, [_p]),
    "agpu_graph_begin": (_i, [_p]),
    "agpu_graph_end": (_i, [_p, C.POINTER(_p)]),
    "agpu_graph_launch": (_i, [_p, _p]),
    "agpu_graph_kernel_count": (C.c_uint64, [_p]),
    "agpu_graph_destroy": (_i, [_p]),
    "agpu_exchange_bytes": (_sz, [_i]),
    "agpu_exchange_post": (_i, [_p, _p, C.POINTER(_p), _i, _i, C.c_uint32]),
    "agpu_exchange_wait": (_i, [_p, _p, _i, C.c_uint32, _p, C.c_uint32]),
    "agpu_h2d": (_i, [_p, _p, _p, _sz]),
    "agpu_d2h": (_i, [_p, _p, _p, _sz]),
    "agpu_d2h_async": (_i, [_p, _p, _p, _sz]),
    "agpu_d2d": (_i, [_p, _p, _p, _sz]),
    "agpu_memset": (_i, [_p, _p, _i, _sz]),
    "agpu_sync": (_i, [_p]),
    "agpu_host_alloc": (_i, [_sz, C.POINTER(_p)]),
    "agpu_host_free": (_i, [_p]),
    "agpu_event_create": (_i, [C.POINTER(_p)]),
    "agpu_event_destroy": (_i, [_p]),
    "agpu_event_record": (_i, [_p, _p]),
    "agpu_event_elapsed_ms": (_i, [_p, _p, C.POINTER(C.c_float)]),
    "agpu_stream_wait_event": (_i, [_p, _p]),
    "agpu_validity_and": (_i, [_p, _u32p, _u32p, _u32p, _sz]),
    "agpu_binary": (_i, [_p, _i, _i, _p, _p, _p, _sz, _u32p, _u32p, _u32p]),
    "agpu_scalar": (_i, [_p, _i, _i, _p, _p, _p, _sz, _u32p, _u32p]),
    "agpu_unary": (_i, [_p, _i, _i, _p, _p, _sz, _u32p, _u32p]),
    "agpu_compare": (_i, [_p, _i, _i, _p, _p, _u32p, _sz, _u32p, _u32p, _u32p]),
    "agpu_shift": (_i, [_p, _i, _i, _p, _u32p, _p, _sz, _u32p, _u32p, _u32p]),
    "agpu_bitmap_binary": (_i, [_p, _i, _u32p, _u32p, _u32p, _sz, _u32p, _u32p, _u32p]),
    "agpu_bitmap_not": (_i, [_p, _u32p, _u32p, _sz, _u32p, _u32p]),
    "agpu_cast": (_i, [_p, _i, _i, _p, _p, _sz, _u32p, _u32p]),
    "agpu_fused_mul_add_gt": (_i, [_p, _p, _p, _p, _p, _u32p, _sz, _u32p, _u32p, _u32p, _u32p, _u32p]),
    "agpu_merge": (_i, [_p, _i, _p, _p, _u32p, _p, _sz, _u32p, _u32p, _u32p, _u32p]),
    "agpu_take": (_i, [_p, _i, _p, _sz, _u32p, _p, _sz, _u32p, _u32p]),
    "agpu_put": (_i, [_p, _i, _p, _sz, _u32p, _p, _sz, _u32p, _sz]),
    "agpu_filter_scratch_bytes": (_sz, [_sz]),
    "agpu_filter_count": (_i, [_p, _u32p, _u32p, _sz, _p, _p]),
    "agpu_filter_count_post": (_i, [_p, _u32p, _u32p, _sz, _p, _p, C.POINTER(_p), _i, _i, C.c_uint32]),
    "agpu_filter_scatter": (_i, [_p, _i, _p, _u32p, _u32p, _u32p, _sz, _p, _p, _u32p, _sz]),
    "agpu_ipc_alloc": (_i, [_p, _sz, C.POINTER(_p)]),
    "agpu_ipc_free": (_i, [_p, _p]),
    "agpu_ipc_export": (_i, [_p, _p, C.c_char_p]),
    "agpu_ipc_open": (_i, [_p, C.c_char_p, C.POINTER(_p)]),
    "agpu_ipc_close": (_i, [_p, _p]),
    "agpu_take_sharded": (_i, [_p, _i, _i, C.POINTER(_p), C.POINTER(_p), C.POINTER(C.c_uint64), _u32p, _p, _sz, _u32p]),
    "agpu_broadcast": (_i, [_p, _i, _p, _p, _sz]),
    "agpu_sum": (_i, [_p, _i, _p, _sz, _p]),
    "agpu_any": (_i, [_p, _u32p, _sz, _u32p]),
    "agpu_all": (_i, [_p, _u32p, _sz, _u32p]),
}



class ChainStep(C.Structure):
    """agpu_chain_step"""
    _fields_ = [("kind", C.c_int32), ("op", C.c_int32), ("operand", C.c_void_p), ("validity", C.c_void_p),
                ("scalar", C.c_float)]


STEP_UNARY, STEP_BINARY_COLUMN, STEP_BINARY_SCALAR, STEP_COMPARE_COLUMN, STEP_COMPARE_SCALAR = range(5)
STEP_BINARY_DEVSCALAR, STEP_COMPARE_DEVSCALAR = 5, 6
STEP_SHIFT_COLUMN = 7
STEP_STORE, STEP_RESET = 8, 9          # agpu_fused_chain_pair only
CHAIN_MAX_STEPS = 8
SIGNATURES["agpu_fused_chain"] = (_i, [_p, _i, _p, _u32p, C.POINTER(ChainStep), _i, _p, _sz, _u32p])
SIGNATURES["agpu_fused_chain_pair"] = (_i, [_p, _i, _p, _u32p, C.POINTER(ChainStep), _i, _p, _p, _sz, _u32p])
SIGNATURES["agpu_fused_chain_int"] = (_i, [_p, _i, _p, _u32p, C.POINTER(ChainStep), _i, _p, _sz, _u32p])

_lib = None


class AgpuError(RuntimeError):
    def __init__(self, code: int, what: str):
        self.code = code
        msg = "?"
        if _lib is not None:
            msg = _lib.agpu_error_string(code).decode()
        super().__init__(f"{what} failed: {msg} (code {code})")


def header_functions() -> list[str]:
    """Names of every function include/agpu.h declares."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(agpu_[a-z0-9_]+)\s*\(", text)))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libagpu.so in-tree with nvcc for sm_100a (see csrc/Makefile)."""
    cmd = ["make", "-C", CSRC, "-j8"] + (["-B"] if force else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode:
        raise RuntimeError("building libagpu.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    """Load libagpu.so, declaring every signature.  Raises if the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        raise AgpuError(code, what)


def device_count() -> int:
    n = C.c_int(0)
    lib().agpu_device_count(C.byref(n))
    return n.value
