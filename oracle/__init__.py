"""Python face of the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
as the checker; never by the arrow_gpu_b200 package.  Functions take and return numpy arrays
on packed buffers exactly like the C ABI does on device buffers.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

# ids shared with include/agpu.h
BOOL, I8, I16, I32, U8, U16, U32, F32, DATE32 = range(9)
ADD, SUB, MUL, DIV, REM, MIN, MAX, AND, OR, XOR, POW = range(11)
NEG, ABS, NOT, SQRT, CBRT, EXP, EXP2, LOG, LOG2, SIN, COS, ACOS, SINH = range(13)
GT, GTEQ, LT, LTEQ, EQ = range(5)
SHL, SHR = range(2)

NP = {I8: np.dtype("i1"), I16: np.dtype("<i2"), I32: np.dtype("<i4"), U8: np.dtype("u1"),
      U16: np.dtype("<u2"), U32: np.dtype("<u4"), F32: np.dtype("<f4"), DATE32: np.dtype("<i4")}

_lib = None


def build(force: bool = False) -> str:
    res = subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), capture_output=True, text=True)
    if res.returncode:
        raise RuntimeError("building liboracle.so failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _chk(rc, what):
    if rc != 0:
        raise ValueError(f"oracle {what}: unsupported/invalid (code {rc})")


def words(n_bits: int) -> int:
    return (n_bits + 31) // 32


def pack_bits(flags) -> np.ndarray:
    flags = np.asarray(flags, dtype=bool)
    out = np.zeros(words(len(flags)) * 4, dtype=np.uint8)
    if len(flags):
        pk = np.packbits(flags, bitorder="little")
        out[: len(pk)] = pk
    return out.view(np.uint32)


def unpack_bits(bits: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(np.ascontiguousarray(bits).view(np.uint8), bitorder="little")[:n].astype(bool)


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(C.c_int(n))


def set_ref_quirks(on: bool) -> None:
    lib().oracle_set_ref_quirks(C.c_int(1 if on else 0))


def _arr(a, dtype):
    return np.ascontiguousarray(np.asarray(a).astype(NP[dtype], copy=False))


def validity_and(va, vb, n_bits):
    if va is None and vb is None:
        return None
    out = np.zeros(words(n_bits), dtype=np.uint32)
    _chk(lib().oracle_validity_and(_p(va), _p(vb), _p(out), C.c_size_t(n_bits)), "validity_and")
    return out


def binary(op, dtype, a, b, out=None):
    a, b = _arr(a, dtype), _arr(b, dtype)
    if out is None:
        out = np.empty(len(a), dtype=NP[dtype])
    _chk(lib().oracle_binary(op, dtype, _p(a), _p(b), _p(out), C.c_size_t(len(a))), "binary")
    return out


def scalar(op, dtype, a, s, out=None):
    a = _arr(a, dtype)
    sv = np.zeros(4, dtype=NP[dtype])  # padded like the reference's scalar buffer
    sv[0] = np.asarray(s).astype(NP[dtype])
    if out is None:
        out = np.empty(len(a), dtype=NP[dtype])
    _chk(lib().oracle_scalar(op, dtype, _p(a), _p(sv), _p(out), C.c_size_t(len(a))), "scalar")
    return out


def unary(op, dtype, a, out=None):
    a = _arr(a, dtype)
    to_f32 = op >= SQRT or (op in (NEG, ABS) and dtype == F32)
    odt = NP[F32] if (to_f32 and dtype != I32) else NP[dtype]
    if out is None:
        out = np.empty(len(a), dtype=odt)
    _chk(lib().oracle_unary(op, dtype, _p(a), _p(out), C.c_size_t(len(a))), "unary")
    return out


def compare(op, dtype, a, b):
    a, b = _arr(a, dtype), _arr(b, dtype)
    out = np.zeros(words(len(a)), dtype=np.uint32)
    _chk(lib().oracle_compare(op, dtype, _p(a), _p(b), _p(out), C.c_size_t(len(a))), "compare")
    return out


def shift(op, dtype, a, counts, out=None):
    a = _arr(a, dtype)
    counts = np.ascontiguousarray(np.asarray(counts).astype(np.uint32, copy=False))
    if out is None:
        out = np.empty(len(a), dtype=NP[dtype])
    _chk(lib().oracle_shift(op, dtype, _p(a), _p(counts), _p(out), C.c_size_t(len(a))), "shift")
    return out


def bitmap_binary(op, a, b, n_bits):
    out = np.zeros(words(n_bits), dtype=np.uint32)
    _chk(lib().oracle_bitmap_binary(op, _p(a), _p(b), _p(out), C.c_size_t(n_bits)), "bitmap_binary")
    return out


def bitmap_not(a, n_bits):
    out = np.zeros(words(n_bits), dtype=np.uint32)
    _chk(lib().oracle_bitmap_not(_p(a), _p(out), C.c_size_t(n_bits)), "bitmap_not")
    return out


def cast(src, dst, a, n=None, out=None):
    if src == BOOL:
        a = np.ascontiguousarray(a, dtype=np.uint32)
        assert n is not None
    else:
        a = _arr(a, src)
        n = len(a)
    if out is None:
        out = np.empty(n, dtype=NP[dst])
    _chk(lib().oracle_cast(src, dst, _p(a), _p(out), C.c_size_t(n)), "cast")
    return out


def merge(dtype, a, b, mask_bits, n=None):
    if dtype == BOOL:
        out = np.zeros(words(n), dtype=np.uint32)
        _chk(lib().oracle_merge(dtype, _p(a), _p(b), _p(mask_bits), _p(out), C.c_size_t(n)), "merge")
        return out
    a, b = _arr(a, dtype), _arr(b, dtype)
    out = np.empty(len(a), dtype=NP[dtype])
    _chk(lib().oracle_merge(dtype, _p(a), _p(b), _p(mask_bits), _p(out), C.c_size_t(len(a))), "merge")
    return out


def merge_validity(va, vb, mask_bits, vmask, n):
    if va is None and vb is None and vmask is None:
        return None
    out = np.zeros(words(n), dtype=np.uint32)
    _chk(lib().oracle_merge_validity(_p(va), _p(vb), _p(mask_bits), _p(vmask), _p(out), C.c_size_t(n)),
         "merge_validity")
    return out


def take(dtype, src, src_len, idx):
    idx = np.ascontiguousarray(np.asarray(idx).astype(np.uint32, copy=False))
    if dtype == BOOL:
        out = np.zeros(words(len(idx)), dtype=np.uint32)
        src = np.ascontiguousarray(src, dtype=np.uint32)
    else:
        src = _arr(src, dtype)
        out = np.empty(len(idx), dtype=NP[dtype])
    _chk(lib().oracle_take(dtype, _p(src), C.c_size_t(src_len), _p(idx), _p(out), C.c_size_t(len(idx))), "take")
    return out


def put(dtype, src, src_idx, dst, dst_idx):
    src_idx = np.ascontiguousarray(np.asarray(src_idx).astype(np.uint32, copy=False))
    dst_idx = np.ascontiguousarray(np.asarray(dst_idx).astype(np.uint32, copy=False))
    dst = np.array(dst, copy=True)
    _chk(lib().oracle_put(dtype, _p(np.ascontiguousarray(src)), _p(src_idx), _p(dst), _p(dst_idx),
                          C.c_size_t(len(src_idx))), "put")
    return dst


def filter(dtype, src, vsrc, mask_bits, vmask):  # noqa: A001
    src = _arr(src, dtype)
    n = len(src)
    out = np.empty(n, dtype=NP[dtype])
    vout = np.zeros(words(n) + 1, dtype=np.uint32) if vsrc is not None else None
    count = C.c_uint64(0)
    _chk(lib().oracle_filter(dtype, _p(src), _p(vsrc), _p(mask_bits), _p(vmask), C.c_size_t(n), _p(out),
                             _p(vout), C.byref(count)), "filter")
    k = count.value
    return out[:k].copy(), (vout[: words(k)].copy() if vout is not None else None), k


def broadcast(dtype, value, n):
    out = np.empty(n, dtype=NP[dtype])
    s = np.asarray([value]).astype(NP[dtype])
    _chk(lib().oracle_broadcast(dtype, _p(s), _p(out), C.c_size_t(n)), "broadcast")
    return out


def sum(dtype, a):  # noqa: A001
    a = _arr(a, dtype)
    out = np.zeros(1, dtype=NP[dtype])
    _chk(lib().oracle_sum(dtype, _p(a), C.c_size_t(len(a)), _p(out)), "sum")
    return out[0]


def any(bits, n_bits) -> bool:  # noqa: A001
    r = C.c_uint32(0)
    _chk(lib().oracle_any(_p(np.ascontiguousarray(bits, dtype=np.uint32)), C.c_size_t(n_bits), C.byref(r)), "any")
    return bool(r.value)


def all(bits, n_bits) -> bool:  # noqa: A001
    r = C.c_uint32(0)
    _chk(lib().oracle_all(_p(np.ascontiguousarray(bits, dtype=np.uint32)), C.c_size_t(n_bits), C.byref(r)), "all")
    return bool(r.value)
