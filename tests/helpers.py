"""Shared test machinery: golden-vector loading, an oracle-side model of the reference's array
layer (values + validity, the host glue of each operator crate), and comparators.

`OArr` + `ORACLE_OPS` restate, on top of oracle/oracle.c, what the reference's host code does
around each shader: which validity rule applies (AND of both / copy), output dtype, lengths.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "reference_vectors.json")

DTYPE_OF = {"Float32ArrayGPU": O.F32, "Int32ArrayGPU": O.I32, "UInt32ArrayGPU": O.U32, "Int16ArrayGPU": O.I16,
            "UInt16ArrayGPU": O.U16, "Int8ArrayGPU": O.I8, "UInt8ArrayGPU": O.U8, "Date32ArrayGPU": O.DATE32,
            "BooleanArrayGPU": O.BOOL}
TYPE_OF_ARROW = {"Float32Type": "Float32ArrayGPU", "Int32Type": "Int32ArrayGPU", "UInt32Type": "UInt32ArrayGPU",
                 "Int16Type": "Int16ArrayGPU", "UInt16Type": "UInt16ArrayGPU", "Int8Type": "Int8ArrayGPU",
                 "UInt8Type": "UInt8ArrayGPU", "Date32Type": "Date32ArrayGPU", "BooleanType": "BooleanArrayGPU"}


def load_cases(*macros):
    cases = json.load(open(GOLDEN))["cases"]
    return [c for c in cases if not macros or c["macro"] in macros]


def decode(v):
    """JSON value -> python value (NaN/inf strings, f32 bit patterns)"""
    if isinstance(v, list):
        return [decode(x) for x in v]
    if isinstance(v, dict):
        return float(np.array([v["f32_bits"]], dtype=np.uint32).view(np.float32)[0])
    if v == "NaN":
        return math.nan
    if v == "Infinity":
        return math.inf
    if v == "-Infinity":
        return -math.inf
    return v


def bits_of(v):
    """expected f32 given as bit pattern (bitcast test) or None"""
    if isinstance(v, dict):
        return v["f32_bits"]
    return None


# ------------------------------------------------------------------------------------------
# oracle-side array model
# ------------------------------------------------------------------------------------------
@dataclass
class OArr:
    dtype: int
    data: np.ndarray                 # values (numpy typed) or bitmap words (uint32) for BOOL
    n: int
    valid: Optional[np.ndarray] = None  # validity bitmap words or None

    @classmethod
    def from_slice(cls, dtype, values):
        if dtype == O.BOOL:
            return cls(dtype, O.pack_bits(values), len(values))
        return cls(dtype, np.asarray(values).astype(O.NP[dtype]), len(values))

    @classmethod
    def from_optional(cls, dtype, values):
        """primitive_array_gpu.rs:22-55 / boolean_gpu.rs:23-50: nulls hold T::default()"""
        valid = O.pack_bits([v is not None for v in values])
        dense = [(False if dtype == O.BOOL else 0) if v is None else v for v in values]
        a = cls.from_slice(dtype, dense)
        a.valid = valid
        return a

    def raw_values(self):
        if self.dtype == O.BOOL:
            return O.unpack_bits(self.data, self.n)
        return self.data[: self.n]

    def values(self):
        raw = self.raw_values()
        if self.valid is None:
            return [x.item() for x in raw]
        ok = O.unpack_bits(self.valid, self.n)
        return [raw[i].item() if ok[i] else None for i in range(self.n)]


_BIN = {"add": O.ADD, "sub": O.SUB, "mul": O.MUL, "div": O.DIV, "min": O.MIN, "max": O.MAX,
        "bitwise_and": O.AND, "bitwise_or": O.OR, "bitwise_xor": O.XOR, "power": O.POW}
_SCALAR = {"add_scalar": O.ADD, "sub_scalar": O.SUB, "mul_scalar": O.MUL, "div_scalar": O.DIV, "rem_scalar": O.REM}
_CMP = {"gt": O.GT, "gteq": O.GTEQ, "lt": O.LT, "lteq": O.LTEQ, "eq": O.EQ}
_UN = {"neg": O.NEG, "abs": O.ABS, "bitwise_not": O.NOT, "sqrt": O.SQRT, "cbrt": O.CBRT, "exp": O.EXP,
       "exp2": O.EXP2, "log": O.LOG, "log2": O.LOG2, "sin": O.SIN, "cos": O.COS, "acos": O.ACOS, "sinh": O.SINH}
_SHIFT = {"bitwise_shl": O.SHL, "bitwise_shr": O.SHR}


def oracle_binary(op: str, a: OArr, b: OArr) -> OArr:
    """validity rule V2 = AND of both bitmaps (null_bit_buffer.rs:206-243)"""
    v = O.validity_and(a.valid, b.valid, a.n)
    if op in _CMP:
        return OArr(O.BOOL, O.compare(_CMP[op], a.dtype, a.data, b.data), a.n, v)
    if op in _SHIFT:
        return OArr(a.dtype, O.shift(_SHIFT[op], a.dtype, a.data, b.data), a.n, v)
    if a.dtype == O.BOOL:
        return OArr(O.BOOL, O.bitmap_binary(_BIN[op], a.data, b.data, a.n), a.n, v)
    return OArr(a.dtype, O.binary(_BIN[op], a.dtype, a.data, b.data), a.n, v)


def oracle_scalar(op: str, a: OArr, s: OArr) -> OArr:
    """validity rule V1 = copy (arithmetic/src/lib.rs:35-38)"""
    return OArr(a.dtype, O.scalar(_SCALAR[op], a.dtype, a.data, s.data[0]), a.n,
                None if a.valid is None else a.valid.copy())


def oracle_unary(op: str, a: OArr) -> OArr:
    v = None if a.valid is None else a.valid.copy()
    if a.dtype == O.BOOL:
        assert op == "bitwise_not"
        return OArr(O.BOOL, O.bitmap_not(a.data, a.n), a.n, v)
    out = O.unary(_UN[op], a.dtype, a.data)
    odt = O.F32 if out.dtype == np.float32 else a.dtype
    return OArr(odt, out, a.n, v)


def oracle_cast(a: OArr, dst: int) -> OArr:
    v = None if a.valid is None else a.valid.copy()
    if a.dtype == O.U32 and dst == O.F32:  # bitcast: buffer copy (cast/src/lib.rs:90-108)
        return OArr(O.F32, a.data.view(np.float32).copy(), a.n, v)
    return OArr(dst, O.cast(a.dtype, dst, a.data, a.n), a.n, v)


def oracle_merge(a: OArr, b: OArr, mask: OArr) -> OArr:
    v = O.merge_validity(a.valid, b.valid, mask.data, mask.valid, a.n)
    return OArr(a.dtype, O.merge(a.dtype, a.data, b.data, mask.data, a.n), a.n, v)


def oracle_take(a: OArr, idx: OArr) -> OArr:
    v = None if a.valid is None else O.take(O.BOOL, a.valid, a.n, idx.data)
    return OArr(a.dtype, O.take(a.dtype, a.data, a.n, idx.data), idx.n, v)


def oracle_put(src: OArr, si: OArr, dst: OArr, di: OArr) -> OArr:
    return OArr(dst.dtype, O.put(src.dtype, src.data, si.data, dst.data, di.data), dst.n)


def oracle_filter(a: OArr, mask: OArr) -> OArr:
    out, vout, k = O.filter(a.dtype, a.data, a.valid, mask.data, mask.valid)
    return OArr(a.dtype, out, k, vout)


# ------------------------------------------------------------------------------------------
# comparators
# ------------------------------------------------------------------------------------------
def float_eq_in_error(left: float, right: float) -> bool:
    """crates/test_macros/src/lib.rs:88-109 — the reference's own float tolerance"""
    if math.isnan(left) != math.isnan(right):
        return False
    if math.isnan(left):
        return True
    if (left == -math.inf) != (right == -math.inf):
        return False
    if (left == math.inf) != (right == math.inf):
        return False
    if math.isinf(left):
        return True
    return abs(abs(left) - abs(right)) <= 0.01


def assert_values(got, expected, *, float_tol: bool, what: str):
    got = [x.item() if hasattr(x, "item") else x for x in got]
    assert len(got) == len(expected), f"{what}: length {len(got)} != {len(expected)}"
    for i, (g, e) in enumerate(zip(got, expected)):
        if e is None or g is None:
            assert g is None and e is None, f"{what}[{i}]: {g!r} != {e!r}\n got {got}\n exp {expected}"
        elif float_tol or isinstance(e, float):
            if float_tol:
                ok = float_eq_in_error(float(e), float(g))
            else:
                ok = (math.isnan(e) and math.isnan(g)) or float(np.float32(e)) == float(g)
            assert ok, f"{what}[{i}]: {g!r} != {e!r}\n got {got}\n exp {expected}"
        else:
            assert g == e, f"{what}[{i}]: {g!r} != {e!r}\n got {got}\n exp {expected}"


def same_f32_bits(got, want) -> bool:
    """bit-for-bit equality of two f32 arrays, except that any NaN equals any NaN: IEEE 754 and
    WGSL leave NaN sign/payload unspecified (x86 SSE produces 0xFFC00000 for invalid operations
    and propagates input payloads, sm_100 produces the canonical 0x7FFFFFFF), and the reference's
    own float comparator treats NaN == NaN (crates/test_macros/src/lib.rs:89-94)."""
    g = np.ascontiguousarray(got, dtype=np.float32)
    w = np.ascontiguousarray(want, dtype=np.float32)
    if g.shape != w.shape:
        return False
    gn, wn = np.isnan(g), np.isnan(w)
    if not np.array_equal(gn, wn):
        return False
    return bool(np.array_equal(g.view(np.uint32)[~gn], w.view(np.uint32)[~wn]))


def ulp_diff(got: np.ndarray, ref: np.ndarray) -> np.ndarray:
    """distance in f32 ULPs; NaN vs NaN and equal infinities count as 0"""
    g = np.asarray(got, dtype=np.float32)
    r = np.asarray(ref, dtype=np.float32)

    def key(x):
        i = x.view(np.int32).astype(np.int64)
        return np.where(i < 0, np.int64(-2147483648) - i, i)

    d = np.abs(key(g) - key(r))
    both_nan = np.isnan(g) & np.isnan(r)
    d = np.where(both_nan, 0, d)
    one_nan = np.isnan(g) ^ np.isnan(r)
    return np.where(one_nan, np.int64(1) << 40, d)
