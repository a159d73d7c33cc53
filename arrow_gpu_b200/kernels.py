"""Operator layer: the host-side mirror of the reference's operator crates
(`arithmetic`, `compare`, `logical`, `cast`, `math`, `trigonometry`, `routines`).

Every trait method of the reference exists here under the same name in two forms, like the
reference: the eager form (`a.add(b)`) and the recording form (`a.add_op(b, pipeline)`), plus the
`*_dyn` / `*_op_dyn` free functions that dispatch on the runtime array type and raise `Panic`
for unsupported pairs (the reference `panic!`s).  Each call is ONE kernel launch through the
C ABI (include/agpu.h): value kernel and validity-bitmap kernel of the reference are fused.

Nothing here computes on the host; there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _ffi
from ._ffi import check, lib
from .array import (ARRAY_TYPES, ArrowComputePipeline, ArrowType, BooleanArrayGPU, Date32ArrayGPU,
                    Float32ArrayGPU, Int8ArrayGPU, Int16ArrayGPU, Int32ArrayGPU, NullBitBufferGpu, Panic,
                    PrimitiveArrayGpu, UInt8ArrayGPU, UInt16ArrayGPU, UInt32ArrayGPU, _new_validity, _vptr,
                    bitmap_words)

_INT_TYPES = (Int32ArrayGPU, UInt32ArrayGPU, Int16ArrayGPU, UInt16ArrayGPU, Int8ArrayGPU, UInt8ArrayGPU)
_I32_BACKED = (Int32ArrayGPU, Date32ArrayGPU)


def _pipeline_for(arr) -> ArrowComputePipeline:
    return ArrowComputePipeline(arr.get_gpu_device(), None)


def _eager(op_fn):
    """default_impl! of the reference: new pipeline -> *_op -> finish.  The pipeline of an eager
    call records nothing (ops enqueue on the stream at once), so one plain pipeline per device
    handle serves every eager call instead of a new object per op."""
    def run(self, *args):
        dev = self.gpu_device
        pipeline = getattr(dev, "_eager_pipeline", None)
        if pipeline is None:
            pipeline = dev._eager_pipeline = ArrowComputePipeline(dev, None)
        _note_foreign_buffers(dev, (self,) + args)
        out = op_fn(self, *args, pipeline)
        if pipeline._lazies:
            pipeline.flush_recorded()
        return out
    return run


def _check_same_len(a, b, what):
    if a.len != b.len:
        raise Panic(f"{what}: length mismatch {a.len} vs {b.len}")


# ==========================================================================================
# low-level launchers (one C-ABI call each).  These run once per op on columns of any size, so
# they touch the private fields directly (after launching a pending recorded chain) and call the
# library without the checking helper on the success path: ~5 us of interpreter time per op is
# the whole cost of a 1 Mi-row column op (BASELINE.json configs[0]).
# ==========================================================================================
def _ready(x):
    if x._lazy is not None:
        x._materialize()


def _alloc_validity(dev, n, va, vb=None):
    """output bitmap iff an input has one (null_bit_buffer.rs:168-204) -> (NullBitBufferGpu | None, ptr | None)"""
    if va is None and vb is None:
        return None, None
    buf = dev.create_empty_buffer(((n + 31) >> 5) * 4)
    return NullBitBufferGpu(buf, n, dev), buf.ptr


def _binary(op: int, a: PrimitiveArrayGpu, b: PrimitiveArrayGpu, out_cls=None, what="binary"):
    n = a.len
    if n != b.len:
        raise Panic(f"{what}: length mismatch {n} vs {b.len}")
    _ready(a), _ready(b)
    dev, va, vb = a.gpu_device, a._null_buffer, b._null_buffer
    out_cls = out_cls or type(a)
    nb, vout = _alloc_validity(dev, n, va, vb)
    data = dev.create_empty_buffer(n * out_cls.ITEMSIZE)
    rc = (_ffi._lib or lib()).agpu_binary(dev.handle, op, a.DTYPE, a._data.ptr, b._data.ptr, data.ptr, n,
                                           va.bit_buffer.ptr if va is not None else None,
                                           vb.bit_buffer.ptr if vb is not None else None, vout)
    if rc:
        check(rc, what)
    return out_cls(data, dev, n, nb)


def _scalar(op: int, a: PrimitiveArrayGpu, s: PrimitiveArrayGpu, what="scalar"):
    """rhs is a 1-element array; validity of `a` is copied (arithmetic/src/lib.rs:35-38)"""
    if s.len != 1:
        raise Panic(f"{what}: scalar operand must have exactly one element")
    _ready(a), _ready(s)
    dev, n, va = a.gpu_device, a.len, a._null_buffer
    nb, vout = _alloc_validity(dev, n, va)
    data = dev.create_empty_buffer(n * a.ITEMSIZE)
    rc = (_ffi._lib or lib()).agpu_scalar(dev.handle, op, a.DTYPE, a._data.ptr, s._data.ptr, data.ptr, n,
                                           va.bit_buffer.ptr if va is not None else None, vout)
    if rc:
        check(rc, what)
    return type(a)(data, dev, n, nb)


def _unary(op: int, a: PrimitiveArrayGpu, out_cls=None, what="unary"):
    _ready(a)
    dev, n, va = a.gpu_device, a.len, a._null_buffer
    out_cls = out_cls or type(a)
    nb, vout = _alloc_validity(dev, n, va)
    data = dev.create_empty_buffer(n * out_cls.ITEMSIZE)
    rc = (_ffi._lib or lib()).agpu_unary(dev.handle, op, a.DTYPE, a._data.ptr, data.ptr, n,
                                          va.bit_buffer.ptr if va is not None else None, vout)
    if rc:
        check(rc, what)
    return out_cls(data, dev, n, nb)


def _compare(op: int, a: PrimitiveArrayGpu, b: PrimitiveArrayGpu, what="compare") -> BooleanArrayGPU:
    n = a.len
    if n != b.len:
        raise Panic(f"{what}: length mismatch {n} vs {b.len}")
    _ready(a), _ready(b)
    dev, va, vb = a.gpu_device, a._null_buffer, b._null_buffer
    nb, vout = _alloc_validity(dev, n, va, vb)
    data = dev.create_empty_buffer(((n + 31) >> 5) * 4)
    rc = (_ffi._lib or lib()).agpu_compare(dev.handle, op, a.DTYPE, a._data.ptr, b._data.ptr, data.ptr, n,
                                            va.bit_buffer.ptr if va is not None else None,
                                            vb.bit_buffer.ptr if vb is not None else None, vout)
    if rc:
        check(rc, what)
    return BooleanArrayGPU(data, dev, n, nb)


def _shift(op: int, a: PrimitiveArrayGpu, counts: UInt32ArrayGPU, what="shift"):
    n = a.len
    if n != counts.len:
        raise Panic(f"{what}: length mismatch {n} vs {counts.len}")
    _ready(a), _ready(counts)
    dev, va, vb = a.gpu_device, a._null_buffer, counts._null_buffer
    nb, vout = _alloc_validity(dev, n, va, vb)
    data = dev.create_empty_buffer(n * a.ITEMSIZE)
    rc = (_ffi._lib or lib()).agpu_shift(dev.handle, op, a.DTYPE, a._data.ptr, counts._data.ptr, data.ptr, n,
                                          va.bit_buffer.ptr if va is not None else None,
                                          vb.bit_buffer.ptr if vb is not None else None, vout)
    if rc:
        check(rc, what)
    return type(a)(data, dev, n, nb)


# ==========================================================================================
# arithmetic  (crates/arithmetic/src/arithmetic_kernels.rs, lib.rs, aggregate_kernels.rs)
# ==========================================================================================
_SCALAR_OPS = {"add": _ffi.ADD, "sub": _ffi.SUB, "mul": _ffi.MUL, "div": _ffi.DIV, "rem": _ffi.REM}

# scalar + : f32, i32, u32, Date32, u16 ; - * / % : f32, i32, u32, Date32  (SURVEY.md §2.2).
# i8/u8/i16 (+ u16 - * / %) are new surface required by BASELINE.json config 2.
_SCALAR_TYPES = (Float32ArrayGPU, Int32ArrayGPU, UInt32ArrayGPU, Date32ArrayGPU, UInt16ArrayGPU,
                 Int16ArrayGPU, Int8ArrayGPU, UInt8ArrayGPU)


def _scalar_compatible(a, s) -> bool:
    if type(a) is type(s):
        return True
    return isinstance(a, _I32_BACKED) and isinstance(s, _I32_BACKED)  # T: Int32Type (types.rs)


def _make_scalar(name, op):
    def scalar_op(self, value, pipeline):
        if not isinstance(self, _SCALAR_TYPES) or not _scalar_compatible(self, value):
            raise Panic(f"Operation {name}_scalar not supported for type {self.get_dtype()} {value.get_dtype()}")
        return _scalar(op, self, value, f"{name}_scalar")
    return scalar_op


for _n, _o in _SCALAR_OPS.items():
    _fn = _make_scalar(_n, _o)
    setattr(PrimitiveArrayGpu, f"{_n}_scalar_op", _fn)        # ArrowScalarAdd::add_scalar_op ...
    setattr(PrimitiveArrayGpu, f"{_n}_scalar", _eager(_fn))    # ArrowScalarAdd::add_scalar ...

_ARRAY_OPS = {"add": _ffi.ADD, "sub": _ffi.SUB, "mul": _ffi.MUL, "div": _ffi.DIV}


def _array_result_cls(a, b):
    """i32 + Date32 pairs give the type of `self` (impl for T: Int32Type, arithmetic/src/i32.rs:103-119);
    add_array_dyn maps (Int32, Date32) to self.add_op, so the result is typed like the lhs."""
    return type(a)


def _make_array(name, op):
    def array_op(self, value, pipeline):
        same = type(self) is type(value) or (isinstance(self, _I32_BACKED) and isinstance(value, _I32_BACKED))
        if not same or isinstance(self, BooleanArrayGPU):
            raise Panic(f"Operation {name} not supported for type {self.get_dtype()} {value.get_dtype()}")
        return _binary(op, self, value, _array_result_cls(self, value), name)
    return array_op


for _n, _o in _ARRAY_OPS.items():
    _fn = _make_array(_n, _o)
    setattr(PrimitiveArrayGpu, f"{_n}_op", _fn)       # ArrowAdd::add_op ...
    setattr(PrimitiveArrayGpu, _n, _eager(_fn))        # ArrowAdd::add ...


def _neg_op(self, pipeline):
    if not isinstance(self, Float32ArrayGPU):
        raise Panic(f"Operation neg_dyn not supported for type {self.get_dtype()}")
    return _unary(_ffi.NEG, self, what="neg")


PrimitiveArrayGpu.neg_op = _neg_op
PrimitiveArrayGpu.neg = _eager(_neg_op)


def _sum_op(self, pipeline):
    """Sum::sum_op (aggregate_kernels.rs:24-52): one-element array, validity ignored"""
    if not isinstance(self, (Float32ArrayGPU, Int32ArrayGPU, UInt32ArrayGPU)):
        raise Panic(f"Operation sum not supported for type {self.get_dtype()}")
    dev = self.gpu_device
    out = type(self).empty(1, dev)
    check(lib().agpu_sum(dev.handle, self.DTYPE, self.data.ptr, self.len, out.data.ptr), "sum")
    return out


PrimitiveArrayGpu.sum_op = _sum_op
PrimitiveArrayGpu.sum = _eager(_sum_op)


def _dyn2(method_name, doc=""):
    """(fn_dyn, fn_op_dyn) pair dispatching to a two-operand method"""
    def op_dyn(data_1, data_2, pipeline):
        return getattr(data_1, method_name)(data_2, pipeline)

    def dyn(data_1, data_2):
        pipeline = _pipeline_for(data_1)
        out = op_dyn(data_1, data_2, pipeline)
        pipeline.finish()
        return out
    dyn.__doc__ = op_dyn.__doc__ = doc
    return dyn, op_dyn


def _dyn1(method_name, doc=""):
    def op_dyn(data, pipeline):
        return getattr(data, method_name)(pipeline)

    def dyn(data):
        pipeline = _pipeline_for(data)
        out = op_dyn(data, pipeline)
        pipeline.finish()
        return out
    dyn.__doc__ = op_dyn.__doc__ = doc
    return dyn, op_dyn


add_scalar_dyn, add_scalar_op_dyn = _dyn2("add_scalar_op", "column + one-element column (the reference's scalar form)")
sub_scalar_dyn, sub_scalar_op_dyn = _dyn2("sub_scalar_op", "column - one-element column")
mul_scalar_dyn, mul_scalar_op_dyn = _dyn2("mul_scalar_op", "column * one-element column")
div_scalar_dyn, div_scalar_op_dyn = _dyn2("div_scalar_op", "column / one-element column")
rem_scalar_dyn, rem_scalar_op_dyn = _dyn2("rem_scalar_op", "column % one-element column")
add_array_dyn, add_array_op_dyn = _dyn2("add_op", "x + y, row by row over both columns")
sub_array_dyn, sub_array_op_dyn = _dyn2("sub_op", "x - y, row by row over both columns")
mul_array_dyn, mul_array_op_dyn = _dyn2("mul_op", "x * y, row by row over both columns")
div_array_dyn, div_array_op_dyn = _dyn2("div_op", "x / y, row by row over both columns")
neg_dyn, neg_op_dyn = _dyn1("neg_op")


def _len_routed(array_op_dyn, scalar_op_dyn):
    """add_dyn & co route by operand length (arithmetic_kernels.rs:101-119): both len 1 or both
    != 1 -> array op; exactly one of length 1 -> scalar op with that operand as the scalar."""
    def op_dyn(input1, input2, pipeline):
        x, y = input1.len, input2.len
        if (x == 1 and y == 1) or (x != 1 and y != 1):
            return array_op_dyn(input1, input2, pipeline)
        if y == 1:
            return scalar_op_dyn(input1, input2, pipeline)
        return scalar_op_dyn(input2, input1, pipeline)

    def dyn(input1, input2):
        pipeline = _pipeline_for(input1)
        out = op_dyn(input1, input2, pipeline)
        pipeline.finish()
        return out
    return dyn, op_dyn


add_dyn, add_op_dyn = _len_routed(add_array_op_dyn, add_scalar_op_dyn)
sub_dyn, sub_op_dyn = _len_routed(sub_array_op_dyn, sub_scalar_op_dyn)
mul_dyn, mul_op_dyn = _len_routed(mul_array_op_dyn, mul_scalar_op_dyn)
div_dyn, div_op_dyn = _len_routed(div_array_op_dyn, div_scalar_op_dyn)

# ==========================================================================================
# compare  (crates/compare/src/lib.rs)
# ==========================================================================================
_CMP = {"gt": _ffi.GT, "gteq": _ffi.GTEQ, "lt": _ffi.LT, "lteq": _ffi.LTEQ, "eq": _ffi.EQ}


def _make_cmp(name, op):
    def cmp_op(self, operand, pipeline):
        if type(self) is not type(operand) or isinstance(self, BooleanArrayGPU):
            raise Panic(f"Operation {name}_dyn not supported for type {self.get_dtype()} {operand.get_dtype()}")
        return _compare(op, self, operand, name)
    return cmp_op


for _n, _o in _CMP.items():
    _fn = _make_cmp(_n, _o)
    setattr(PrimitiveArrayGpu, f"{_n}_op", _fn)     # Compare::gt_op ...
    setattr(PrimitiveArrayGpu, _n, _eager(_fn))      # Compare::gt ...

gt_dyn, gt_op_dyn = _dyn2("gt_op", "row-wise predicate x > y -> BooleanArrayGPU")
gteq_dyn, gteq_op_dyn = _dyn2("gteq_op", "row-wise predicate x >= y -> BooleanArrayGPU")
lt_dyn, lt_op_dyn = _dyn2("lt_op", "row-wise predicate x < y -> BooleanArrayGPU")
lteq_dyn, lteq_op_dyn = _dyn2("lteq_op", "row-wise predicate x <= y -> BooleanArrayGPU")
eq_dyn, eq_op_dyn = _dyn2("eq_op", "row-wise predicate x == y -> BooleanArrayGPU")


def _make_minmax(name, op):
    def mm_op(self, operand, pipeline):
        if type(self) is not type(operand) or isinstance(self, BooleanArrayGPU):
            raise Panic(f"Operation {name}_dyn not supported for type {self.get_dtype()} {operand.get_dtype()}")
        return _binary(op, self, operand, what=name)
    return mm_op


for _n, _o in {"min": _ffi.MIN, "max": _ffi.MAX}.items():
    _fn = _make_minmax(_n, _o)
    setattr(PrimitiveArrayGpu, f"{_n}_op", _fn)     # MinMax::min_op / max_op
    setattr(PrimitiveArrayGpu, _n, _eager(_fn))

min_dyn, min_op_dyn = _dyn2("min_op", "min(x, y), row by row over both columns")
max_dyn, max_op_dyn = _dyn2("max_op", "max(x, y), row by row over both columns")

# ==========================================================================================
# logical  (crates/logical/src/lib.rs, boolean.rs)
# ==========================================================================================
_LOGICAL = {"bitwise_and": _ffi.AND, "bitwise_or": _ffi.OR, "bitwise_xor": _ffi.XOR}


def _make_logical(name, op):
    def logical_op(self, operand, pipeline):
        if type(self) is not type(operand) or not isinstance(self, _INT_TYPES):
            raise Panic(f"Operation {name}_dyn not supported for type {self.get_dtype()} {operand.get_dtype()}")
        return _binary(op, self, operand, what=name)
    return logical_op


for _n, _o in _LOGICAL.items():
    _fn = _make_logical(_n, _o)
    setattr(PrimitiveArrayGpu, f"{_n}_op", _fn)
    setattr(PrimitiveArrayGpu, _n, _eager(_fn))


def _not_op(self, pipeline):
    if not isinstance(self, _INT_TYPES):
        raise Panic(f"Operation bitwise_not_dyn not supported for type {self.get_dtype()}")
    return _unary(_ffi.NOT, self, what="bitwise_not")


PrimitiveArrayGpu.bitwise_not_op = _not_op
PrimitiveArrayGpu.bitwise_not = _eager(_not_op)


def _make_shift(name, op):
    def shift_op(self, operand, pipeline):
        if not isinstance(self, _INT_TYPES) or not isinstance(operand, UInt32ArrayGPU):
            raise Panic(f"Operation {name}_dyn not supported for type {self.get_dtype()} {operand.get_dtype()}")
        return _shift(op, self, operand, name)
    return shift_op


for _n, _o in {"bitwise_shl": _ffi.SHL, "bitwise_shr": _ffi.SHR}.items():
    _fn = _make_shift(_n, _o)
    setattr(PrimitiveArrayGpu, f"{_n}_op", _fn)
    setattr(PrimitiveArrayGpu, _n, _eager(_fn))


def _make_bool_logical(name, op):
    def bool_op(self, operand, pipeline):
        if not isinstance(operand, BooleanArrayGPU):
            raise Panic(f"Operation {name}_dyn not supported for type {self.get_dtype()} {operand.get_dtype()}")
        _check_same_len(self, operand, name)
        dev = self.gpu_device
        nb = _new_validity(dev, self.len, self.null_buffer, operand.null_buffer)
        out = BooleanArrayGPU.empty(self.len, dev, nb)
        check(lib().agpu_bitmap_binary(dev.handle, op, self.data.ptr, operand.data.ptr, out.data.ptr, self.len,
                                       _vptr(self.null_buffer), _vptr(operand.null_buffer), _vptr(nb)), name)
        return out
    return bool_op


for _n, _o in _LOGICAL.items():
    _fn = _make_bool_logical(_n, _o)
    setattr(BooleanArrayGPU, f"{_n}_op", _fn)
    setattr(BooleanArrayGPU, _n, _eager(_fn))


def _bool_not_op(self, pipeline):
    dev = self.gpu_device
    nb = _new_validity(dev, self.len, self.null_buffer)
    out = BooleanArrayGPU.empty(self.len, dev, nb)
    check(lib().agpu_bitmap_not(dev.handle, self.data.ptr, out.data.ptr, self.len, _vptr(self.null_buffer),
                                _vptr(nb)), "bitwise_not")
    return out


BooleanArrayGPU.bitwise_not_op = _bool_not_op
BooleanArrayGPU.bitwise_not = _eager(_bool_not_op)


def _bool_shift_unsupported(self, operand, pipeline=None):
    raise Panic("shift is not defined for BooleanArrayGPU (empty shader, logical/src/boolean.rs:14)")


BooleanArrayGPU.bitwise_shl_op = BooleanArrayGPU.bitwise_shr_op = _bool_shift_unsupported
BooleanArrayGPU.bitwise_shl = BooleanArrayGPU.bitwise_shr = _bool_shift_unsupported


def _bool_reduce(fn_name):
    def run(self) -> bool:
        """LogicalContains (logical/src/boolean.rs:106-147); validity is ignored like the reference"""
        dev = self.gpu_device
        flag = dev.create_empty_buffer(4)
        check(getattr(lib(), fn_name)(dev.handle, self.data.ptr, self.len, flag.ptr), fn_name)
        return bool(dev.retrive_data(flag, 4).view(np.uint32)[0])
    return run


BooleanArrayGPU.any = _bool_reduce("agpu_any")
BooleanArrayGPU.all = _bool_reduce("agpu_all")

bitwise_and_dyn, bitwise_and_op_dyn = _dyn2("bitwise_and_op", "x & y, row by row over both columns")
bitwise_or_dyn, bitwise_or_op_dyn = _dyn2("bitwise_or_op", "x | y, row by row over both columns")
bitwise_xor_dyn, bitwise_xor_op_dyn = _dyn2("bitwise_xor_op", "x ^ y, row by row over both columns")
bitwise_shl_dyn, bitwise_shl_op_dyn = _dyn2("bitwise_shl_op", "x << y, row by row over both columns")
bitwise_shr_dyn, bitwise_shr_op_dyn = _dyn2("bitwise_shr_op", "x >> y, row by row over both columns")
bitwise_not_dyn, bitwise_not_op_dyn = _dyn1("bitwise_not_op", "!x of every row")

# ==========================================================================================
# cast  (crates/cast/src/lib.rs)
# ==========================================================================================
# the matrix of cast_dyn (cast/src/lib.rs:135-161)
_CAST_MATRIX = {
    Int8ArrayGPU: (UInt8ArrayGPU, UInt16ArrayGPU, UInt32ArrayGPU, Int16ArrayGPU, Int32ArrayGPU, Float32ArrayGPU),
    Int16ArrayGPU: (Int32ArrayGPU, UInt16ArrayGPU, UInt32ArrayGPU, Float32ArrayGPU),
    UInt8ArrayGPU: (UInt16ArrayGPU, UInt32ArrayGPU, Int8ArrayGPU, Int16ArrayGPU, Int32ArrayGPU, Float32ArrayGPU),
    UInt16ArrayGPU: (UInt32ArrayGPU, Int16ArrayGPU, Int32ArrayGPU, Float32ArrayGPU),
    Float32ArrayGPU: (UInt8ArrayGPU,),
    BooleanArrayGPU: (Float32ArrayGPU,),
}


def _cast_to(self, into_cls, pipeline, matrix=_CAST_MATRIX, what="cast"):
    if into_cls not in matrix.get(type(self), ()):
        raise Panic(f"Casting not supported for type {self.get_dtype()} {into_cls.ARROW_TYPE}")
    dev = self.gpu_device
    nb = _new_validity(dev, self.len, self.null_buffer)
    out = into_cls.empty(self.len, dev, nb)
    check(lib().agpu_cast(dev.handle, self.DTYPE, into_cls.DTYPE, self.data.ptr, out.data.ptr, self.len,
                          _vptr(self.null_buffer), _vptr(nb)), what)
    return out


def _cast_op(self, into, pipeline):
    """Cast<T>::cast_op; `into` is the target array class or its ArrowType"""
    return _cast_to(self, ARRAY_TYPES[into] if isinstance(into, ArrowType) else into, pipeline)


def _bitcast_op(self, into, pipeline):
    """BitCast<T>::bitcast_op — only u32 -> f32 exists (cast/src/lib.rs:187-192)"""
    cls = ARRAY_TYPES[into] if isinstance(into, ArrowType) else into
    return _cast_to(self, cls, pipeline, {UInt32ArrayGPU: (Float32ArrayGPU,)}, "bitcast")


for _cls in (PrimitiveArrayGpu, BooleanArrayGPU):
    _cls.cast_op = _cast_op
    _cls.cast = _eager(_cast_op)
PrimitiveArrayGpu.bitcast_op = _bitcast_op
PrimitiveArrayGpu.bitcast = _eager(_bitcast_op)

cast_dyn, cast_op_dyn = _dyn2("cast_op", "every row converted to the element type `T`")
bitcast_dyn, bitcast_op_dyn = _dyn2("bitcast_op", "the same bits of every row read as `T`")

# ==========================================================================================
# math  (crates/math/src/lib.rs)   trigonometry  (crates/trigonometry/src/lib.rs)
# ==========================================================================================
_FLOAT_UNARY = {"sqrt": _ffi.SQRT, "cbrt": _ffi.CBRT, "exp": _ffi.EXP, "exp2": _ffi.EXP2, "log": _ffi.LOG,
                "log2": _ffi.LOG2}


def _make_float_unary(name, op):
    def fn(self, pipeline):
        if not isinstance(self, Float32ArrayGPU):
            raise Panic(f"Operation {name}_dyn not supported for type {self.get_dtype()}")
        return _unary(op, self, what=name)
    return fn


for _n, _o in _FLOAT_UNARY.items():
    _fn = _make_float_unary(_n, _o)
    setattr(PrimitiveArrayGpu, f"{_n}_op", _fn)    # FloatMathUnary::sqrt_op ...
    setattr(PrimitiveArrayGpu, _n, _eager(_fn))


def _abs_op(self, pipeline):
    if not isinstance(self, (Float32ArrayGPU, Int32ArrayGPU)):
        raise Panic(f"Operation abs_dyn not supported for type {self.get_dtype()}")
    return _unary(_ffi.ABS, self, what="abs")


def _power_op(self, other, pipeline):
    if type(self) is not type(other) or not isinstance(self, (Float32ArrayGPU, Int32ArrayGPU)):
        raise Panic(f"Operation power_dyn not supported for type {self.get_dtype()} {other.get_dtype()}")
    return _binary(_ffi.POW, self, other, what="power")


PrimitiveArrayGpu.abs_op = _abs_op
PrimitiveArrayGpu.abs = _eager(_abs_op)
PrimitiveArrayGpu.power_op = _power_op
PrimitiveArrayGpu.power = _eager(_power_op)

abs_dyn, abs_op_dyn = _dyn1("abs_op", "abs(x) of every row")
sqrt_dyn, sqrt_op_dyn = _dyn1("sqrt_op", "square_root(x) of every row")
cbrt_dyn, cbrt_op_dyn = _dyn1("cbrt_op", "cube_root(x) of every row")
exp_dyn, exp_op_dyn = _dyn1("exp_op", "e^x of every row")
exp2_dyn, exp2_op_dyn = _dyn1("exp2_op", "2^x of every row")
log_dyn, log_op_dyn = _dyn1("log_op", "log(x) of every row")
log2_dyn, log2_op_dyn = _dyn1("log2_op", "log_to_base_2(x) of every row")
power_dyn, power_op_dyn = _dyn2("power_op", "x ^ y, row by row over both columns")

_TRIG_INT = (Int8ArrayGPU, UInt8ArrayGPU, Int16ArrayGPU, UInt16ArrayGPU)


def _make_trig(name, op, allow_int):
    def fn(self, pipeline):
        ok = isinstance(self, Float32ArrayGPU) or (allow_int and isinstance(self, _TRIG_INT))
        if not ok:
            raise Panic(f"Operation {name}_op_dyn not supported for type {self.get_dtype()}")
        return _unary(op, self, Float32ArrayGPU, name)   # int columns: cast fused into the kernel
    return fn


for _n, _o, _ai in (("sin", _ffi.SIN, True), ("cos", _ffi.COS, True), ("acos", _ffi.ACOS, False),
                    ("sinh", _ffi.SINH, True)):
    _fn = _make_trig(_n, _o, _ai)
    setattr(PrimitiveArrayGpu, f"{_n}_op", _fn)    # Trigonometric::sin_op ..., Hyperbolic::sinh_op
    setattr(PrimitiveArrayGpu, _n, _eager(_fn))

sin_dyn, sin_op_dyn = _dyn1("sin_op", "sin(x) of every row")
cos_dyn, cos_op_dyn = _dyn1("cos_op", "cos(x) of every row")
acos_dyn, acos_op_dyn = _dyn1("acos_op", "acos(x) of every row")
sinh_dyn, sinh_op_dyn = _dyn1("sinh_op", "sinh(x) of every row")

# ==========================================================================================
# routines  (crates/routines/src/lib.rs, merge.rs, take.rs, put.rs, bool.rs)
# ==========================================================================================
_TAKE_PUT_TYPES = (Date32ArrayGPU, UInt32ArrayGPU, Int32ArrayGPU, Float32ArrayGPU, BooleanArrayGPU,
                   # 8/16-bit take/put are `todo!()` in the reference (routines/src/i16.rs:5-6): new surface
                   Int16ArrayGPU, UInt16ArrayGPU, Int8ArrayGPU, UInt8ArrayGPU)


def _merge_op(self, other, mask, pipeline):
    """Swizzle::merge_op (routines/src/lib.rs:82-120, bool.rs:49-87) + merge_null_buffers_op"""
    if type(self) is not type(other) or not isinstance(mask, BooleanArrayGPU):
        raise Panic(f"Merge Operation not supported between {self.get_dtype()} and {other.get_dtype()}")
    _check_same_len(self, other, "merge")
    _check_same_len(self, mask, "merge")
    dev = self.gpu_device
    nb = _new_validity(dev, self.len, self.null_buffer, other.null_buffer, mask.null_buffer)
    out = type(self).empty(self.len, dev, nb)
    check(lib().agpu_merge(dev.handle, self.DTYPE, self.data.ptr, other.data.ptr, mask.data.ptr, out.data.ptr,
                           self.len, _vptr(self.null_buffer), _vptr(other.null_buffer),
                           _vptr(mask.null_buffer), _vptr(nb)), "merge")
    return out


def _take_op(self, indexes, pipeline):
    """Swizzle::take_op (lib.rs:122-143, bool.rs:89-100); result length = indexes.len (Q8)"""
    if not isinstance(self, _TAKE_PUT_TYPES) or not isinstance(indexes, UInt32ArrayGPU):
        raise Panic(f"Take Operation not supported for {self.get_dtype()}")
    dev = self.gpu_device
    nb = _new_validity(dev, indexes.len, self.null_buffer)
    out = type(self).empty(indexes.len, dev, nb)
    check(lib().agpu_take(dev.handle, self.DTYPE, self.data.ptr, self.len, indexes.data.ptr, out.data.ptr,
                          indexes.len, _vptr(self.null_buffer), _vptr(nb)), "take")
    return out


def _put_op(self, src_indexes, dst, dst_indexes, pipeline):
    """Swizzle::put_op (lib.rs:145-170): scatter into `dst` in place; arrays with validity are
    `todo!()` in the reference and rejected here"""
    if type(self) is not type(dst) or not isinstance(self, _TAKE_PUT_TYPES):
        raise Panic(f"Put Operation not supported for {self.get_dtype()} and {dst.get_dtype()}")
    if self.null_buffer is not None or dst.null_buffer is not None:
        raise NotImplementedError("put with validity bitmaps is todo!() in the reference (routines/src/lib.rs:164-169)")
    _check_same_len(src_indexes, dst_indexes, "put")
    dev = self.gpu_device
    check(lib().agpu_put(dev.handle, self.DTYPE, self.data.ptr, self.len, src_indexes.data.ptr, dst.data.ptr, dst.len,
                         dst_indexes.data.ptr, src_indexes.len), "put")


class FilterPlan:
    """state between the two halves of a filter: per-tile counts / scanned offsets (scratch) and
    the device-side total.  A sharded caller exchanges totals between the halves."""

    def __init__(self, array, mask, scratch, total):
        self.array, self.mask, self.scratch, self.total = array, mask, scratch, total


def _filter_count_op(self, mask, pipeline, total_ptr=None, post=None) -> FilterPlan:
    """first half of filter: count the selected rows (one pass over the mask bits).  `total_ptr`
    may point at caller-owned device memory (8 bytes) that receives the count.  `post` =
    (peer slot pointers, rank, world, seq) of a sharded.CountExchange: the count kernel itself then
    stores this shard's total into every peer's slot area (agpu_filter_count_post)."""
    if not isinstance(mask, BooleanArrayGPU) or isinstance(self, BooleanArrayGPU):
        raise Panic(f"Filter Operation not supported for {self.get_dtype()}")
    _check_same_len(self, mask, "filter")
    dev = self.gpu_device
    l = lib()
    scratch = dev.create_empty_buffer(l.agpu_filter_scratch_bytes(self.len))
    total = None if total_ptr is not None else dev.create_empty_buffer(8)
    tptr = total_ptr if total_ptr is not None else total.ptr
    if post is None:
        check(l.agpu_filter_count(dev.handle, mask.data.ptr, _vptr(mask.null_buffer), self.len, scratch.ptr, tptr), "filter_count")
    else:
        ptrs, rank, world, seq = post
        check(l.agpu_filter_count_post(dev.handle, mask.data.ptr, _vptr(mask.null_buffer), self.len, scratch.ptr, tptr, ptrs, rank,
                                       world, seq), "filter_count_post")
    return FilterPlan(self, mask, scratch, total)


def _filter_scatter_op(self, plan: FilterPlan, count: int, pipeline):
    """second half of filter: compact values (and validity) into a `count`-row array.  `count` may
    be an upper bound on the selected rows (a sharded caller launches this before the exact total
    has reached the host): buffers are sized for it, rows beyond it are dropped by the kernel."""
    dev = self.gpu_device
    mask = plan.mask
    out = type(self).empty(count, dev)
    vout = None
    if self.null_buffer is not None:
        vout = dev.create_empty_buffer(bitmap_words(count) * 4)
        out.null_buffer = NullBitBufferGpu(vout, count, dev)
    check(lib().agpu_filter_scatter(dev.handle, self.DTYPE, self.data.ptr, _vptr(self.null_buffer), mask.data.ptr,
                                    _vptr(mask.null_buffer), self.len, plan.scratch.ptr, out.data.ptr,
                                    vout.ptr if vout else None, count), "filter_scatter")
    return out


def _filter_op(self, mask, pipeline):
    """New surface (BASELINE.json config 5; the reference has no filter): keep rows whose mask
    bit is set and valid, order preserving.  Needs the selected count on the host to size the
    output — the one extra synchronisation of this op."""
    plan = _filter_count_op(self, mask, pipeline)
    count = int(self.gpu_device.retrive_data(plan.total, 8).view(np.uint64)[0])
    return _filter_scatter_op(self, plan, count, pipeline)


for _cls in (PrimitiveArrayGpu, BooleanArrayGPU):
    _cls.merge_op = _merge_op
    _cls.merge = _eager(_merge_op)
    _cls.take_op = _take_op
    _cls.take = _eager(_take_op)
    _cls.put_op = _put_op
    _cls.put = _eager(_put_op)
PrimitiveArrayGpu.filter_op = _filter_op
PrimitiveArrayGpu.filter = _eager(_filter_op)
PrimitiveArrayGpu.filter_count_op = _filter_count_op
PrimitiveArrayGpu.filter_scatter_op = _filter_scatter_op


def merge_op_dyn(operand_1, operand_2, mask, pipeline):
    return operand_1.merge_op(operand_2, mask, pipeline)


def merge_dyn(operand_1, operand_2, mask):
    pipeline = ArrowComputePipeline(operand_1.get_gpu_device(), "merge")
    out = merge_op_dyn(operand_1, operand_2, mask, pipeline)
    pipeline.finish()
    return out


def take_op_dyn(operand_1, indexes, pipeline):
    return operand_1.take_op(indexes, pipeline)


def take_dyn(operand_1, indexes):
    pipeline = ArrowComputePipeline(operand_1.get_gpu_device(), "take")
    out = take_op_dyn(operand_1, indexes, pipeline)
    pipeline.finish()
    return out


def put_op_dyn(src, src_indexes, dst, dst_indexes, pipeline):
    src.put_op(src_indexes, dst, dst_indexes, pipeline)


def put_dyn(src, src_indexes, dst, dst_indexes):
    pipeline = ArrowComputePipeline(src.get_gpu_device(), "put")
    put_op_dyn(src, src_indexes, dst, dst_indexes, pipeline)
    pipeline.finish()


def filter_op_dyn(operand_1, mask, pipeline):
    return operand_1.filter_op(mask, pipeline)


def filter_dyn(operand_1, mask):
    pipeline = ArrowComputePipeline(operand_1.get_gpu_device(), "filter")
    out = filter_op_dyn(operand_1, mask, pipeline)
    pipeline.finish()
    return out


# ==========================================================================================
# fused expression (BASELINE.json config 3):  ((a * b) + c) > d  in one pass
# ==========================================================================================
def fused_mul_add_gt_op(a, b, c, d, pipeline) -> BooleanArrayGPU:
    """Equivalent to gt_op_dyn(add_op_dyn(mul_op_dyn(a, b), c), d) on one pipeline
    (the recorded-chain pattern of crates/arrow/examples/simple.rs:45-72), bit-identical to it,
    in a single kernel: 16.75 B/row instead of 33.25 B/row."""
    for x in (a, b, c, d):
        if not isinstance(x, Float32ArrayGPU):
            raise Panic(f"fused_mul_add_gt not supported for type {x.get_dtype()}")
        _check_same_len(a, x, "fused_mul_add_gt")
    dev = a.gpu_device
    _note_foreign_buffers(dev, (a, b, c, d))
    nb = _new_validity(dev, a.len, a.null_buffer, b.null_buffer, c.null_buffer, d.null_buffer)
    out = BooleanArrayGPU.empty(a.len, dev, nb)
    check(lib().agpu_fused_mul_add_gt(dev.handle, a.data.ptr, b.data.ptr, c.data.ptr, d.data.ptr, out.data.ptr,
                                      a.len, _vptr(a.null_buffer), _vptr(b.null_buffer), _vptr(c.null_buffer),
                                      _vptr(d.null_buffer), _vptr(nb)), "fused_mul_add_gt")
    return out


def fused_mul_add_gt(a, b, c, d) -> BooleanArrayGPU:
    pipeline = _pipeline_for(a)
    out = fused_mul_add_gt_op(a, b, c, d, pipeline)
    pipeline.finish()
    return out


# ==========================================================================================
# general fused linear chains (north_star (2)): one kernel for  cast -> op -> op -> ... [-> compare]
# ==========================================================================================
_CHAIN_UNARY = {"neg": _ffi.NEG, "abs": _ffi.ABS, "sqrt": _ffi.SQRT, "cbrt": _ffi.CBRT, "exp": _ffi.EXP,
                "exp2": _ffi.EXP2, "log": _ffi.LOG, "log2": _ffi.LOG2, "sin": _ffi.SIN, "cos": _ffi.COS,
                "acos": _ffi.ACOS, "sinh": _ffi.SINH}
_CHAIN_BINARY = {"add": _ffi.ADD, "sub": _ffi.SUB, "mul": _ffi.MUL, "div": _ffi.DIV, "rem": _ffi.REM,
                 "min": _ffi.MIN, "max": _ffi.MAX, "power": _ffi.POW}
_CHAIN_COMPARE = {"gt": _ffi.GT, "gteq": _ffi.GTEQ, "lt": _ffi.LT, "lteq": _ffi.LTEQ, "eq": _ffi.EQ}
_CHAIN_INPUT = (Float32ArrayGPU, Int8ArrayGPU, UInt8ArrayGPU, Int16ArrayGPU, UInt16ArrayGPU)


class DeviceScalar:
    """a ONE-element array used as the scalar operand of a chain step (the reference passes
    scalars this way); read on the device, never copied back to the host.  Float32ArrayGPU for
    fused_chain, the column's own type for fused_chain_int."""

    def __init__(self, array):
        if not isinstance(array, PrimitiveArrayGpu) or array.len != 1:
            raise Panic("DeviceScalar needs a one-element array")
        self.array = array


def _encode_chain_steps(data, steps, arr, at, validities, what="fused_chain"):
    """fills arr[at : at + len(steps)] with the agpu_chain_step form of `steps`; appends the operand
    columns' validity buffers to `validities`; returns True when the last step is a compare"""
    is_pred = False
    for k, step in enumerate(steps):
        name, operand = step[0], (step[1] if len(step) > 1 else None)
        slot = arr[at + k]
        if name in _CHAIN_UNARY and operand is None:
            slot.kind, slot.op = _ffi.STEP_UNARY, _CHAIN_UNARY[name]
            continue
        if name in _CHAIN_BINARY:
            op, kinds = _CHAIN_BINARY[name], (_ffi.STEP_BINARY_COLUMN, _ffi.STEP_BINARY_SCALAR, _ffi.STEP_BINARY_DEVSCALAR)
        elif name in _CHAIN_COMPARE:
            if k != len(steps) - 1:
                raise Panic("a compare can only end a fused chain")
            op, kinds, is_pred = (_CHAIN_COMPARE[name],
                                  (_ffi.STEP_COMPARE_COLUMN, _ffi.STEP_COMPARE_SCALAR, _ffi.STEP_COMPARE_DEVSCALAR), True)
        else:
            raise Panic(f"{what}: unknown step {step!r}")
        slot.op = op
        if isinstance(operand, DeviceScalar):
            if not isinstance(operand.array, Float32ArrayGPU):
                raise Panic(f"{what}: a device scalar must be a one-element Float32ArrayGPU")
            slot.kind, slot.operand = kinds[2], operand.array.data.ptr
        elif isinstance(operand, Float32ArrayGPU):
            _check_same_len(data, operand, what)
            slot.kind, slot.operand, slot.validity = kinds[0], operand.data.ptr, _vptr(operand.null_buffer)
            validities.append(operand.null_buffer)
        elif isinstance(operand, (int, float, np.floating, np.integer)):
            slot.kind, slot.scalar = kinds[1], float(operand)
        else:
            raise Panic(f"{what}: operand of {name!r} must be a Float32ArrayGPU or a number")
    return is_pred


def fused_chain_op(data, steps, pipeline):
    """Evaluate a linear chain in ONE kernel.  `steps` is a list of
         ("sqrt",)                 unary f32 op on the running value
         ("mul", other)            binary op with a Float32ArrayGPU column or a python float
         ("gt", other)             compare (only as the last step) -> BooleanArrayGPU
    The running value starts as cast<f32>(data) (f32, i8, u8, i16 or u16 column).  The result is
    bit-identical to applying the same `*_op`s one after another on a pipeline; validity is the AND
    of all columns' bitmaps.  Example: fused_chain(a, [("mul", b), ("add", c), ("gt", d)])."""
    if not isinstance(data, _CHAIN_INPUT):
        raise Panic(f"fused_chain not supported for type {data.get_dtype()}")
    if not 1 <= len(steps) <= _ffi.CHAIN_MAX_STEPS:
        raise Panic(f"fused_chain takes 1..{_ffi.CHAIN_MAX_STEPS} steps")
    arr = (_ffi.ChainStep * len(steps))()
    validities = [data.null_buffer]
    is_pred = _encode_chain_steps(data, steps, arr, 0, validities)
    dev = data.gpu_device
    _note_foreign_buffers(dev, [data] + [st[1] for st in steps if len(st) > 1 and isinstance(st[1], PrimitiveArrayGpu)])
    nb = _new_validity(dev, data.len, *validities)
    out = (BooleanArrayGPU if is_pred else Float32ArrayGPU).empty(data.len, dev, nb)
    check(lib().agpu_fused_chain(dev.handle, data.DTYPE, data.data.ptr, _vptr(data.null_buffer), arr, len(steps),
                                 out.data.ptr, data.len, _vptr(nb)), "fused_chain")
    return out


# ---- two results of one source column in one pass (agpu_fused_chain_pair) ----
_PAIR_STEPS = {"neg", "abs", "sqrt", "add", "sub", "mul", "div", "rem", "min", "max"} | set(_CHAIN_COMPARE)


def _step_columns(steps):
    return [st[1] for st in steps if len(st) > 1 and isinstance(st[1], PrimitiveArrayGpu)]


def _validity_set(data, steps):
    """the device bitmaps a chain's validity is the AND of (by address)"""
    return {nb.bit_buffer.ptr for nb in [data.null_buffer] + [c.null_buffer for c in _step_columns(steps)] if nb is not None}


def pair_eligible(data, value_steps, pred_steps) -> bool:
    """can `value_steps` and `pred_steps` (both starting at `data`) run as ONE agpu_fused_chain_pair
    kernel with results identical to the two chains run separately?"""
    if not isinstance(data, Float32ArrayGPU) or not value_steps or not pred_steps:
        return False
    if len(value_steps) + len(pred_steps) + 2 > _ffi.CHAIN_MAX_STEPS:
        return False
    if any(st[0] not in _PAIR_STEPS for st in value_steps + pred_steps):
        return False
    if any(st[0] in _CHAIN_COMPARE for st in value_steps) or pred_steps[-1][0] not in _CHAIN_COMPARE:
        return False
    if any(st[0] in _CHAIN_COMPARE for st in pred_steps[:-1]):
        return False
    cols = _step_columns(value_steps + pred_steps)
    if any(not isinstance(c, Float32ArrayGPU) or c.len != data.len for c in cols):
        return False
    if len({c.data.ptr for c in cols}) > _MAX_CHAIN_COLS:
        return False
    # the kernel writes ONE validity bitmap (AND of every input's): right for both results only
    # when both chains depend on the same bitmaps
    return _validity_set(data, value_steps) == _validity_set(data, pred_steps)


def fused_chain_pair_op(data, value_steps, pred_steps, pipeline):
    """`fused_chain_op(data, value_steps)` and `fused_chain_op(data, pred_steps)` in ONE kernel: the
    source and every operand column are read once, e.g. the first benchmark program of the
    reference, s = a + b; g = a > b:  fused_chain_pair(a, [("add", b)], [("gt", b)]) -> (s, g).
    f32 columns, arithmetic steps (neg abs sqrt + - * / % min max) and a closing compare in the
    second chain; both chains must depend on the same validity bitmaps (`pair_eligible`).  The two
    results share one validity buffer (bitmaps are never modified in place)."""
    if not pair_eligible(data, value_steps, pred_steps):
        raise Panic("fused_chain_pair: the two chains cannot share one kernel (see pair_eligible)")
    nv, npred = len(value_steps), len(pred_steps)
    arr = (_ffi.ChainStep * (nv + npred + 2))()
    validities = [data.null_buffer]
    _encode_chain_steps(data, value_steps, arr, 0, validities, "fused_chain_pair")
    arr[nv].kind, arr[nv + 1].kind = _ffi.STEP_STORE, _ffi.STEP_RESET
    _encode_chain_steps(data, pred_steps, arr, nv + 2, validities, "fused_chain_pair")
    dev = data.gpu_device
    _note_foreign_buffers(dev, [data] + _step_columns(value_steps + pred_steps))
    n = data.len
    vbuf = dev.create_empty_buffer(bitmap_words(n) * 4) if any(v is not None for v in validities) else None
    value = Float32ArrayGPU.empty(n, dev, NullBitBufferGpu(vbuf, n, dev) if vbuf is not None else None)
    pred = BooleanArrayGPU.empty(n, dev, NullBitBufferGpu(vbuf, n, dev) if vbuf is not None else None)
    check(lib().agpu_fused_chain_pair(dev.handle, data.DTYPE, data.data.ptr, _vptr(data.null_buffer), arr, len(arr),
                                      value.data.ptr, pred.data.ptr, n, vbuf.ptr if vbuf is not None else None),
          "fused_chain_pair")
    return value, pred


def fused_chain_pair(data, value_steps, pred_steps):
    pipeline = _pipeline_for(data)
    out = fused_chain_pair_op(data, value_steps, pred_steps, pipeline)
    pipeline.finish()
    return out


def fused_chain(data, steps):
    pipeline = _pipeline_for(data)
    out = fused_chain_op(data, steps, pipeline)
    pipeline.finish()
    return out


_INT_CHAIN_UNARY = {"bitwise_not": _ffi.NOT, "abs": _ffi.ABS}
_INT_CHAIN_SHIFT = {"bitwise_shl": _ffi.SHL, "bitwise_shr": _ffi.SHR}
_INT_CHAIN_BINARY = {"add": _ffi.ADD, "sub": _ffi.SUB, "mul": _ffi.MUL, "div": _ffi.DIV, "rem": _ffi.REM,
                     "min": _ffi.MIN, "max": _ffi.MAX, "bitwise_and": _ffi.AND, "bitwise_or": _ffi.OR,
                     "bitwise_xor": _ffi.XOR, "power": _ffi.POW}


def fused_chain_int_op(data, steps, pipeline):
    """fused_chain_op on an INTEGER column: the running value, the operand columns and the scalars
    all have the type of `data`; each step is the stand-alone integer kernel of that op (wrap in
    the column's width, x/0 = x, x%0 = 0).  Steps:
         ("bitwise_not",)  ("abs",)  [abs: Int32 only]
         ("add", other) ... sub mul div rem min max bitwise_and bitwise_or bitwise_xor power [Int32]
         ("bitwise_shl", counts)  ("bitwise_shr", counts)   counts = UInt32ArrayGPU, one shift per chain
         ("gt", other) ... gteq lt lteq eq          (only as the last step) -> BooleanArrayGPU
    `other` = a column of the same type and length, a DeviceScalar / one-element array of the same
    type, or a python int (uploaded as a one-element array)."""
    if not isinstance(data, _INT_TYPES + (Date32ArrayGPU,)):
        raise Panic(f"fused_chain_int not supported for type {data.get_dtype()}")
    if not 1 <= len(steps) <= _ffi.CHAIN_MAX_STEPS:
        raise Panic(f"fused_chain_int takes 1..{_ffi.CHAIN_MAX_STEPS} steps")
    arr = (_ffi.ChainStep * len(steps))()
    validities = [data.null_buffer]
    keep = []           # one-element arrays made here must outlive the launch call
    is_pred = False
    for k, step in enumerate(steps):
        name, operand = step[0], (step[1] if len(step) > 1 else None)
        if name in ("abs", "power") and not isinstance(data, _I32_BACKED):
            raise Panic(f"fused_chain_int: {name} not supported for type {data.get_dtype()}")   # math/src/i32.rs
        if name in _INT_CHAIN_UNARY and operand is None:
            arr[k].kind, arr[k].op = _ffi.STEP_UNARY, _INT_CHAIN_UNARY[name]
            continue
        if name in _INT_CHAIN_SHIFT:
            if not isinstance(operand, UInt32ArrayGPU) or isinstance(data, Date32ArrayGPU):
                raise Panic(f"fused_chain_int: {name} takes a UInt32ArrayGPU of per-row counts")
            _check_same_len(data, operand, "fused_chain_int")
            arr[k].kind, arr[k].op = _ffi.STEP_SHIFT_COLUMN, _INT_CHAIN_SHIFT[name]
            arr[k].operand, arr[k].validity = operand.data.ptr, _vptr(operand.null_buffer)
            validities.append(operand.null_buffer)
            continue
        if name in _INT_CHAIN_BINARY:
            op, kinds = _INT_CHAIN_BINARY[name], (_ffi.STEP_BINARY_COLUMN, _ffi.STEP_BINARY_DEVSCALAR)
        elif name in _CHAIN_COMPARE:
            if k != len(steps) - 1:
                raise Panic("a compare can only end a fused chain")
            op, kinds, is_pred = _CHAIN_COMPARE[name], (_ffi.STEP_COMPARE_COLUMN, _ffi.STEP_COMPARE_DEVSCALAR), True
        else:
            raise Panic(f"fused_chain_int: unknown step {step!r}")
        arr[k].op = op
        if isinstance(operand, (int, np.integer)):
            operand = DeviceScalar(type(data).from_slice([operand], data.gpu_device))
            keep.append(operand)
        if isinstance(operand, DeviceScalar):
            operand = operand.array
            scalar = True
        else:
            scalar = isinstance(operand, PrimitiveArrayGpu) and operand.len == 1 and data.len != 1
        if type(operand) is not type(data):
            raise Panic(f"fused_chain_int: operand of {name!r} must be a {type(data).__name__}")
        if scalar:
            arr[k].kind, arr[k].operand = kinds[1], operand.data.ptr
        else:
            _check_same_len(data, operand, "fused_chain_int")
            arr[k].kind, arr[k].operand, arr[k].validity = kinds[0], operand.data.ptr, _vptr(operand.null_buffer)
            validities.append(operand.null_buffer)
    dev = data.gpu_device
    _note_foreign_buffers(dev, [data] + [st[1] for st in steps if len(st) > 1 and isinstance(st[1], PrimitiveArrayGpu)])
    nb = _new_validity(dev, data.len, *validities)
    out = (BooleanArrayGPU if is_pred else type(data)).empty(data.len, dev, nb)
    check(lib().agpu_fused_chain_int(dev.handle, data.DTYPE, data.data.ptr, _vptr(data.null_buffer), arr, len(steps),
                                     out.data.ptr, data.len, _vptr(nb)), "fused_chain_int")
    return out


def fused_chain_int(data, steps):
    pipeline = _pipeline_for(data)
    out = fused_chain_int_op(data, steps, pipeline)
    pipeline.finish()
    return out


# ==========================================================================================
# auto-fusion: ArrowComputePipeline(device, fuse=True)
# ==========================================================================================
# The reference's recorded-chain pattern (crates/arrow/examples/simple.rs:45-72) issues one dispatch
# per `*_op`.  On a fusing pipeline eligible ops only RECORD: the returned Float32ArrayGPU carries
# (source column, steps) instead of a buffer.  Applying another eligible op to it extends the chain;
# a compare ends it and launches the single fused kernel at once; `finish()` launches every chain
# that was not absorbed into a longer one; reading `data` / `null_buffer` / values of a recorded
# array launches its own chain on demand.  Anything not eligible falls back to the plain kernel
# (its lazy operands are launched first), so results never differ from fuse=False.
import weakref  # noqa: E402

_FUSE_UNARY = set(_CHAIN_UNARY)
_FUSE_BINARY = {"add", "sub", "mul", "div", "min", "max", "power"}
_FUSE_SCALAR = {"add_scalar": "add", "sub_scalar": "sub", "mul_scalar": "mul", "div_scalar": "div", "rem_scalar": "rem"}
_FUSE_COMPARE = set(_CHAIN_COMPARE)
_MAX_CHAIN_COLS = 3


_FUSE_INT_BINARY = {"add", "sub", "mul", "div", "min", "max", "bitwise_and", "bitwise_or", "bitwise_xor"}


class _LazyChain:
    """mode "f32": steps of agpu_fused_chain on cast<f32>(source); mode "int": steps of
    agpu_fused_chain_int in the source's own integer type"""

    def __init__(self, source, steps, pipeline, mode="f32"):
        self.source, self.steps, self.pipeline, self.mode, self.consumed = source, steps, pipeline, mode, False

    def n_cols(self):
        return sum(1 for st in self.steps if len(st) > 1 and isinstance(st[1], PrimitiveArrayGpu))

    def evaluate(self):
        if self.mode == "int":
            return fused_chain_int_op(self.source, self.steps, self.pipeline)
        if not self.steps:      # a bare int -> f32 cast that nothing was chained onto
            return _cast_to(self.source, Float32ArrayGPU, self.pipeline)
        return fused_chain_op(self.source, self.steps, self.pipeline)


def _lazy_array(source, steps, pipeline, mode="f32"):
    cls = Float32ArrayGPU if mode == "f32" else type(source)
    arr = cls(None, source.gpu_device, source.len, None)
    arr._lazy = _LazyChain(source, steps, pipeline, mode)
    pipeline._lazies.append(weakref.ref(arr))
    return arr


def _extend(self, step, pipeline, mode="f32"):
    """new recorded array = chain of `self` + step (or a fresh chain starting at `self`)"""
    lazy = self._lazy
    adds_col = len(step) > 1 and isinstance(step[1], PrimitiveArrayGpu)
    if (lazy is not None and lazy.pipeline is pipeline and lazy.mode == mode and len(lazy.steps) < _ffi.CHAIN_MAX_STEPS
            and (not adds_col or lazy.n_cols() < _MAX_CHAIN_COLS)):
        lazy.consumed = True
        return lazy.source, lazy.steps + [step]
    return self, [step]      # `self` (concrete, or launched on demand) becomes the source


def _try_fuse_int(base, self, operand, pipeline):
    """integer columns: wrapping arithmetic, min/max, bitwise logic, scalar ops and a closing
    compare of ONE integer type record into an agpu_fused_chain_int chain"""
    if base == "bitwise_not" and operand is None:
        return _lazy_array(*_extend(self, (base,), pipeline, "int"), pipeline, "int")
    if base == "abs" and operand is None and isinstance(self, Int32ArrayGPU):
        return _lazy_array(*_extend(self, (base,), pipeline, "int"), pipeline, "int")
    if base in _INT_CHAIN_SHIFT:
        if not isinstance(operand, UInt32ArrayGPU) or operand.len != self.len:
            return None
        lazy = self._lazy
        if lazy is not None and lazy.mode == "int" and any(st[0] in _INT_CHAIN_SHIFT for st in lazy.steps):
            source, steps = self, [(base, operand)]      # one shift per kernel: start a new chain here
        else:
            source, steps = _extend(self, (base, operand), pipeline, "int")
        return _lazy_array(source, steps, pipeline, "int")
    if type(operand) is not type(self):
        return None
    if base in _FUSE_SCALAR:
        if operand.len != 1:
            return None
        return _lazy_array(*_extend(self, (_FUSE_SCALAR[base], DeviceScalar(operand)), pipeline, "int"), pipeline, "int")
    if operand.len != self.len:
        return None
    if base in _FUSE_INT_BINARY or (base == "power" and isinstance(self, Int32ArrayGPU)):
        return _lazy_array(*_extend(self, (base, operand), pipeline, "int"), pipeline, "int")
    if base in _FUSE_COMPARE:
        source, steps = _extend(self, (base, operand), pipeline, "int")
        return fused_chain_int_op(source, steps, pipeline)    # a predicate ends the chain: launch now
    return None


def _try_fuse(name, self, args):
    """returns the recorded/fused result, or None when the op is not eligible"""
    pipeline = args[-1] if args and isinstance(args[-1], ArrowComputePipeline) else None
    if pipeline is None or not pipeline.fuse or not isinstance(self, PrimitiveArrayGpu):
        return None
    base = name[:-3]
    is_f32 = isinstance(self, Float32ArrayGPU)
    operand = args[0] if len(args) > 1 else None
    if base == "cast" and isinstance(self, _TRIG_INT) and self._lazy is None:
        into = ARRAY_TYPES[operand] if isinstance(operand, ArrowType) else operand
        if into is Float32ArrayGPU:
            return _lazy_array(self, [], pipeline)
        return None
    if base in _FUSE_UNARY and operand is None:
        if is_f32 or (base in ("sin", "cos", "sinh") and isinstance(self, _TRIG_INT)):
            return _lazy_array(*_extend(self, (base,), pipeline), pipeline)
    if isinstance(self, _INT_TYPES):
        return _try_fuse_int(base, self, operand, pipeline)
    if not is_f32 or not isinstance(operand, Float32ArrayGPU):
        return None
    if base in _FUSE_SCALAR:
        if operand.len != 1:
            return None
        return _lazy_array(*_extend(self, (_FUSE_SCALAR[base], DeviceScalar(operand)), pipeline), pipeline)
    if operand.len != self.len:
        return None
    if base in _FUSE_BINARY:
        return _lazy_array(*_extend(self, (base, operand), pipeline), pipeline)
    if base in _FUSE_COMPARE:
        source, steps = _extend(self, (base, operand), pipeline)
        partner = _pending_value_chain(pipeline, source, steps)
        if partner is not None:
            # a recorded value chain over the same source is still waiting (s = a + b; g = a > b):
            # both results come out of ONE kernel that reads the shared columns once
            value, pred = fused_chain_pair_op(source, partner._lazy.steps, steps, pipeline)
            partner._lazy = None
            partner._data, partner._null_buffer = value._data, value._null_buffer
            return pred
        return fused_chain_op(source, steps, pipeline)       # a predicate ends the chain: launch now
    return None


def _pending_value_chain(pipeline, source, pred_steps):
    """the most recently recorded, not yet launched f32 value chain of `pipeline` that starts at the
    same source column as the predicate chain about to be launched and may share its kernel"""
    for ref in reversed(pipeline._lazies):
        arr = ref()
        lazy = arr._lazy if arr is not None else None
        if (lazy is None or lazy.consumed or lazy.mode != "f32" or lazy.pipeline is not pipeline or lazy.source is not source
                or any(c is arr for c in _step_columns(pred_steps))):      # the predicate reads this very result
            continue
        # looking at the operands launches the ones that are still recorded chains; if one of them
        # depended on `arr`, that launched `arr` too and there is nothing left to pair with
        if pair_eligible(source, lazy.steps, pred_steps) and arr._lazy is lazy:
            return arr
    return None


def _fusing(name, fn):
    def wrapper(self, *args, **kwargs):
        if not kwargs:
            out = _try_fuse(name, self, args)
            if out is not None:
                return out
        return fn(self, *args, **kwargs)
    wrapper.__name__ = getattr(fn, "__name__", name)
    wrapper.__doc__ = fn.__doc__
    return wrapper


for _name, _fn in list(vars(PrimitiveArrayGpu).items()):
    if _name.endswith("_op") and callable(_fn) and not isinstance(_fn, (classmethod, staticmethod)):
        setattr(PrimitiveArrayGpu, _name, _fusing(_name, _fn))


# ==========================================================================================
# profiling hook (the reference's `profile` feature, gpu_utils/compute_query.rs): every `*_op`
# recorded on a pipeline created with profile=True is bracketed by a CUDA event pair
# ==========================================================================================
def _note_foreign_buffers(dev, operands) -> None:
    """ops run on `dev` (= self.gpu_device).  A column whose buffer was allocated through another
    handle of the same GPU (uploaded on a copy stream, say) is recorded as used by `dev`, so that
    dropping it cannot hand the block out again while this op still reads it."""
    for x in operands:
        if isinstance(x, PrimitiveArrayGpu):
            buf, nb = x._data, x._null_buffer
        elif isinstance(x, BooleanArrayGPU):
            buf, nb = x.data, x.null_buffer
        else:
            continue
        if buf is not None and buf.device is not dev and buf._kind == "pool" and buf._owned:
            dev.record_use(buf)
        if nb is not None and nb.bit_buffer.device is not dev and nb.bit_buffer._kind == "pool" and nb.bit_buffer._owned:
            dev.record_use(nb.bit_buffer)


def _profiled(name, fn):
    in_place = name == "put_op"

    def wrapper(self, *args, **kwargs):
        pipeline = next((a for a in reversed(args) if isinstance(a, ArrowComputePipeline)), None)
        _note_foreign_buffers(self.gpu_device, (self,) + args)
        if in_place and pipeline is not None and pipeline._lazies:
            # put_op mutates `dst` right away; chains recorded earlier on this pipeline that read it
            # have to run first, as they do in the reference's encoder order
            pipeline.flush_recorded()
        if pipeline is not None and pipeline.profile:
            start = pipeline.device.record_event()
            out = fn(self, *args, **kwargs)
            pipeline.queries.append((name, start, pipeline.device.record_event()))
            return out
        return fn(self, *args, **kwargs)
    wrapper.__name__ = getattr(fn, "__name__", name)
    wrapper.__doc__ = fn.__doc__
    return wrapper


for _cls in (PrimitiveArrayGpu, BooleanArrayGPU):
    for _name, _fn in list(vars(_cls).items()):
        if _name.endswith("_op") and callable(_fn) and not isinstance(_fn, (classmethod, staticmethod)):
            setattr(_cls, _name, _profiled(_name, _fn))
