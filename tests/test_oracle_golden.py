"""CPU: pin the oracle against EVERY golden vector of the reference's own unit tests
(tests/golden/reference_vectors.json, extracted from the reference's test macros) and the
hand-written reference tests.  This is what makes `oracle/` a trusted checker."""
import numpy as np
import pytest

import oracle as O
from helpers import (DTYPE_OF, TYPE_OF_ARROW, OArr, assert_values, decode, load_cases, oracle_binary, oracle_cast,
                     oracle_merge, oracle_put, oracle_scalar, oracle_take, oracle_unary)


def ids(cases):
    return [c["name"] for c in cases]


UNARY = load_cases("test_unary_op", "test_unary_op_float")


@pytest.mark.parametrize("case", UNARY, ids=ids(UNARY))
def test_unary(case):
    a = OArr.from_slice(DTYPE_OF[case["input_type"]], decode(case["input"]))
    out = oracle_unary(case["op"], a)
    assert out.dtype == DTYPE_OF[case["output_type"]]
    assert_values(out.raw_values(), decode(case["expected"]), float_tol=case["macro"].endswith("float"),
                  what=case["name"])


SCALAR = load_cases("test_scalar_op", "test_float_scalar_op")


@pytest.mark.parametrize("case", SCALAR, ids=ids(SCALAR))
def test_scalar(case):
    a = OArr.from_slice(DTYPE_OF[case["input_type"]], decode(case["input"]))
    s = OArr.from_slice(DTYPE_OF[case["scalar_type"]], [decode(case["scalar"])])
    out = oracle_scalar(case["op"], a, s)
    assert_values(out.raw_values(), decode(case["expected"]), float_tol="float" in case["macro"], what=case["name"])


ARRAY = load_cases("test_array_op", "test_float_array_op")


@pytest.mark.parametrize("case", ARRAY, ids=ids(ARRAY))
def test_array(case):
    a = OArr.from_optional(DTYPE_OF[case["lhs_type"]], decode(case["lhs"]))
    b = OArr.from_optional(DTYPE_OF[case["rhs_type"]], decode(case["rhs"]))
    out = oracle_binary(case["op"], a, b)
    assert_values(out.values(), decode(case["expected"]), float_tol="float" in case["macro"], what=case["name"])


CAST = load_cases("test_cast_op", "test_bitcast_op")


@pytest.mark.parametrize("case", CAST, ids=ids(CAST))
def test_cast(case):
    a = OArr.from_slice(DTYPE_OF[case["input_type"]], decode(case["input"]))
    out = oracle_cast(a, DTYPE_OF[TYPE_OF_ARROW[case["cast_type"]]])
    assert out.dtype == DTYPE_OF[case["output_type"]]
    if case["macro"] == "test_bitcast_op":
        exp_bits = [np.array([decode(x)], dtype=np.float32).view(np.uint32)[0] if not isinstance(x, dict)
                    else x["f32_bits"] for x in case["expected"]]
        assert list(out.raw_values().view(np.uint32)) == exp_bits
    else:
        assert_values(out.raw_values(), decode(case["expected"]), float_tol=False, what=case["name"])


BROADCAST = load_cases("test_broadcast")


@pytest.mark.parametrize("case", BROADCAST, ids=ids(BROADCAST))
def test_broadcast(case):
    dt = DTYPE_OF[case["output_type"]]
    v = decode(case["value"])
    if dt == O.BOOL:
        got = np.full(case["length"], bool(v))  # host-built in the reference (boolean_gpu.rs:119-135)
    else:
        got = O.broadcast(dt, v, case["length"])
    assert list(got) == [v] * case["length"]


SUM = load_cases("test_sum")


@pytest.mark.parametrize("case", SUM, ids=ids(SUM))
def test_sum(case):
    dt = DTYPE_OF[case["input_type"]]
    arr = O.broadcast(dt, decode(case["base"]), case["size"])
    got = O.sum(dt, arr)
    exp = decode(case["expected"])
    if dt == O.U32:
        exp %= 1 << 32
    assert got == np.asarray(exp).astype(O.NP[dt])


MERGE = load_cases("test_merge_op")


@pytest.mark.parametrize("case", MERGE, ids=ids(MERGE))
def test_merge(case):
    a = OArr.from_optional(DTYPE_OF[case["lhs_type"]], decode(case["lhs"]))
    b = OArr.from_optional(DTYPE_OF[case["rhs_type"]], decode(case["rhs"]))
    m = OArr.from_optional(O.BOOL, decode(case["mask"]))
    out = oracle_merge(a, b, m)
    assert_values(out.values(), decode(case["expected"]), float_tol=False, what=case["name"])


TAKE = load_cases("test_take_op")


@pytest.mark.parametrize("case", TAKE, ids=ids(TAKE))
def test_take(case):
    dt = DTYPE_OF[case["lhs_type"]]
    a = OArr.from_optional(dt, decode(case["lhs"])) if case["lhs_optional"] else OArr.from_slice(dt, decode(case["lhs"]))
    idx = OArr.from_slice(O.U32, decode(case["rhs"]))
    out = oracle_take(a, idx)
    got = out.values() if case["lhs_optional"] else list(out.raw_values())
    assert_values(got, decode(case["expected"]), float_tol=False, what=case["name"])


PUT = load_cases("test_put_op")


@pytest.mark.parametrize("case", PUT, ids=ids(PUT))
def test_put(case):
    dt = DTYPE_OF[case["array_type"]]
    src = OArr.from_slice(dt, decode(case["src"]))
    dst = OArr.from_slice(dt, decode(case["dst"]))
    out = oracle_put(src, OArr.from_slice(O.U32, case["src_indexes"]), dst, OArr.from_slice(O.U32, case["dst_indexes"]))
    assert_values(list(out.raw_values()), decode(case["expected"]), float_tol=False, what=case["name"])


def test_every_reference_macro_case_is_covered():
    covered = len(UNARY) + len(SCALAR) + len(ARRAY) + len(CAST) + len(BROADCAST) + len(SUM) + len(MERGE) + len(TAKE) + len(PUT)
    assert covered == len(load_cases()) == 202


# ---- hand-written reference tests (not macro generated) ---------------------------------
def test_f32_array_from_optional_vec_and_null_and():
    """crates/array/src/array/f32_gpu.rs:91-123"""
    a = OArr.from_optional(O.F32, [0.0, 1.0, None, None, 4.0])
    assert list(a.raw_values()) == [0.0, 1.0, 0.0, 0.0, 4.0]
    assert a.valid.view(np.uint8)[0] == 0b00010011
    b = OArr.from_optional(O.F32, [1.0, 2.0, None, 4.0, None])
    assert list(b.raw_values()) == [1.0, 2.0, 0.0, 4.0, 0.0]
    assert b.valid.view(np.uint8)[0] == 0b00001011
    assert O.validity_and(b.valid, a.valid, 5).view(np.uint8)[0] == 0b00000011


def test_boolean_values():
    """crates/array/src/array/boolean_gpu.rs:208-228"""
    values = [True, True, False, None] * 101
    a = OArr.from_optional(O.BOOL, values)
    assert list(a.raw_values()) == [True, True, False, False] * 101
    assert a.values() == values


def test_large_f32_scalar_add():
    """crates/arithmetic/src/f32.rs:189-207 (10 Mi rows)"""
    n = 1024 * 1024 * 10
    x = np.arange(n, dtype=np.float32)
    out = O.scalar(O.ADD, O.F32, x, 100.0)
    assert np.array_equal(out, x + np.float32(100.0))


def test_any():
    """crates/logical/src/boolean.rs:259-283"""
    assert O.any(O.pack_bits([True, True, False, True, False]), 5) is True
    assert O.any(O.pack_bits([True] * 16384), 16384) is True
    data = [False] * 16384
    assert O.any(O.pack_bits(data), 16384) is False
    data += [True] * 16384
    assert O.any(O.pack_bits(data), 32768) is True


def test_all():
    """crates/logical/src/boolean.rs:285-319"""
    assert O.all(O.pack_bits([True, True, False, True, False]), 5) is False
    assert O.all(O.pack_bits([True] * 100), 100) is True
    assert O.all(O.pack_bits([False] * 100), 100) is False
    n = 1024 * 1024 * 2
    assert O.all(O.pack_bits(np.zeros(n, bool)), n) is False
    data = np.ones(n + 1, bool)
    assert O.all(O.pack_bits(data[:n]), n) is True
    data[n] = False
    assert O.all(O.pack_bits(data), n + 1) is False


def test_ref_quirks_switch():
    """the reference's u32 min/max shader compares as signed"""
    a, b = np.array([1, 0xFFFFFFFF], np.uint32), np.array([2, 1], np.uint32)
    assert list(O.binary(O.MAX, O.U32, a, b)) == [2, 0xFFFFFFFF]
    O.set_ref_quirks(True)
    try:
        assert list(O.binary(O.MAX, O.U32, a, b)) == [2, 1]
    finally:
        O.set_ref_quirks(False)
