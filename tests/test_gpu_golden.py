"""GPU: every golden vector of the reference's unit tests through the product path
(arrow_gpu_b200 -> ctypes -> libagpu.so -> sm_100a kernels), with the typed method AND the
`_dyn` function exactly as the reference's test macros do (crates/test_macros/src/lib.rs), and
cross-checked bit-for-bit against the oracle on the same inputs."""
import numpy as np
import pytest

import arrow_gpu_b200 as ag
from arrow_gpu_b200 import kernels as K
import oracle as O
from helpers import (DTYPE_OF, TYPE_OF_ARROW, OArr, assert_values, decode, load_cases, oracle_binary, oracle_cast,
                     oracle_merge, oracle_scalar, oracle_take, oracle_unary, same_f32_bits)

pytestmark = pytest.mark.gpu


def ids(cases):
    return [c["name"] for c in cases]


def cls(name):
    return ag.ARRAY_BY_NAME[name]


def same_bits(gpu_arr, oarr: OArr, what):
    """GPU result vs oracle result: identical data buffers over the first n rows and identical
    validity bits (integer/bitmap/cast/index work must be bit-exact)"""
    assert gpu_arr.len == oarr.n, what
    g = gpu_arr.raw_values()
    o = oarr.raw_values()
    if g.dtype == np.float32:
        assert same_f32_bits(g, o), what
    else:
        assert np.array_equal(g, o), what
    if oarr.valid is None:
        assert gpu_arr.null_buffer is None, what
    else:
        assert np.array_equal(gpu_arr.null_buffer.flags(), O.unpack_bits(oarr.valid, oarr.n)), what


EXACT_F32_OPS = {"neg", "abs", "sqrt", "add", "sub", "mul", "div", "min", "max"}

UNARY = load_cases("test_unary_op", "test_unary_op_float")


@pytest.mark.parametrize("case", UNARY, ids=ids(UNARY))
def test_unary(case, device):
    arr = cls(case["input_type"]).from_slice(decode(case["input"]), device)
    expected = decode(case["expected"])
    tol = case["macro"].endswith("float")
    out = getattr(arr, case["op"])()
    assert type(out) is cls(case["output_type"])
    assert_values(out.raw_values(), expected, float_tol=tol, what=case["name"])
    if case["op_dyn"]:
        out_dyn = getattr(K, case["op_dyn"])(arr)
        assert_values(out_dyn.raw_values(), expected, float_tol=tol, what=case["name"] + " (dyn)")
    o = oracle_unary(case["op"], OArr.from_slice(DTYPE_OF[case["input_type"]], decode(case["input"])))
    if o.dtype != O.F32 or case["op"] in EXACT_F32_OPS:
        same_bits(out, o, case["name"])


SCALAR = load_cases("test_scalar_op", "test_float_scalar_op")


@pytest.mark.parametrize("case", SCALAR, ids=ids(SCALAR))
def test_scalar(case, device):
    arr = cls(case["input_type"]).from_slice(decode(case["input"]), device)
    sc = cls(case["scalar_type"]).from_slice([decode(case["scalar"])], device)
    expected = decode(case["expected"])
    tol = "float" in case["macro"]
    out = getattr(arr, case["op"])(sc)
    assert_values(out.raw_values(), expected, float_tol=tol, what=case["name"])
    out_dyn = getattr(K, case["op_dyn"])(arr, sc)
    assert type(out_dyn) is cls(case["output_type"])
    assert_values(out_dyn.raw_values(), expected, float_tol=tol, what=case["name"] + " (dyn)")
    o = oracle_scalar(case["op"], OArr.from_slice(DTYPE_OF[case["input_type"]], decode(case["input"])),
                      OArr.from_slice(DTYPE_OF[case["scalar_type"]], [decode(case["scalar"])]))
    same_bits(out, o, case["name"])  # f32 + - * / % are single correctly rounded ops: bit-exact


ARRAY = load_cases("test_array_op", "test_float_array_op")


@pytest.mark.parametrize("case", ARRAY, ids=ids(ARRAY))
def test_array(case, device):
    a = cls(case["lhs_type"]).from_optional_slice(decode(case["lhs"]), device)
    b = cls(case["rhs_type"]).from_optional_slice(decode(case["rhs"]), device)
    expected = decode(case["expected"])
    tol = "float" in case["macro"]
    out = getattr(a, case["op"])(b)
    assert_values(out.values(), expected, float_tol=tol, what=case["name"])
    if case["op_dyn"]:
        out_dyn = getattr(K, case["op_dyn"])(a, b)
        assert type(out_dyn) is cls(case["output_type"])
        assert_values(out_dyn.values(), expected, float_tol=tol, what=case["name"] + " (dyn)")
    o = oracle_binary(case["op"], OArr.from_optional(DTYPE_OF[case["lhs_type"]], decode(case["lhs"])),
                      OArr.from_optional(DTYPE_OF[case["rhs_type"]], decode(case["rhs"])))
    if case["op"] != "power" or o.dtype != O.F32:
        same_bits(out, o, case["name"])


CAST = load_cases("test_cast_op", "test_bitcast_op")


@pytest.mark.parametrize("case", CAST, ids=ids(CAST))
def test_cast(case, device):
    arr = cls(case["input_type"]).from_slice(decode(case["input"]), device)
    into = ag.ArrowType[case["cast_type"]]
    bitcast = case["macro"] == "test_bitcast_op"
    out = arr.bitcast(cls(case["output_type"])) if bitcast else arr.cast(cls(case["output_type"]))
    out_dyn = (K.bitcast_dyn if bitcast else K.cast_dyn)(arr, into)
    assert type(out) is type(out_dyn) is cls(case["output_type"])
    o = oracle_cast(OArr.from_slice(DTYPE_OF[case["input_type"]], decode(case["input"])),
                    DTYPE_OF[TYPE_OF_ARROW[case["cast_type"]]])
    same_bits(out, o, case["name"])
    same_bits(out_dyn, o, case["name"] + " (dyn)")
    if not bitcast:
        assert_values(out.raw_values(), decode(case["expected"]), float_tol=False, what=case["name"])


BROADCAST = load_cases("test_broadcast")


@pytest.mark.parametrize("case", BROADCAST, ids=ids(BROADCAST))
def test_broadcast(case, device):
    v = decode(case["value"])
    out = cls(case["output_type"]).broadcast(v, case["length"], device)
    assert list(out.raw_values()) == [v] * case["length"]


SUM = load_cases("test_sum")


@pytest.mark.parametrize("case", SUM, ids=ids(SUM))
def test_sum(case, device):
    c = cls(case["input_type"])
    arr = c.broadcast(decode(case["base"]), case["size"], device)
    got = arr.sum().raw_values()
    exp = decode(case["expected"])
    if c is ag.UInt32ArrayGPU:
        exp %= 1 << 32
    assert list(got) == [np.asarray(exp).astype(c.NP)]


MERGE = load_cases("test_merge_op")


@pytest.mark.parametrize("case", MERGE, ids=ids(MERGE))
def test_merge(case, device):
    a = cls(case["lhs_type"]).from_optional_slice(decode(case["lhs"]), device)
    b = cls(case["rhs_type"]).from_optional_slice(decode(case["rhs"]), device)
    m = ag.BooleanArrayGPU.from_optional_slice(decode(case["mask"]), device)
    expected = decode(case["expected"])
    out = getattr(a, case["op"])(b, m)
    assert_values(out.values(), expected, float_tol=False, what=case["name"])
    if case["op_dyn"]:
        assert_values(getattr(K, case["op_dyn"])(a, b, m).values(), expected, float_tol=False, what=case["name"])
    o = oracle_merge(OArr.from_optional(DTYPE_OF[case["lhs_type"]], decode(case["lhs"])),
                     OArr.from_optional(DTYPE_OF[case["rhs_type"]], decode(case["rhs"])),
                     OArr.from_optional(O.BOOL, decode(case["mask"])))
    same_bits(out, o, case["name"])


TAKE = load_cases("test_take_op")


@pytest.mark.parametrize("case", TAKE, ids=ids(TAKE))
def test_take(case, device):
    c = cls(case["lhs_type"])
    a = (c.from_optional_slice if case["lhs_optional"] else c.from_slice)(decode(case["lhs"]), device)
    idx = ag.UInt32ArrayGPU.from_slice(decode(case["rhs"]), device)
    expected = decode(case["expected"])
    out = a.take(idx)
    got = out.values() if case["lhs_optional"] else list(out.raw_values())
    assert_values(got, expected, float_tol=False, what=case["name"])
    if case["op_dyn"]:
        assert_values(getattr(K, case["op_dyn"])(a, idx).values(), expected, float_tol=False, what=case["name"])
    dt = DTYPE_OF[case["lhs_type"]]
    oa = OArr.from_optional(dt, decode(case["lhs"])) if case["lhs_optional"] else OArr.from_slice(dt, decode(case["lhs"]))
    same_bits(out, oracle_take(oa, OArr.from_slice(O.U32, decode(case["rhs"]))), case["name"])


PUT = load_cases("test_put_op")


@pytest.mark.parametrize("case", PUT, ids=ids(PUT))
def test_put(case, device):
    c = cls(case["array_type"])
    src = c.from_slice(decode(case["src"]), device)
    si = ag.UInt32ArrayGPU.from_slice(case["src_indexes"], device)
    di = ag.UInt32ArrayGPU.from_slice(case["dst_indexes"], device)
    expected = decode(case["expected"])
    dst = c.from_slice(decode(case["dst"]), device)
    src.put(si, dst, di)
    assert_values(list(dst.raw_values()), expected, float_tol=False, what=case["name"])
    if case["op_dyn"]:
        dst2 = c.from_slice(decode(case["dst"]), device)
        getattr(K, case["op_dyn"])(src, si, dst2, di)
        assert_values(list(dst2.raw_values()), expected, float_tol=False, what=case["name"] + " (dyn)")


# ---- hand-written reference tests ----------------------------------------------------------
def test_f32_array_from_optional_vec_and_null_and(device):
    """crates/array/src/array/f32_gpu.rs:91-123"""
    a = ag.Float32ArrayGPU.from_optional_slice([0.0, 1.0, None, None, 4.0], device)
    assert list(a.raw_values()) == [0.0, 1.0, 0.0, 0.0, 4.0]
    assert list(a.null_buffer.raw_values()) == [0b00010011]
    b = ag.Float32ArrayGPU.from_optional_slice([1.0, 2.0, None, 4.0, None], device)
    assert list(b.null_buffer.raw_values()) == [0b00001011]
    merged = ag.NullBitBufferGpu.merge_null_bit_buffer(b.null_buffer, a.null_buffer)
    assert list(merged.raw_values()) == [0b00000011]


def test_boolean_values(device):
    """crates/array/src/array/boolean_gpu.rs:208-228"""
    values = [True, True, False, None] * 101
    arr = ag.BooleanArrayGPU.from_optional_slice(values, device)
    assert list(arr.raw_values()) == [True, True, False, False] * 101
    assert arr.values() == values


def test_large_f32_array(device):
    """crates/arithmetic/src/f32.rs:189-207 (10 Mi rows, scalar add)"""
    n = 1024 * 1024 * 10
    x = np.arange(n, dtype=np.float32)
    out = ag.Float32ArrayGPU.from_slice(x, device).add_scalar(ag.Float32ArrayGPU.from_slice([100.0], device))
    assert np.array_equal(out.raw_values(), x + np.float32(100.0))


def test_any(device):
    """crates/logical/src/boolean.rs:259-283"""
    B = ag.BooleanArrayGPU
    assert B.from_slice([True, True, False, True, False], device).any() is True
    assert B.from_slice([True] * 16384, device).any() is True
    data = [False] * 16384
    assert B.from_slice(data, device).any() is False
    assert B.from_slice(data + [True] * 16384, device).any() is True


def test_all(device):
    """crates/logical/src/boolean.rs:285-319 (ignored in the reference's CI; passes here)"""
    B = ag.BooleanArrayGPU
    assert B.from_slice([True, True, False, True, False], device).all() is False
    assert B.from_slice([True] * 100, device).all() is True
    assert B.from_slice([False] * 100, device).all() is False
    n = 1024 * 1024 * 2
    assert B.from_slice(np.zeros(n, bool), device).all() is False
    data = np.ones(n + 1, bool)
    assert B.from_slice(data[:n], device).all() is True
    data[n] = False
    assert B.from_slice(data, device).all() is False


def test_example_simple(device):
    """crates/arrow/examples/simple.rs: eager op, dyn op and a recorded two-op pipeline"""
    vals = np.arange(100, dtype=np.float32)
    arr = ag.Float32ArrayGPU.from_slice(vals, device)
    sc = ag.Float32ArrayGPU.from_slice([20.0], device)
    assert arr.add_scalar(sc).values() == list(vals + 20.0)
    assert K.add_scalar_dyn(arr, sc).values() == list(vals + 20.0)
    pipeline = ag.ArrowComputePipeline(device, "example")
    r1 = K.add_scalar_op_dyn(arr, sc, pipeline)
    r2 = K.mul_scalar_op_dyn(r1, sc, pipeline)
    pipeline.finish()
    assert r2.values() == list((vals + 20.0) * 20.0)


def test_unsupported_pairs_panic(device):
    """the reference panic!s on unsupported dtype pairs (arithmetic_kernels.rs:92-97 etc.)"""
    f = ag.Float32ArrayGPU.from_slice([1.0], device)
    i = ag.Int32ArrayGPU.from_slice([1, 2], device)
    u8 = ag.UInt8ArrayGPU.from_slice([1, 2], device)
    for fn, args in ((K.add_array_dyn, (i, f)), (K.bitwise_and_dyn, (f, f)), (K.gt_dyn, (i, u8)),
                     (K.sqrt_dyn, (i,)), (K.acos_dyn, (u8,)), (K.neg_dyn, (i,))):
        with pytest.raises(ag.Panic):
            fn(*args)
    with pytest.raises(ag.Panic):
        K.cast_dyn(i, ag.ArrowType.Float32Type)  # not in the reference's cast matrix


def test_pipeline_profile_feature(device):
    """the reference's `profile` feature (gpu_utils/compute_query.rs): per-op timestamps"""
    arr = ag.Float32ArrayGPU.from_slice(np.arange(1 << 20, dtype=np.float32), device)
    sc = ag.Float32ArrayGPU.from_slice([2.0], device)
    pipeline = ag.ArrowComputePipeline(device, "profiled", profile=True)
    r = K.mul_scalar_op_dyn(K.add_scalar_op_dyn(arr, sc, pipeline), sc, pipeline)
    K.gt_op_dyn(r, arr, pipeline)
    pipeline.finish()
    res = pipeline.wait_for_results()
    assert [n for n, _ in res] == ["add_scalar_op", "mul_scalar_op", "gt_op"]
    assert all(0.0 < ms < 50.0 for _, ms in res)


def test_arrow_interop_roundtrip(device):
    pa = pytest.importorskip("pyarrow")
    import pyarrow.compute as pc
    cases = [pa.array([1, None, -3, 4, None], pa.int8()), pa.array([1.5, 2.5, None], pa.float32()),
             pa.array([True, None, False, True] * 11, pa.bool_()), pa.array(list(range(100)), pa.uint16()),
             pa.array([None, 7], pa.date32()), pa.array([], pa.int32()), pa.array(list(range(50)), pa.int32())[3:40]]
    for a in cases:
        g = ag.from_arrow(a, device)
        assert g.len == len(a)
        assert ag.to_arrow(g).equals(a if a.offset == 0 else pa.concat_arrays([a])), a.type
    # same results as pyarrow.compute where the two agree (SURVEY.md §8c cross-check list)
    x = pa.array([100, 127, -128, None, 5], pa.int8())
    y = pa.array([100, 1, -1, 3, None], pa.int8())
    gx, gy = ag.from_arrow(x, device), ag.from_arrow(y, device)
    assert ag.to_arrow(gx.add(gy)).equals(pc.add(x, y))            # wrap-around
    assert ag.to_arrow(gx.gt(gy)).equals(pc.greater(x, y))
    assert ag.to_arrow(gx.max(gy)).to_pylist() == pc.max_element_wise(x, y, skip_nulls=False).to_pylist()
