"""GPU: randomized bit-parity of every (op, dtype) against the oracle on the same seeded inputs,
at ragged sizes (0, 1, word/tile boundaries +-1, > one tile), with and without nulls and through
unaligned views; ULP bounds for f32 transcendentals; size-independent properties at large N."""
import numpy as np
import pytest

import arrow_gpu_b200 as ag
from arrow_gpu_b200 import kernels as K
from arrow_gpu_b200 import _ffi
import oracle as O
from helpers import OArr, oracle_binary, oracle_cast, oracle_scalar, oracle_unary, same_f32_bits, ulp_diff

pytestmark = pytest.mark.gpu

INT_CLS = {O.I8: ag.Int8ArrayGPU, O.U8: ag.UInt8ArrayGPU, O.I16: ag.Int16ArrayGPU, O.U16: ag.UInt16ArrayGPU,
           O.I32: ag.Int32ArrayGPU, O.U32: ag.UInt32ArrayGPU}
ALL_CLS = dict(INT_CLS)
ALL_CLS[O.F32] = ag.Float32ArrayGPU
ALL_CLS[O.DATE32] = ag.Date32ArrayGPU
NAMES = {O.I8: "i8", O.U8: "u8", O.I16: "i16", O.U16: "u16", O.I32: "i32", O.U32: "u32", O.F32: "f32", O.DATE32: "date32"}
SIZES = [0, 1, 3, 31, 32, 33, 127, 1000, 4096, 16384 + 5, 70001]


def stable_seed(*parts) -> int:
    """same seed in every process (the builtin hash of a str is salted per interpreter)"""
    import zlib
    return zlib.crc32(repr(parts).encode())


def rand_vals(rng, dtype, n, special=True):
    if dtype == O.F32:
        x = rng.uniform(-1000, 1000, n).astype(np.float32)
        if special and n >= 8:
            x[:8] = [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-40, -1e-40, 3.4e38]
        return x
    info = np.iinfo(O.NP[dtype])
    x = rng.integers(info.min, int(info.max) + 1, n, dtype=np.int64).astype(O.NP[dtype])
    if special and n >= 6:
        x[:6] = np.array([0, 1, -1 if info.min < 0 else info.max, info.min, info.max, 2]).astype(O.NP[dtype])
    return x


def make(rng, dtype, n, nulls, device, vals=None):
    vals = rand_vals(rng, dtype, n) if vals is None else vals
    valid = rng.random(n) < 0.85 if nulls else None
    g = ALL_CLS[dtype].from_numpy(vals, valid, device)
    o = OArr(dtype, vals.copy(), n, O.pack_bits(valid) if nulls else None)
    return g, o


def assert_same(g, o: OArr, what):
    assert g.len == o.n, what
    gv, ov = g.raw_values(), o.raw_values()
    if gv.dtype == np.float32:
        assert same_f32_bits(gv, ov), what
    else:
        assert np.array_equal(gv, ov), what
    if o.valid is None:
        assert g.null_buffer is None, what
    else:
        # whole words must match: padding bits are zero on both sides
        raw = g.gpu_device.retrive_data(g.null_buffer.bit_buffer, O.words(o.n) * 4).view(np.uint32)
        assert np.array_equal(raw, o.valid), what + " (validity)"


BIN_INT = ["add", "sub", "mul", "div", "min", "max", "bitwise_and", "bitwise_or", "bitwise_xor"]
OPS_NEW_SURFACE = {"div"}  # array/array int div is not a reference method name: use the C ABI id


@pytest.mark.parametrize("dtype", list(INT_CLS), ids=lambda d: NAMES[d])
@pytest.mark.parametrize("op", ["add", "sub", "mul", "min", "max", "bitwise_and", "bitwise_or", "bitwise_xor"])
def test_int_binary(op, dtype, device):
    rng = np.random.default_rng(stable_seed(op, dtype))
    for n in SIZES:
        for nulls in (False, True):
            a, oa = make(rng, dtype, n, nulls, device)
            b, ob = make(rng, dtype, n, nulls and n % 2 == 0, device)
            assert_same(getattr(a, op)(b), oracle_binary(op, oa, ob), f"{op} {NAMES[dtype]} n={n} nulls={nulls}")


@pytest.mark.parametrize("dtype", list(INT_CLS) + [O.F32], ids=lambda d: NAMES[d])
@pytest.mark.parametrize("op", ["add_scalar", "sub_scalar", "mul_scalar", "div_scalar", "rem_scalar"])
def test_scalar(op, dtype, device):
    rng = np.random.default_rng(stable_seed(op, dtype))
    scalars = [3, 0, 1] if dtype != O.F32 else [3.5, 0.0, -0.25]
    if np.dtype(O.NP[dtype]).kind == "i":
        scalars.append(-1)  # MIN / -1 and MIN % -1
    for n in (0, 5, 1000, 70001):
        for s in scalars:
            a, oa = make(rng, dtype, n, n % 2 == 1, device)
            sc = ALL_CLS[dtype].from_slice([s], device)
            want = oracle_scalar(op, oa, OArr.from_slice(dtype, [s]))
            assert_same(getattr(a, op)(sc), want, f"{op} {NAMES[dtype]} n={n} s={s}")


@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "min", "max"])
def test_f32_binary(op, device):
    rng = np.random.default_rng(7)
    for n in SIZES:
        a, oa = make(rng, O.F32, n, True, device)
        b, ob = make(rng, O.F32, n, False, device)
        if n >= 16:  # 0/0, x/0, -0 vs +0 for min/max, NaN pairs
            ob.data[:16] = [0.0, 0.0, np.nan, -0.0, 0.0, 5.0, np.inf, -np.inf, 0.0, -0.0, np.nan, 1.0, 2.0, 0.0, 0.0, 0.0]
            oa.data[8:16] = [-0.0, 0.0, np.nan, np.nan, 0.0, 1.0, -1.0, np.inf]
            a = ag.Float32ArrayGPU.from_numpy(oa.data, O.unpack_bits(oa.valid, n), device)
            b = ag.Float32ArrayGPU.from_numpy(ob.data, None, device)
        assert_same(getattr(a, op)(b), oracle_binary(op, oa, ob), f"{op} f32 n={n}")


@pytest.mark.parametrize("dtype", list(ALL_CLS), ids=lambda d: NAMES[d])
@pytest.mark.parametrize("op", ["gt", "gteq", "lt", "lteq", "eq"])
def test_compare(op, dtype, device):
    rng = np.random.default_rng(stable_seed(op, dtype))
    for n in SIZES:
        a, oa = make(rng, dtype, n, True, device)
        b, ob = make(rng, dtype, n, n % 2 == 0, device)
        if n > 20:  # force equal pairs
            ob.data[10:20] = oa.data[10:20]
            b = ALL_CLS[dtype].from_numpy(ob.data, None if ob.valid is None else O.unpack_bits(ob.valid, n), device)
        got = getattr(a, op)(b)
        want = oracle_binary(op, oa, ob)
        raw = device.retrive_data(got.data, O.words(n) * 4).view(np.uint32)
        assert np.array_equal(raw, want.data), f"{op} {NAMES[dtype]} n={n}"   # whole words: padding bits zero
        assert np.array_equal(got.null_buffer.flags(), O.unpack_bits(want.valid, n))


@pytest.mark.parametrize("dtype", list(INT_CLS), ids=lambda d: NAMES[d])
@pytest.mark.parametrize("op", ["bitwise_shl", "bitwise_shr"])
def test_shift(op, dtype, device):
    rng = np.random.default_rng(stable_seed(op, dtype))
    width = O.NP[dtype].itemsize * 8
    for n in SIZES:
        a, oa = make(rng, dtype, n, True, device)
        counts = rng.integers(0, width, n).astype(np.uint32)
        if n > 40:  # counts >= width and >= 32: WGSL takes the count mod 32 on the widened lane (Q17)
            counts[20:30] = rng.integers(width, 64, 10)
            counts[30:34] = [width, 31, 32, 0xFFFFFFFF]   # (the oracle evaluates the i16 shr helper literally there)
        c = ag.UInt32ArrayGPU.from_numpy(counts, None, device)
        oc = OArr(O.U32, counts, n)
        assert_same(getattr(a, op)(c), oracle_binary(op, oa, oc), f"{op} {NAMES[dtype]} n={n}")


@pytest.mark.parametrize("dtype", list(INT_CLS), ids=lambda d: NAMES[d])
def test_not(dtype, device):
    rng = np.random.default_rng(dtype)
    for n in SIZES:
        a, oa = make(rng, dtype, n, n % 2 == 1, device)
        assert_same(a.bitwise_not(), oracle_unary("bitwise_not", oa), f"not {NAMES[dtype]} n={n}")


CASTS = [(O.I8, O.U8), (O.I8, O.U16), (O.I8, O.U32), (O.I8, O.I16), (O.I8, O.I32), (O.I8, O.F32),
         (O.I16, O.I32), (O.I16, O.U16), (O.I16, O.U32), (O.I16, O.F32),
         (O.U8, O.U16), (O.U8, O.U32), (O.U8, O.I8), (O.U8, O.I16), (O.U8, O.I32), (O.U8, O.F32),
         (O.U16, O.U32), (O.U16, O.I16), (O.U16, O.I32), (O.U16, O.F32), (O.F32, O.U8)]


@pytest.mark.parametrize("pair", CASTS, ids=lambda p: f"{NAMES[p[0]]}->{NAMES[p[1]]}")
def test_cast(pair, device):
    src, dst = pair
    rng = np.random.default_rng(src * 16 + dst)
    for n in SIZES:
        vals = None
        if src == O.F32:
            vals = rng.uniform(-10, 70000, n).astype(np.float32)
            if n >= 8:
                vals[:8] = [0.0, -0.0, -1.0, 255.0, 256.0, np.nan, np.inf, 5e9]
        a, oa = make(rng, src, n, n % 2 == 1, device, vals)
        assert_same(a.cast(ALL_CLS[dst]), oracle_cast(oa, dst), f"cast n={n}")


def test_cast_bool_to_f32_and_bitcast(device):
    rng = np.random.default_rng(3)
    for n in SIZES:
        flags = rng.random(n) < 0.5
        valid = rng.random(n) < 0.9
        b = ag.BooleanArrayGPU.from_numpy(flags, valid, device)
        want = oracle_cast(OArr(O.BOOL, O.pack_bits(flags), n, O.pack_bits(valid)), O.F32)
        assert_same(b.cast(ag.Float32ArrayGPU), want, f"bool->f32 n={n}")
        u = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
        g = ag.UInt32ArrayGPU.from_numpy(u, None, device).bitcast(ag.Float32ArrayGPU)
        assert np.array_equal(g.raw_values().view(np.uint32), u)


def test_bool_logical(device):
    rng = np.random.default_rng(5)
    for n in SIZES:
        fa, fb = rng.random(n) < 0.5, rng.random(n) < 0.5
        va, vb = rng.random(n) < 0.9, rng.random(n) < 0.9
        a = ag.BooleanArrayGPU.from_numpy(fa, va, device)
        b = ag.BooleanArrayGPU.from_numpy(fb, vb, device)
        oa = OArr(O.BOOL, O.pack_bits(fa), n, O.pack_bits(va))
        ob = OArr(O.BOOL, O.pack_bits(fb), n, O.pack_bits(vb))
        for op in ("bitwise_and", "bitwise_or", "bitwise_xor"):
            got, want = getattr(a, op)(b), oracle_binary(op, oa, ob)
            assert np.array_equal(device.retrive_data(got.data, O.words(n) * 4).view(np.uint32), want.data), (op, n)
            assert np.array_equal(got.null_buffer.flags(), O.unpack_bits(want.valid, n))
        got, want = a.bitwise_not(), oracle_unary("bitwise_not", oa)
        assert np.array_equal(device.retrive_data(got.data, O.words(n) * 4).view(np.uint32), want.data), ("not", n)


def test_i32_abs_power_neg(device):
    rng = np.random.default_rng(11)
    n = 5000
    x = rng.integers(-50, 50, n).astype(np.int32)
    p = rng.integers(-6, 12, n).astype(np.int32)
    x[:4] = [np.iinfo(np.int32).min, 0, -1, 1]
    p[:4] = [1, -3, -5, -7]
    a, b = ag.Int32ArrayGPU.from_numpy(x, None, device), ag.Int32ArrayGPU.from_numpy(p, None, device)
    assert np.array_equal(a.power(b).raw_values(), O.binary(O.POW, O.I32, x, p))
    assert np.array_equal(a.abs().raw_values(), O.unary(O.ABS, O.I32, x))
    f = rand_vals(rng, O.F32, n)
    assert same_f32_bits(ag.Float32ArrayGPU.from_numpy(f, None, device).neg().raw_values(), O.unary(O.NEG, O.F32, f))


# ---- f32 transcendentals: stated ULP bounds against the correctly rounded value -------------
# (op, input range, max ULP).  sqrt/abs/neg must be exact.  Bounds are the documented CUDA libm
# bounds (CUDA C Programming Guide, "Mathematical Functions"): sinf/cosf 2, expf/exp2f 2, logf 1,
# log2f 1, powf 4, sinhf 3, acosf 2; cbrt = cbrtf + an exponent correction towards the reference's pow(|x|, 1/3f): 2.
ULP_CASES = [("sqrt", (0, 1e6), 0), ("exp", (-20, 20), 2), ("exp2", (-30, 30), 2), ("log", (1e-6, 1e6), 1),
             ("log2", (1e-6, 1e6), 1), ("sin", (-100, 100), 2), ("cos", (-100, 100), 2), ("acos", (-1, 1), 2),
             ("sinh", (-10, 10), 3), ("cbrt", (-1e6, 1e6), 2), ("abs", (-1e6, 1e6), 0)]


@pytest.mark.parametrize("op,rng_,bound", ULP_CASES, ids=[c[0] for c in ULP_CASES])
def test_f32_math_ulp(op, rng_, bound, device):
    rng = np.random.default_rng(30)
    n = 1 << 20
    x = rng.uniform(rng_[0], rng_[1], n).astype(np.float32)
    got = getattr(ag.Float32ArrayGPU.from_numpy(x, None, device), op)().raw_values()
    want = oracle_unary(op, OArr(O.F32, x, n)).raw_values()
    d = ulp_diff(got, want)
    assert d.max() <= bound, f"{op}: max {d.max()} ULP at x={x[d.argmax()]!r} got {got[d.argmax()]!r} want {want[d.argmax()]!r}"


def test_f32_math_special_values(device):
    x = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, -1.0, 1.0, 1e-45, 3e38], np.float32)
    a = ag.Float32ArrayGPU.from_numpy(x, None, device)
    for op in ("sqrt", "exp", "exp2", "log", "log2", "sin", "cos", "acos", "sinh", "cbrt"):
        got = getattr(a, op)().raw_values()
        want = oracle_unary(op, OArr(O.F32, x, len(x))).raw_values()
        with np.errstate(all="ignore"):
            assert np.array_equal(np.isnan(got), np.isnan(want)), (op, got, want)
            assert np.array_equal(np.isinf(got), np.isinf(want)), (op, got, want)
            fin = np.isfinite(want)
            assert ulp_diff(got[fin], want[fin]).max(initial=0) <= 4, (op, got, want)


def test_f32_power_ulp(device):
    rng = np.random.default_rng(31)
    n = 1 << 18
    x = rng.uniform(0, 50, n).astype(np.float32)
    y = rng.uniform(-4, 4, n).astype(np.float32)
    got = ag.Float32ArrayGPU.from_numpy(x, None, device).power(ag.Float32ArrayGPU.from_numpy(y, None, device)).raw_values()
    want = O.binary(O.POW, O.F32, x, y)
    assert ulp_diff(got, want).max() <= 4


@pytest.mark.parametrize("dtype", [O.I8, O.U8, O.I16, O.U16], ids=lambda d: NAMES[d])
def test_int_trig_fused_cast(dtype, device):
    """trigonometry/compute_shaders/{i8,u8,i16,u16}: cast to f32 fused with sin/cos/sinh"""
    rng = np.random.default_rng(dtype)
    n = 50001
    vals = rand_vals(rng, dtype, n)
    a, oa = make(rng, dtype, n, True, device, vals)
    for op, bound in (("sin", 2), ("cos", 2), ("sinh", 3)):
        got = getattr(a, op)()
        assert type(got) is ag.Float32ArrayGPU
        want = oracle_unary(op, oa)
        d = ulp_diff(got.raw_values(), want.raw_values())
        # |x| up to 65535 leaves CUDA's fast sinf/cosf path above 105615 only; still 2 ULP
        assert d.max() <= bound, (op, NAMES[dtype], d.max())
        assert np.array_equal(got.null_buffer.flags(), O.unpack_bits(want.valid, n))


# ---- unaligned device pointers: same results through the element-wise fallback kernels ----------
def test_unaligned_pointers(device):
    l = _ffi.lib()
    rng = np.random.default_rng(9)
    n = 10007
    for dtype in (O.I8, O.I16, O.F32):
        es = O.NP[dtype].itemsize
        x, y = rand_vals(rng, dtype, n + 4), rand_vals(rng, dtype, n + 4)
        bx, by = device.create_gpu_buffer_with_data(x), device.create_gpu_buffer_with_data(y)
        out = device.create_empty_buffer((n + 4) * es)
        bits = device.create_empty_buffer(O.words(n) * 4)
        off = es  # one element: breaks 16-byte alignment
        _ffi.check(l.agpu_binary(device.handle, O.ADD, dtype, bx.ptr + off, by.ptr + off, out.ptr + off, n,
                                 None, None, None), "binary")
        got = device.retrive_data(out, (n + 4) * es).view(O.NP[dtype])[1:n + 1]
        want = O.binary(O.ADD, dtype, x[1:n + 1], y[1:n + 1])
        assert same_f32_bits(got, want) if dtype == O.F32 else np.array_equal(got, want)
        _ffi.check(l.agpu_compare(device.handle, O.GT, dtype, bx.ptr + off, by.ptr + off, bits.ptr, n,
                                  None, None, None), "compare")
        assert np.array_equal(device.retrive_data(bits, O.words(n) * 4).view(np.uint32),
                              O.compare(O.GT, dtype, x[1:n + 1], y[1:n + 1]))


# ---- fused expression == unfused reference chain, bit for bit ----------------------------------
def test_fused_mul_add_gt(device):
    rng = np.random.default_rng(20)
    for n in SIZES + [1 << 20]:
        cols, ocols = [], []
        for k in range(4):
            g, o = make(rng, O.F32, n, k != 2, device, rng.uniform(-10, 10, n).astype(np.float32))
            cols.append(g)
            ocols.append(o)
        fused = K.fused_mul_add_gt(*cols)
        pipeline = ag.ArrowComputePipeline(device, "chain")
        chain = K.gt_op_dyn(K.add_op_dyn(K.mul_op_dyn(cols[0], cols[1], pipeline), cols[2], pipeline), cols[3], pipeline)
        pipeline.finish()
        want = oracle_binary("gt", oracle_binary("add", oracle_binary("mul", ocols[0], ocols[1]), ocols[2]), ocols[3])
        for got in (fused, chain):
            assert np.array_equal(device.retrive_data(got.data, O.words(n) * 4).view(np.uint32), want.data), n
            assert np.array_equal(device.retrive_data(got.null_buffer.bit_buffer, O.words(n) * 4).view(np.uint32),
                                  want.valid), n


# ---- sum: same tree order as the reference => bit-identical f32 result ---------------------------
def test_sum_of_an_empty_column_is_zero(device):
    """the reference's `while new_length != 1` loop never ends for len 0 (aggregate_kernels.rs:26-44);
    both sides here define the result as 0"""
    for cls, dt in ((ag.Float32ArrayGPU, O.F32), (ag.Int32ArrayGPU, O.I32), (ag.UInt32ArrayGPU, O.U32)):
        got = cls.from_numpy(np.zeros(0, O.NP[dt]), None, device).sum().raw_values()
        assert got.tolist() == [0] and np.asarray(O.sum(dt, np.zeros(0, O.NP[dt]))).item() == 0


def test_sum_bit_exact(device):
    rng = np.random.default_rng(40)
    for n in (1, 255, 256, 257, 65536, 65537, 1_000_003, 5_000_000):
        x = rng.uniform(-1000, 1000, n).astype(np.float32)
        got = ag.Float32ArrayGPU.from_numpy(x, None, device).sum().raw_values()[0]
        assert np.float32(got).view(np.uint32) == np.float32(O.sum(O.F32, x)).view(np.uint32), n
        i = rng.integers(-2**31, 2**31, n).astype(np.int32)
        assert ag.Int32ArrayGPU.from_numpy(i, None, device).sum().raw_values()[0] == O.sum(O.I32, i)


# ---- size-independent properties at BASELINE.json scale (256 Mi rows of i8) -----------------------
def test_large_properties_i8(device):
    n = 1 << 28
    rng = np.random.default_rng(10)
    x = rng.integers(-128, 128, n, dtype=np.int8)
    y = rng.integers(-128, 128, n, dtype=np.int8)
    a, b = ag.Int8ArrayGPU.from_numpy(x, None, device), ag.Int8ArrayGPU.from_numpy(y, None, device)
    # (a + b) - b == a under wrap-around; xor is an involution; gt/lteq are complementary
    assert np.array_equal(a.add(b).sub(b).raw_values(), x)
    assert np.array_equal(a.bitwise_xor(b).bitwise_xor(b).raw_values(), x)
    gt, le = a.gt(b), a.lteq(b)
    assert gt.bitwise_xor(le).all() is True
    assert gt.bitwise_and(le).any() is False
    # checksum of the sum against numpy on a strided sample + exact count of a > b
    s = a.add(b).raw_values()
    idx = np.arange(0, n, 9973)
    assert np.array_equal(s[idx], (x[idx].astype(np.int16) + y[idx]).astype(np.int8))
    assert int(gt.raw_values().sum()) == int((x > y).sum())


def test_more_than_2_32_rows(device):
    """64-bit row indexing: n = 2^32 + 77 int8 rows (4.3 GB per column).  add / gt / cast / filter are
    checked against numpy over the whole column (bitmap bit indices and byte offsets exceed 32 bits)."""
    n = (1 << 32) + 77
    rng = np.random.default_rng(77)
    x = rng.integers(-128, 128, n, dtype=np.int8)
    y = np.roll(x, 12345)          # a second column without another RNG pass
    a, b = ag.Int8ArrayGPU.from_numpy(x, None, device), ag.Int8ArrayGPU.from_numpy(y, None, device)
    s = a.add(b).raw_values()
    assert np.array_equal(s[-(1 << 20):], (x[-(1 << 20):].astype(np.int16) + y[-(1 << 20):]).astype(np.int8))
    assert np.array_equal(s[:: 65537], (x[:: 65537].astype(np.int16) + y[:: 65537]).astype(np.int8))
    del s
    g = a.gt(b)
    bits = np.unpackbits(device.retrive_data(g.data, O.words(n) * 4), bitorder="little")[:n].view(bool)
    want = x > y
    assert np.array_equal(bits[-(1 << 20):], want[-(1 << 20):])
    assert int(bits.sum()) == int(want.sum())
    # keep ~0.1 % of the rows, including the very last one
    keep = (x == 127) & (y > 100)
    keep[-1] = True
    m = ag.BooleanArrayGPU(device.create_gpu_buffer_with_data(np.packbits(keep, bitorder="little")), device, n, None)
    out = a.filter(m).raw_values()
    assert np.array_equal(out, x[keep])
    tail = ag.Int8ArrayGPU(ag.ArrowGpuBuffer(device, a.data.ptr + (1 << 32), 77, owned=False), device, 77, None)
    assert np.array_equal(tail.cast(ag.Int32ArrayGPU).raw_values(), x[1 << 32:].astype(np.int32))


# ---- general fused chains: bit-identical to the same ops applied one by one ---------------------
CHAINS = [
    [("mul", "B"), ("add", "C"), ("gt", "D")],
    [("mul", 2.5), ("add", "B"), ("sqrt",), ("lt", 30.0)],
    [("abs",), ("sqrt",), ("mul", "B"), ("sub", 1.0), ("max", "C")],
    [("sin",), ("mul", "B"), ("add", "C"), ("cos",), ("gteq", 0.25)],
    [("div", "B"), ("min", 3.0), ("exp",), ("log2",), ("neg",), ("rem", 0.37), ("add", "C"), ("eq", "D")],
    [("exp2",)],
    [("power", 2.0), ("add", 1.0), ("log",), ("sinh",)],
]


def _unfused(a, steps, cols, device):
    cur = a if isinstance(a, ag.Float32ArrayGPU) else a.cast(ag.Float32ArrayGPU)
    for step in steps:
        name = step[0]
        if len(step) == 1:
            cur = getattr(cur, name)()
            continue
        rhs = cols[step[1]] if isinstance(step[1], str) else step[1]
        if isinstance(rhs, ag.Float32ArrayGPU):
            cur = getattr(cur, name)(rhs)
        elif name in ("add", "sub", "mul", "div", "rem"):
            cur = getattr(cur, name + "_scalar")(ag.Float32ArrayGPU.from_slice([rhs], device))
        else:  # min/max/power/compare with a constant: the unfused API needs a broadcast column
            cur = getattr(cur, name)(ag.Float32ArrayGPU.broadcast(rhs, cur.len, device))
    return cur


@pytest.mark.parametrize("in_dtype", [O.F32, O.I8, O.U8, O.I16, O.U16], ids=lambda d: NAMES[d])
@pytest.mark.parametrize("chain", range(len(CHAINS)))
def test_fused_chain_equals_unfused(chain, in_dtype, device):
    rng = np.random.default_rng(1000 * chain + in_dtype)
    steps = CHAINS[chain]
    for n in (0, 1, 5, 33, 4097, 70001):
        if in_dtype == O.F32:
            vals = rng.uniform(-8, 8, n).astype(np.float32)
            if n >= 8:
                vals[:8] = [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-40, 7.5, -7.5]
        else:
            vals = rand_vals(rng, in_dtype, n)
        a, _ = make(rng, in_dtype, n, True, device, vals)
        cols = {}
        for k, name in enumerate("BCD"):
            cols[name], _ = make(rng, O.F32, n, k != 1, device, rng.uniform(-4, 4, n).astype(np.float32))
        fused = K.fused_chain(a, [(s[0], cols[s[1]]) if len(s) > 1 and isinstance(s[1], str) else s for s in steps])
        want = _unfused(a, steps, cols, device)
        assert type(fused) is type(want) and fused.len == want.len == n
        if isinstance(want, ag.BooleanArrayGPU):
            assert np.array_equal(device.retrive_data(fused.data, O.words(n) * 4), device.retrive_data(want.data, O.words(n) * 4))
        else:
            assert same_f32_bits(fused.raw_values(), want.raw_values()), (chain, n)
        assert np.array_equal(device.retrive_data(fused.null_buffer.bit_buffer, O.words(n) * 4),
                              device.retrive_data(want.null_buffer.bit_buffer, O.words(n) * 4))


def test_fused_chain_vs_oracle_and_limits(device):
    rng = np.random.default_rng(77)
    n = 100_003
    a, oa = make(rng, O.F32, n, True, device, rng.uniform(-10, 10, n).astype(np.float32))
    b, ob = make(rng, O.F32, n, True, device, rng.uniform(-10, 10, n).astype(np.float32))
    got = K.fused_chain(a, [("mul", b), ("add", 0.5), ("abs",), ("sqrt",), ("gt", 2.0)])
    t = oracle_unary("sqrt", oracle_unary("abs", oracle_scalar("add_scalar", oracle_binary("mul", oa, ob), OArr.from_slice(O.F32, [0.5]))))
    want = O.compare(O.GT, O.F32, t.data, np.full(n, 2.0, np.float32))
    assert np.array_equal(device.retrive_data(got.data, O.words(n) * 4).view(np.uint32), want)
    with pytest.raises(ag.Panic):
        K.fused_chain(a, [("gt", b), ("add", 1.0)])          # a compare must end the chain
    with pytest.raises(ag.Panic):
        K.fused_chain(ag.Int32ArrayGPU.from_slice([1], device), [("sqrt",)])
    with pytest.raises(ag.Panic):
        K.fused_chain(a, [("neg",)] * 9)


# ---- auto-fusing pipeline: same bits as the unfused pipeline, fewer launches ---------------------
def _record(pipeline, a, b, c, d, s, i8):
    """a recorded program in the reference's `*_op_dyn` style (crates/arrow/examples/simple.rs:45-72)"""
    r1 = K.mul_op_dyn(a, b, pipeline)
    r2 = K.add_op_dyn(r1, c, pipeline)
    m1 = K.gt_op_dyn(r2, d, pipeline)                         # chain 1: mul, add, gt
    t1 = K.sqrt_op_dyn(K.abs_op_dyn(K.mul_scalar_op_dyn(a, s, pipeline), pipeline), pipeline)
    t2 = K.sub_op_dyn(t1, b, pipeline)                        # chain 2: mul_scalar, abs, sqrt, sub  (kept as f32)
    u1 = K.sin_op_dyn(i8, pipeline)                           # chain 3 starts at an int8 column (fused cast)
    u2 = K.max_op_dyn(K.add_scalar_op_dyn(u1, s, pipeline), c, pipeline)
    w = K.add_op_dyn(t2, u2, pipeline)                        # both operands are recorded chains
    m2 = K.lteq_op_dyn(K.exp_op_dyn(w, pipeline), d, pipeline)
    both = K.bitwise_and_op_dyn(m1, m2, pipeline)             # not fusable: plain kernel on two bitmaps
    return {"r2": r2, "m1": m1, "t2": t2, "u2": u2, "w": w, "m2": m2, "both": both}


def test_auto_fusing_pipeline(device):
    rng = np.random.default_rng(99)
    for n in (0, 7, 4097, 250_001):
        cols = [make(rng, O.F32, n, k % 2 == 0, device, rng.uniform(-5, 5, n).astype(np.float32))[0] for k in range(4)]
        s = ag.Float32ArrayGPU.from_slice([1.75], device)
        i8 = make(rng, O.I8, n, True, device)[0]
        plain = ag.ArrowComputePipeline(device, "plain")
        want = _record(plain, *cols, s, i8)
        plain.finish()
        l0 = device.launch_count()
        fusing = ag.ArrowComputePipeline(device, "fused", fuse=True)
        got = _record(fusing, *cols, s, i8)
        fusing.finish()
        fused_launches = device.launch_count() - l0
        for key in want:
            g, w = got[key], want[key]
            assert type(g) is type(w) and g.len == w.len == n, key
            if isinstance(w, ag.BooleanArrayGPU):
                assert np.array_equal(device.retrive_data(g.data, O.words(n) * 4), device.retrive_data(w.data, O.words(n) * 4)), (key, n)
            else:
                assert same_f32_bits(g.raw_values(), w.raw_values()), (key, n)
            gv = None if g.null_buffer is None else device.retrive_data(g.null_buffer.bit_buffer, O.words(n) * 4)
            wv = None if w.null_buffer is None else device.retrive_data(w.null_buffer.bit_buffer, O.words(n) * 4)
            assert (gv is None) == (wv is None) and (gv is None or np.array_equal(gv, wv)), (key, n)
        if n:
            # 15 ops recorded; fused: chains end at m1, t2, u2, w, m2 (+ r2 read back on demand) + the bitmap AND
            assert fused_launches <= 8, fused_launches


def test_auto_fusion_launch_counts(device):
    n = 1 << 20
    rng = np.random.default_rng(5)
    a, b, c, d = (ag.Float32ArrayGPU.from_numpy(rng.uniform(-3, 3, n).astype(np.float32), None, device) for _ in range(4))
    l0 = device.launch_count()
    p = ag.ArrowComputePipeline(device, fuse=True)
    m = K.gt_op_dyn(K.add_op_dyn(K.mul_op_dyn(a, b, p), c, p), d, p)
    p.finish()
    assert device.launch_count() - l0 == 1                 # mul, add, gt -> one kernel
    assert np.array_equal(m.raw_values(), (a.raw_values() * b.raw_values() + c.raw_values()) > d.raw_values())
    # a result nobody keeps is never computed; a kept intermediate is computed on demand
    l0 = device.launch_count()
    p = ag.ArrowComputePipeline(device, fuse=True)
    r1 = K.mul_op_dyn(a, b, p)
    r2 = K.sqrt_op_dyn(K.abs_op_dyn(r1, p), p)
    del r2
    p.finish()
    assert device.launch_count() - l0 == 0
    assert np.array_equal(r1.raw_values(), a.raw_values() * b.raw_values())
    assert device.launch_count() - l0 == 1
    # more than 8 steps / 3 operand columns split into several kernels, still identical
    p = ag.ArrowComputePipeline(device, fuse=True)
    x = a
    for k in range(11):
        x = K.add_op_dyn(K.mul_scalar_op_dyn(x, ag.Float32ArrayGPU.from_slice([0.5], device), p), (b, c, d)[k % 3], p)
    p.finish()
    ref = a.raw_values()
    for k in range(11):
        ref = ref * np.float32(0.5) + (b, c, d)[k % 3].raw_values()
    assert same_f32_bits(x.raw_values(), ref)


# ---- integer chains (agpu_fused_chain_int) --------------------------------------------------------
INT_CHAINS = [
    [("add", "col"), ("bitwise_and", "col"), ("mul", "scalar")],
    [("bitwise_not",), ("sub", "col"), ("min", "col"), ("max", "col")],
    [("mul", "col"), ("div", "scalar"), ("rem", "scalar"), ("bitwise_xor", "col"), ("gt", "col")],
    [("bitwise_or", "scalar"), ("div", "col"), ("lteq", "scalar")],
    [("rem", "col"), ("eq", "col")],
    [("bitwise_shl", "cnt"), ("add", "col")],
    [("add", "col"), ("bitwise_shr", "cnt"), ("bitwise_and", "scalar"), ("gt", "col")],
    [("bitwise_not",), ("bitwise_shr", "cnt")],
]


@pytest.mark.parametrize("dtype", list(INT_CLS), ids=lambda d: NAMES[d])
@pytest.mark.parametrize("chain", range(len(INT_CHAINS)))
def test_fused_chain_int_equals_unfused_and_oracle(chain, dtype, device):
    rng = np.random.default_rng(stable_seed("intchain", chain, dtype))
    for n in (0, 1, 33, 4099, 70001):
        a, oa = make(rng, dtype, n, True, device)
        steps, want, o = [], a, oa
        for step in INT_CHAINS[chain]:
            name = step[0]
            if len(step) == 1:
                want, o = getattr(want, name)(), oracle_unary(name, o)
                steps.append(step)
                continue
            if step[1] == "cnt":          # per-row u32 counts, above the lane width too (the `& 31` rule)
                cnt = rng.integers(0, 40, n).astype(np.uint32)
                cvalid = rng.random(n) < 0.9 if rng.random() < 0.5 else None
                gc = ag.UInt32ArrayGPU.from_numpy(cnt, cvalid, device)
                oc = OArr(O.U32, cnt, n, None if cvalid is None else O.pack_bits(cvalid))
                steps.append((name, gc))
                want, o = getattr(want, name)(gc), oracle_binary(name, o, oc)
                continue
            if step[1] == "col":
                g, og = make(rng, dtype, n, bool(rng.random() < 0.5), device)
                if name in ("div", "rem") and n > 8:
                    og.data[7] = 0          # the divide-by-zero rules inside a chain
                    g = ALL_CLS[dtype].from_numpy(og.data, None if og.valid is None else O.unpack_bits(og.valid, n), device)
                steps.append((name, g))
                if name == "rem":            # array % array has no method name in the reference: C ABI id
                    want = K._binary(ag._ffi.REM, want, g, what="rem")
                    o = OArr(dtype, O.binary(O.REM, dtype, o.data, og.data), n, O.validity_and(o.valid, og.valid, n))
                else:
                    want, o = getattr(want, name)(g), oracle_binary(name, o, og)
            else:
                sv = rand_vals(rng, dtype, 1, special=False)
                if name in ("div", "rem") and rng.random() < 0.3:
                    sv[0] = 0
                gs = ALL_CLS[dtype].from_numpy(sv, None, device)
                steps.append((name, K.DeviceScalar(gs)))
                if name in ("add", "sub", "mul", "div", "rem"):
                    want, o = getattr(want, name + "_scalar")(gs), oracle_scalar(name + "_scalar", o, OArr(dtype, sv, 1))
                else:                        # min/max/logic/compare with a scalar: broadcast column on the unfused side
                    col = ALL_CLS[dtype].from_numpy(np.full(n, sv[0]), None, device)
                    ocol = OArr(dtype, np.full(n, sv[0]).astype(O.NP[dtype]), n)
                    want, o = getattr(want, name)(col), oracle_binary(name, o, ocol)
        l0 = device.launch_count()
        got = K.fused_chain_int(a, steps)
        assert device.launch_count() - l0 == (1 if n else 0)
        if o.dtype == O.BOOL:
            words = O.words(n) * 4
            assert np.array_equal(device.retrive_data(got.data, words), device.retrive_data(want.data, words)), (chain, n)
            assert np.array_equal(device.retrive_data(got.data, words).view(np.uint32), o.data), (chain, n)
            assert np.array_equal(device.retrive_data(got.null_buffer.bit_buffer, words).view(np.uint32), o.valid), (chain, n)
        else:
            assert_same(got, o, f"int chain {chain} {NAMES[dtype]} n={n}")
            assert np.array_equal(got.raw_values(), want.raw_values())


def test_fused_chain_int_limits(device):
    a = ag.Int16ArrayGPU.from_slice([1, 2, 3], device)
    with pytest.raises(ag.Panic):
        K.fused_chain_int(a, [("gt", a), ("add", a)])
    with pytest.raises(ag.Panic):
        K.fused_chain_int(a, [("add", ag.Int8ArrayGPU.from_slice([1, 2, 3], device))])
    with pytest.raises(ag.Panic):
        K.fused_chain_int(ag.Float32ArrayGPU.from_slice([1.0], device), [("bitwise_not",)])
    with pytest.raises(ag.Panic):
        K.fused_chain_int(a, [("abs",)])                    # abs is an Int32 op (math/src/i32.rs)
    assert K.fused_chain_int(a, [("add", 5), ("mul", a)]).raw_values().tolist() == [6, 14, 24]
    cnt = ag.UInt32ArrayGPU.from_slice([1, 2, 35], device)
    assert K.fused_chain_int(a, [("bitwise_shl", cnt), ("add", 1)]).raw_values().tolist() == [3, 9, 25]   # 35 & 31 = 3
    with pytest.raises(ag._ffi.AgpuError):
        K.fused_chain_int(a, [("bitwise_shl", cnt), ("bitwise_shr", cnt)])      # one shift step per chain
    i32 = ag.Int32ArrayGPU.from_slice([-3, 2, -2**31], device)
    assert K.fused_chain_int(i32, [("abs",), ("power", ag.Int32ArrayGPU.from_slice([2, 3, 1], device))]).raw_values().tolist() \
        == [9, 8, -2**31]


def test_auto_fusion_int_chain_launch_counts(device):
    rng = np.random.default_rng(5)
    n = 100_001
    a, b, c = (make(rng, O.U8, n, k == 0, device)[0] for k in range(3))
    s = ag.UInt8ArrayGPU.from_slice([7], device)

    def program(p):
        t = K.bitwise_and_op_dyn(K.add_op_dyn(a, b, p), c, p)
        return K.gt_op_dyn(K.mul_scalar_op_dyn(t, s, p), b, p), t

    plain = ag.ArrowComputePipeline(device)
    l0 = device.launch_count()
    want_m, want_t = program(plain)
    plain.finish()
    n_plain = device.launch_count() - l0
    fusing = ag.ArrowComputePipeline(device, fuse=True)
    l0 = device.launch_count()
    got_m, got_t = program(fusing)
    fusing.finish()
    n_fused = device.launch_count() - l0
    assert n_plain == 4 and n_fused == 1, (n_plain, n_fused)     # add, and, mul_scalar, gt in ONE kernel
    words = O.words(n) * 4
    assert np.array_equal(device.retrive_data(got_m.data, words), device.retrive_data(want_m.data, words))
    assert np.array_equal(device.retrive_data(got_m.null_buffer.bit_buffer, words),
                          device.retrive_data(want_m.null_buffer.bit_buffer, words))
    l0 = device.launch_count()
    assert np.array_equal(got_t.raw_values(), want_t.raw_values())   # absorbed prefix: launched when read
    assert device.launch_count() - l0 == 1
