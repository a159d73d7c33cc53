"""CPU: independent cross-checks of the oracle on random inputs, beyond the reference's 5-11-element
vectors.  (1) numpy, for every op whose semantics numpy shares exactly (two's-complement wrap,
compares, bit ops, widening casts, select/gather); (2) the WGSL edge rules written out as
one-liners (x/0, MIN/-1, shift count & 31, saturating f32->u32); (3) pyarrow.compute for the ops
SURVEY.md §8c lists as agreeing (int8 wrap, NaN compares, NaN-ignoring max, null-AND, if_else,
take, filter).  The oracle's sub-word paths are word-level restatements of the shaders, so this
also proves those restatements equal the per-element semantics."""
import numpy as np
import pytest

import oracle as O

INTS = [O.I8, O.U8, O.I16, O.U16, O.I32, O.U32]
N = 20011


def rnd(rng, dt, n=N):
    if dt == O.F32:
        x = rng.uniform(-1e4, 1e4, n).astype(np.float32)
        x[:6] = [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-40]
        return x
    info = np.iinfo(O.NP[dt])
    x = rng.integers(info.min, int(info.max) + 1, n, dtype=np.int64).astype(O.NP[dt])
    x[:4] = np.array([info.min, info.max, 0, 1]).astype(O.NP[dt])
    return x


@pytest.mark.parametrize("dt", INTS)
def test_wrapping_arithmetic_and_bit_ops_vs_numpy(dt):
    rng = np.random.default_rng(dt)
    a, b = rnd(rng, dt), rnd(rng, dt)
    with np.errstate(over="ignore"):
        assert np.array_equal(O.binary(O.ADD, dt, a, b), a + b)
        assert np.array_equal(O.binary(O.SUB, dt, a, b), a - b)
        assert np.array_equal(O.binary(O.MUL, dt, a, b), a * b)
        assert np.array_equal(O.scalar(O.ADD, dt, a, 3), a + O.NP[dt].type(3))
        assert np.array_equal(O.scalar(O.MUL, dt, a, 3), a * O.NP[dt].type(3))
    assert np.array_equal(O.binary(O.AND, dt, a, b), a & b)
    assert np.array_equal(O.binary(O.OR, dt, a, b), a | b)
    assert np.array_equal(O.binary(O.XOR, dt, a, b), a ^ b)
    assert np.array_equal(O.unary(O.NOT, dt, a), ~a)
    assert np.array_equal(O.binary(O.MIN, dt, a, b), np.minimum(a, b))
    assert np.array_equal(O.binary(O.MAX, dt, a, b), np.maximum(a, b))
    for op, f in ((O.GT, np.greater), (O.GTEQ, np.greater_equal), (O.LT, np.less), (O.LTEQ, np.less_equal), (O.EQ, np.equal)):
        assert np.array_equal(O.unpack_bits(O.compare(op, dt, a, b), N), f(a, b))


@pytest.mark.parametrize("dt", INTS)
def test_wgsl_division_rules(dt):
    """x/0 = x, x%0 = 0, MIN/-1 = MIN, MIN%-1 = 0, otherwise truncation toward zero"""
    rng = np.random.default_rng(100 + dt)
    a, b = rnd(rng, dt), rnd(rng, dt)
    b[:50] = 0
    info = np.iinfo(O.NP[dt])
    if info.min < 0:
        a[50:60], b[50:60] = info.min, -1
    A, B = a.astype(np.int64), b.astype(np.int64)
    safe = np.where(B == 0, 1, B)
    q = np.where(B == 0, A, np.sign(A) * np.sign(safe) * (np.abs(A) // np.abs(safe)))
    r = np.where(B == 0, 0, A - q * safe)
    if info.min < 0:   # the widened MIN/-1 does not overflow in 32 bits for 8/16-bit lanes: it wraps on narrowing
        if dt == O.I32:
            ov = (A == info.min) & (B == -1)
            q, r = np.where(ov, A, q), np.where(ov, 0, r)
    assert np.array_equal(O.binary(O.DIV, dt, a, b), q.astype(O.NP[dt]))
    assert np.array_equal(O.binary(O.REM, dt, a, b), r.astype(O.NP[dt]))


@pytest.mark.parametrize("dt", INTS)
def test_shift_count_mod_32_on_widened_lane(dt):
    rng = np.random.default_rng(200 + dt)
    a = rnd(rng, dt)
    c = rng.integers(0, 64, N).astype(np.uint32)
    wide = a.astype(np.int64)
    s = (c & 31).astype(np.int64)
    bits = O.NP[dt].itemsize * 8
    shl = ((wide << s) & ((1 << bits) - 1)).astype(np.uint64).astype(O.NP[dt] if np.iinfo(O.NP[dt]).min == 0 else f"u{bits // 8}").view(O.NP[dt])
    shr = (wide >> s).astype(O.NP[dt])          # numpy >> on int64 is arithmetic, i.e. sign-filling for signed lanes
    assert np.array_equal(O.shift(O.SHL, dt, a, c), shl)
    assert np.array_equal(O.shift(O.SHR, dt, a, c), shr)


def test_casts_vs_numpy():
    rng = np.random.default_rng(3)
    pairs = [(O.I8, O.U8), (O.I8, O.U16), (O.I8, O.U32), (O.I8, O.I16), (O.I8, O.I32), (O.I8, O.F32), (O.I16, O.I32),
             (O.I16, O.U16), (O.I16, O.U32), (O.I16, O.F32), (O.U8, O.U16), (O.U8, O.U32), (O.U8, O.I8), (O.U8, O.I16),
             (O.U8, O.I32), (O.U8, O.F32), (O.U16, O.U32), (O.U16, O.I16), (O.U16, O.I32), (O.U16, O.F32)]
    for s, d in pairs:
        a = rnd(rng, s)
        assert np.array_equal(O.cast(s, d, a), a.astype(O.NP[d])), (s, d)   # C-style wrap == sign/zero extend + truncate
    f = rng.uniform(-10, 70000, N).astype(np.float32)
    f[:6] = [-1.0, np.nan, np.inf, 255.9, 256.0, 5e9]
    f64 = f.astype(np.float64)
    want = np.where(np.isnan(f64) | (f64 <= 0), 0, np.minimum(np.trunc(np.nan_to_num(f64, nan=0.0, posinf=4294967295.0)), 4294967295.0)).astype(np.uint64) % 256
    assert np.array_equal(O.cast(O.F32, O.U8, f), want.astype(np.uint8))
    flags = rng.random(N) < 0.5
    assert np.array_equal(O.cast(O.BOOL, O.F32, O.pack_bits(flags), N), flags.astype(np.float32))


def test_f32_exact_ops_vs_numpy():
    rng = np.random.default_rng(4)
    a, b = rnd(rng, O.F32), rnd(rng, O.F32)
    b[6:12] = [0.0, -0.0, np.nan, np.inf, 1.0, -1.0]
    same = lambda x, y: np.array_equal(np.isnan(x), np.isnan(y)) and np.array_equal(x[~np.isnan(x)].view(np.uint32), y[~np.isnan(y)].view(np.uint32))  # noqa: E731
    with np.errstate(all="ignore"):
        assert same(O.binary(O.ADD, O.F32, a, b), a + b)
        assert same(O.binary(O.SUB, O.F32, a, b), a - b)
        assert same(O.binary(O.MUL, O.F32, a, b), a * b)
        assert same(O.binary(O.DIV, O.F32, a, b), a / b)
        assert same(O.unary(O.SQRT, O.F32, np.abs(a)), np.sqrt(np.abs(a)))
        assert same(O.unary(O.NEG, O.F32, a), -a)
        assert same(O.unary(O.ABS, O.F32, a), np.abs(a))
        assert same(O.binary(O.MAX, O.F32, a, b), np.fmax(a, b))      # NaN-ignoring
        assert same(O.binary(O.MIN, O.F32, a, b), np.fmin(a, b))
        for op, f in ((O.GT, np.greater), (O.GTEQ, np.greater_equal), (O.LT, np.less), (O.LTEQ, np.less_equal), (O.EQ, np.equal)):
            assert np.array_equal(O.unpack_bits(O.compare(op, O.F32, a, b), N), f(a, b))
    assert O.binary(O.MAX, O.F32, [-0.0], [0.0]).view(np.uint32)[0] == 0            # -0 < +0
    assert O.binary(O.MIN, O.F32, [0.0], [-0.0]).view(np.uint32)[0] == 0x80000000


def test_routines_vs_numpy():
    rng = np.random.default_rng(5)
    for dt in INTS + [O.F32]:
        a, b = rnd(rng, dt), rnd(rng, dt)
        m = rng.random(N) < 0.5
        got = O.merge(dt, a, b, O.pack_bits(m))
        assert np.array_equal(got.view(np.uint8), np.where(m, a, b).view(np.uint8))
        idx = rng.integers(0, N, 3 * N // 2).astype(np.uint32)
        assert np.array_equal(O.take(dt, a, N, idx).view(np.uint8), a[idx].view(np.uint8))
        out, _, k = O.filter(dt, a, None, O.pack_bits(m), None)
        assert k == m.sum() and np.array_equal(out.view(np.uint8), a[m].view(np.uint8))
    va, vb, vm = (O.pack_bits(rng.random(N) < 0.8) for _ in range(3))
    mbits = O.pack_bits(rng.random(N) < 0.5)
    assert np.array_equal(O.merge_validity(va, vb, mbits, vm, N), ((va & mbits) | (vb & ~mbits)) & vm)


def test_sum_tree_order_is_the_reference_shader_order():
    """the pairwise tree of aggregate.wgsl written out with an explicit shared array"""
    rng = np.random.default_rng(6)
    x = rng.uniform(-1000, 1000, 70001).astype(np.float32)

    def shader_pass(v):
        groups = (len(v) + 255) // 256
        pad = np.zeros(groups * 256, np.float32)
        pad[: len(v)] = v
        sh = pad.reshape(groups, 256).copy()
        s = 1
        while s < 256:
            idx = np.arange(0, 256, 2 * s)
            sh[:, idx] = sh[:, idx] + sh[:, idx + s]
            s *= 2
        return sh[:, 0].copy()
    v = x
    while True:
        v = shader_pass(v)
        if len(v) == 1:
            break
    assert np.float32(O.sum(O.F32, x)).view(np.uint32) == v[0].view(np.uint32)
    i = rng.integers(-2**31, 2**31, 70001).astype(np.int32)
    assert O.sum(O.I32, i) == np.int32(i.astype(np.int64).sum() & 0xFFFFFFFF if (i.astype(np.int64).sum() & 0xFFFFFFFF) < 2**31 else (i.astype(np.int64).sum() & 0xFFFFFFFF) - 2**32)


def test_pyarrow_agrees_where_survey_says_it_should():
    pa = pytest.importorskip("pyarrow")
    import pyarrow.compute as pc
    rng = np.random.default_rng(7)
    a, b = rnd(rng, O.I8), rnd(rng, O.I8)
    assert np.array_equal(O.binary(O.ADD, O.I8, a, b), pc.add(pa.array(a), pa.array(b)).to_numpy())
    assert np.array_equal(O.binary(O.MUL, O.I8, a, b), pc.multiply(pa.array(a), pa.array(b)).to_numpy())
    f, g = rnd(rng, O.F32), rnd(rng, O.F32)
    assert np.array_equal(O.unpack_bits(O.compare(O.GT, O.F32, f, g), N), pc.greater(pa.array(f), pa.array(g)).to_numpy(zero_copy_only=False))
    mx = pc.max_element_wise(pa.array(f), pa.array(g)).to_numpy()
    got = O.binary(O.MAX, O.F32, f, g)
    assert np.array_equal(np.isnan(got), np.isnan(mx)) and np.array_equal(got[~np.isnan(got)], mx[~np.isnan(mx)])
    keep, kvalid, valid = rng.random(N) < 0.4, rng.random(N) < 0.9, rng.random(N) < 0.9
    i = rnd(rng, O.I32)
    out, vout, k = O.filter(O.I32, i, O.pack_bits(valid), O.pack_bits(keep), O.pack_bits(kvalid))
    want = pc.filter(pa.array(i, mask=~valid), pa.array(keep, mask=~kvalid), null_selection_behavior="drop")
    got = [int(v) if ok else None for v, ok in zip(out, O.unpack_bits(vout, k))]
    assert got == want.to_pylist()
    sel = pc.if_else(pa.array(keep), pa.array(i), pa.array(i[::-1].copy())).to_numpy()
    assert np.array_equal(O.merge(O.I32, i, i[::-1].copy(), O.pack_bits(keep)), sel)
