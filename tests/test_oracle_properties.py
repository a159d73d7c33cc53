"""CPU: algebraic properties of the oracle (oracle/oracle.c) on random inputs — a second line of
defence next to the reference's golden vectors: an oracle that violated one of these could not be
the reference's algorithm.  hypothesis drives sizes/seeds; everything is bit-exact integer work."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle as O

INTS = [O.I8, O.U8, O.I16, O.U16, O.I32, O.U32]
WIDTH = {O.I8: 8, O.U8: 8, O.I16: 16, O.U16: 16, O.I32: 32, O.U32: 32}
sizes = st.sampled_from([0, 1, 2, 31, 32, 33, 100, 257, 1000])
seeds = st.integers(0, 2**31 - 1)
dtypes = st.sampled_from(INTS)
fast = settings(max_examples=60, deadline=None)
pytestmark = pytest.mark.timeout(120)


def ints(rng, dtype, n):
    info = np.iinfo(O.NP[dtype])
    x = rng.integers(info.min, int(info.max) + 1, n, dtype=np.int64).astype(O.NP[dtype])
    if n >= 4:
        x[:4] = np.array([0, info.min, info.max, 1]).astype(O.NP[dtype])
    return x


@fast
@given(dtypes, sizes, seeds)
def test_add_sub_are_inverse_and_commutative(dtype, n, seed):
    rng = np.random.default_rng(seed)
    a, b = ints(rng, dtype, n), ints(rng, dtype, n)
    s = O.binary(O.ADD, dtype, a, b)
    assert np.array_equal(O.binary(O.SUB, dtype, s, b), a)                 # wrap-around is a group
    assert np.array_equal(s, O.binary(O.ADD, dtype, b, a))
    assert np.array_equal(O.binary(O.MUL, dtype, a, b), O.binary(O.MUL, dtype, b, a))
    wide = (a.astype(np.int64) + b.astype(np.int64)) & ((1 << WIDTH[dtype]) - 1)
    assert np.array_equal(s.astype(np.int64) & ((1 << WIDTH[dtype]) - 1), wide)


@fast
@given(dtypes, sizes, seeds)
def test_div_rem_reconstruct_the_dividend(dtype, n, seed):
    """a == (a / s) * s + (a % s) for s != 0 (and not MIN / -1); x / 0 = x and x % 0 = 0 (Q12)"""
    rng = np.random.default_rng(seed)
    a = ints(rng, dtype, n)
    s = int(ints(rng, dtype, 8)[5])
    q, r = O.scalar(O.DIV, dtype, a, s), O.scalar(O.REM, dtype, a, s)
    if s == 0:
        assert np.array_equal(q, a) and not r.any()
        return
    back = O.binary(O.ADD, dtype, O.scalar(O.MUL, dtype, q, s), r)
    assert np.array_equal(back, a)
    if s > 0 and n:
        assert (np.abs(r.astype(np.int64)) < s).all()
        assert ((r.astype(np.int64) == 0) | (np.sign(r.astype(np.int64)) == np.sign(a.astype(np.int64)))).all()


@fast
@given(dtypes, sizes, seeds)
def test_compare_relations(dtype, n, seed):
    rng = np.random.default_rng(seed)
    a, b = ints(rng, dtype, n), ints(rng, dtype, n)
    if n > 5:
        b[5:n:3] = a[5:n:3]                                                # some equal pairs
    gt, lt = O.compare(O.GT, dtype, a, b), O.compare(O.LT, dtype, b, a)
    ge, le = O.compare(O.GTEQ, dtype, a, b), O.compare(O.LTEQ, dtype, a, b)
    eq = O.compare(O.EQ, dtype, a, b)
    assert np.array_equal(gt, lt)
    assert np.array_equal(O.unpack_bits(gt, n), a > b)
    assert np.array_equal(O.unpack_bits(ge, n), a >= b)
    assert np.array_equal(O.unpack_bits(ge, n) & O.unpack_bits(le, n), O.unpack_bits(eq, n))
    last = O.words(n) * 32 - n                                              # padding bits are zero
    if n and last:
        assert gt[-1] >> (32 - last) == 0 and eq[-1] >> (32 - last) == 0
    mn, mx = O.binary(O.MIN, dtype, a, b), O.binary(O.MAX, dtype, a, b)
    assert np.array_equal(mn, np.minimum(a, b)) and np.array_equal(mx, np.maximum(a, b))


@fast
@given(dtypes, sizes, seeds)
def test_logic_and_shift_identities(dtype, n, seed):
    rng = np.random.default_rng(seed)
    a, b = ints(rng, dtype, n), ints(rng, dtype, n)
    nota = O.unary(O.NOT, dtype, a)
    assert np.array_equal(O.unary(O.NOT, dtype, nota), a)
    assert np.array_equal(O.binary(O.XOR, dtype, O.binary(O.XOR, dtype, a, b), b), a)
    # De Morgan
    assert np.array_equal(O.unary(O.NOT, dtype, O.binary(O.AND, dtype, a, b)),
                          O.binary(O.OR, dtype, nota, O.unary(O.NOT, dtype, b)))
    w = WIDTH[dtype]
    c = rng.integers(0, w, n).astype(np.uint32)
    shl = O.shift(O.SHL, dtype, a, c)
    want = ((a.astype(np.int64) << c.astype(np.int64)) & ((1 << w) - 1))
    assert np.array_equal(shl.astype(np.int64) & ((1 << w) - 1), want)
    shr = O.shift(O.SHR, dtype, a, c)
    assert np.array_equal(shr.astype(np.int64), a.astype(np.int64) >> c.astype(np.int64))   # arithmetic for signed
    assert np.array_equal(O.shift(O.SHL, dtype, a, c + 32), shl)                               # count & 31


@fast
@given(st.sampled_from([O.I8, O.U8, O.I16, O.U16, O.I32, O.U32, O.F32]), sizes, seeds)
def test_merge_take_put_filter(dtype, n, seed):
    rng = np.random.default_rng(seed)
    if dtype == O.F32:
        a, b = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    else:
        a, b = ints(rng, dtype, n), ints(rng, dtype, n)
    flags = rng.random(n) < 0.4
    m = O.pack_bits(flags)
    assert np.array_equal(O.merge(dtype, a, b, m, n), np.where(flags, a, b))
    assert np.array_equal(O.merge(dtype, a, a, m, n), a)
    perm = rng.permutation(n).astype(np.uint32)
    taken = O.take(dtype, a, n, perm)
    assert np.array_equal(taken, a[perm])
    if dtype in (O.I32, O.U32, O.F32) and n:
        back = O.put(dtype, taken, np.arange(n, dtype=np.uint32), np.zeros(n, O.NP[dtype]), perm)
        assert np.array_equal(back, a)                                      # put undoes take
    vflags = rng.random(n) < 0.8
    mflags = rng.random(n) < 0.9
    out, vout, k = O.filter(dtype, a, O.pack_bits(vflags), m, O.pack_bits(mflags))
    keep = flags & mflags
    assert k == int(keep.sum())
    assert np.array_equal(out[:k], a[keep])
    assert np.array_equal(O.unpack_bits(vout, k), vflags[keep])
    # filtering twice with an all-true mask changes nothing
    again, _v, k2 = O.filter(dtype, out[:k], None, O.pack_bits(np.ones(k, bool)), None)
    assert k2 == k and np.array_equal(again[:k], out[:k])


@fast
@given(sizes, seeds)
def test_casts_and_sums(n, seed):
    rng = np.random.default_rng(seed)
    i8, u8, i16, u16 = (ints(rng, t, n) for t in (O.I8, O.U8, O.I16, O.U16))
    assert np.array_equal(O.cast(O.I8, O.I32, i8), i8.astype(np.int32))
    assert np.array_equal(O.cast(O.I8, O.U16, i8), i8.astype(np.int16).view(np.uint16))     # sign-extend, keep low bits
    assert np.array_equal(O.cast(O.U8, O.I16, u8), u8.astype(np.int16))
    assert np.array_equal(O.cast(O.I16, O.F32, i16), i16.astype(np.float32))
    assert np.array_equal(O.cast(O.F32, O.U8, O.cast(O.U8, O.F32, u8)), u8)               # exact round trip
    assert np.array_equal(O.cast(O.U16, O.U32, u16), u16.astype(np.uint32))
    i32 = ints(rng, O.I32, n)
    assert int(np.asarray(O.sum(O.I32, i32)).astype(np.int64)) == int(((int(i32.astype(np.int64).sum()) + 2**31) % 2**32) - 2**31)
    small = rng.integers(-1000, 1000, n).astype(np.float32)                 # exactly representable partial sums
    assert float(np.asarray(O.sum(O.F32, small))) == float(small.astype(np.float64).sum())


def test_i16_shr_helper_is_an_arithmetic_shift_for_every_count():
    """logical/compute_shaders/i16/shift.wgsl:31-41 evaluated literally (the oracle does) equals
    `lo16(sx(a) >> (count & 31))` for ALL counts, also where `16u - shift_value` wraps — so the
    kernel's formula and the shader agree beyond the pinned 0..15 range."""
    vals = np.array([-32768, -32767, -12345, -2, -1, 0, 1, 2, 12345, 32767], dtype=np.int16)
    counts = np.array(list(range(0, 70)) + [255, 256, 2**31, 2**32 - 1], dtype=np.uint32)
    a = np.repeat(vals, len(counts))
    c = np.tile(counts, len(vals))
    got = O.shift(O.SHR, O.I16, a, c)
    want = (a.astype(np.int32) >> (c & 31).astype(np.int32)).astype(np.int16)
    assert np.array_equal(got, want)
