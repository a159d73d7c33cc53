"""GPU: agpu_fused_chain_pair — a value chain and a predicate chain of one source column in ONE
kernel (BASELINE.json configs[0]: s = a + b; g = a > b).  Both results must be bit-identical to the
two chains run separately, and to the oracle; the auto-fusing pipeline must produce the same arrays
with one launch instead of two."""
import numpy as np
import pytest

import arrow_gpu_b200 as ag
from arrow_gpu_b200 import kernels as K
import oracle as O
from helpers import OArr, oracle_binary

pytestmark = pytest.mark.gpu

SPECIALS = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-40, 7.5, -7.5, 3.4e38, -3.4e38], dtype=np.float32)


def column(rng, n, with_validity, device, lo=-8, hi=8):
    vals = rng.uniform(lo, hi, n).astype(np.float32)
    if n >= len(SPECIALS):
        vals[rng.permutation(n)[: len(SPECIALS)]] = SPECIALS
    valid = (rng.random(n) < 0.85) if with_validity else None
    return ag.Float32ArrayGPU.from_numpy(vals, valid, device), vals, valid


def same_bits(x, y):
    x, y = np.asarray(x, np.float32), np.asarray(y, np.float32)
    nan = np.isnan(x) & np.isnan(y)            # any NaN equals any NaN (tests/helpers.py, test_macros/src/lib.rs:89-94)
    return np.array_equal(np.where(nan, 0, x.view(np.uint32)), np.where(nan, 0, y.view(np.uint32)))


PROGRAMS = [
    ([("add", "B")], [("gt", "B")]),                                             # config 1
    ([("mul", "B"), ("add", "C")], [("sub", "D"), ("lteq", "B")]),               # three distinct columns
    ([("abs",), ("sqrt",), ("div", "B")], [("neg",), ("min", "C"), ("eq", 2.5)]),
    ([("rem", 3.0), ("max", "B")], [("mul", 0.5), ("gteq", "B")]),
    ([("sub", "S")], [("add", "S"), ("lt", "S")]),                               # one-element device scalars
]


def resolve(steps, cols):
    out = []
    for st in steps:
        if len(st) > 1 and isinstance(st[1], str):
            out.append((st[0], K.DeviceScalar(cols["S"]) if st[1] == "S" else cols[st[1]]))
        else:
            out.append(st)
    return out


@pytest.mark.parametrize("prog", range(len(PROGRAMS)))
@pytest.mark.parametrize("validity", ["none", "source", "all"])
def test_pair_equals_the_two_chains_run_separately(prog, validity, device):
    """validity "source": only the source column has a bitmap; "all": every column has one — then
    the two chains of some programs depend on different bitmaps and must be refused"""
    rng = np.random.default_rng(50 + prog)
    for n in (0, 1, 5, 33, 1023, 4097, 70001, (1 << 20) + 3):
        a, _, _ = column(rng, n, validity != "none", device)
        cols = {k: column(rng, n, validity == "all", device, -4, 4)[0] for k in "BCD"}
        cols["S"] = ag.Float32ArrayGPU.from_slice([1.25], device)
        vs, ps = resolve(PROGRAMS[prog][0], cols), resolve(PROGRAMS[prog][1], cols)
        used = [{st[1] for st in chain if len(st) > 1 and st[1] in tuple("BCD")} for chain in PROGRAMS[prog]]
        same_bitmaps = validity != "all" or used[0] == used[1]
        assert K.pair_eligible(a, vs, ps) == same_bitmaps
        if not same_bitmaps:
            with pytest.raises(ag.Panic):
                K.fused_chain_pair(a, vs, ps)
            continue
        value, pred = K.fused_chain_pair(a, vs, ps)
        want_v, want_p = K.fused_chain(a, vs), K.fused_chain(a, ps)
        assert type(value) is ag.Float32ArrayGPU and type(pred) is ag.BooleanArrayGPU and value.len == pred.len == n
        assert same_bits(value.raw_values(), want_v.raw_values()), (prog, n)
        assert np.array_equal(device.retrive_data(pred.data, O.words(n) * 4), device.retrive_data(want_p.data, O.words(n) * 4)), (prog, n)
        if validity != "none":
            for got, want in ((value, want_v), (pred, want_p)):
                assert np.array_equal(device.retrive_data(got.null_buffer.bit_buffer, O.words(n) * 4),
                                      device.retrive_data(want.null_buffer.bit_buffer, O.words(n) * 4)), (prog, n)
        else:
            assert value.null_buffer is None and pred.null_buffer is None


@pytest.mark.parametrize("binop", ["add", "sub", "mul", "div", "min"])
@pytest.mark.parametrize("cmp", ["gt", "gteq", "lt", "lteq", "eq"])
def test_bin_cmp_pairs_take_the_dedicated_kernel_with_identical_results(binop, cmp, device):
    """value = a binop b, predicate = a cmp b / a cmp c (add sub mul div: dedicated kernel; min: interpreter)"""
    rng = np.random.default_rng(hash((binop, cmp)) % 1000)
    for n in (3, 4100, 300_007):
        a, x, _ = column(rng, n, True, device)
        b, y, _ = column(rng, n, True, device)
        c, z, _ = column(rng, n, False, device)
        y[: n // 2] = x[: n // 2]                       # ties, so that gteq / lteq / eq see both outcomes
        b = ag.Float32ArrayGPU.from_numpy(y, None, device)
        for other in (b, c):
            value, pred = K.fused_chain_pair(a, [(binop, b)], [(cmp, other)])
            assert same_bits(value.raw_values(), getattr(a, binop)(b).raw_values()), (binop, cmp, n)
            assert np.array_equal(pred.raw_values(), getattr(a, cmp)(other).raw_values()), (binop, cmp, n)
            assert np.array_equal(value.null_buffer.flags(), a.null_buffer.flags())


def test_config1_program_against_the_oracle_and_unaligned_columns(device):
    rng = np.random.default_rng(1)
    n = (1 << 20) + 77
    a, x, va = column(rng, n, True, device, -1000, 1000)
    b, y, vb = column(rng, n, True, device, -1000, 1000)
    s, g = K.fused_chain_pair(a, [("add", b)], [("gt", b)])
    oa, ob = OArr(O.F32, x, n, O.pack_bits(va)), OArr(O.F32, y, n, O.pack_bits(vb))
    want = oracle_binary("add", oa, ob)
    assert same_bits(s.raw_values(), want.data)
    assert np.array_equal(device.retrive_data(g.data, O.words(n) * 4).view(np.uint32), O.compare(O.GT, O.F32, x, y))
    assert np.array_equal(device.retrive_data(s.null_buffer.bit_buffer, O.words(n) * 4).view(np.uint32), want.valid)
    assert s.null_buffer.bit_buffer is g.null_buffer.bit_buffer
    # an operand that is not 16-byte aligned takes the row-per-thread kernel: same results
    big = ag.Float32ArrayGPU.from_numpy(np.concatenate([[0.0], y]).astype(np.float32), None, device)
    shifted = ag.Float32ArrayGPU(ag.ArrowGpuBuffer(device, big.data.ptr + 4, n * 4, owned=False), device, n, None)
    a2 = ag.Float32ArrayGPU.from_numpy(x, None, device)
    s2, g2 = K.fused_chain_pair(a2, [("add", shifted)], [("gt", shifted)])
    assert same_bits(s2.raw_values(), want.data)
    assert np.array_equal(device.retrive_data(g2.data, O.words(n) * 4).view(np.uint32), O.compare(O.GT, O.F32, x, y))


def test_fusing_pipeline_pairs_add_and_gt_into_one_launch(device):
    rng = np.random.default_rng(9)
    n = (1 << 18) + 5
    a, x, va = column(rng, n, True, device)
    b, y, vb = column(rng, n, True, device)
    plain = ag.ArrowComputePipeline(device)
    s0, g0 = K.add_op_dyn(a, b, plain), K.gt_op_dyn(a, b, plain)
    plain.finish()
    before = device.launch_count()
    p = ag.ArrowComputePipeline(device, fuse=True)
    s1 = K.add_op_dyn(a, b, p)
    g1 = K.gt_op_dyn(a, b, p)
    p.finish()
    assert device.launch_count() - before == 1
    assert same_bits(s1.raw_values(), s0.raw_values())
    assert np.array_equal(g1.raw_values(), g0.raw_values())
    assert np.array_equal(s1.null_buffer.flags(), s0.null_buffer.flags()) and np.array_equal(g1.null_buffer.flags(), g0.null_buffer.flags())
    # recorded once into a CUDA graph, replayed: still one kernel per submit, same results
    cap = ag.ArrowComputePipeline(device, "pair", fuse=True, capture=True)
    s2 = K.add_op_dyn(a, b, cap)
    g2 = K.gt_op_dyn(a, b, cap)
    cap.finish()
    assert cap.graph.kernels == 1
    cap.replay()
    assert same_bits(s2.raw_values(), s0.raw_values()) and np.array_equal(g2.raw_values(), g0.raw_values())
    # the value result keeps working as an operand of later ops
    t = s1.mul(b)
    assert same_bits(t.raw_values(), s0.mul(b).raw_values())
