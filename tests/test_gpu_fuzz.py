"""GPU: random operator programs through the public array API, checked value-for-value and
bitmap-for-bitmap against the oracle-side model (helpers.OArr) after every step — eagerly and
recorded on a fusing ArrowComputePipeline.  Only bit-exact ops take part (integer, bitmap, cast,
indexing, f32 + - * / min max neg abs sqrt and compares); transcendentals have their own ULP tests."""
import numpy as np
import pytest

import arrow_gpu_b200 as ag
import oracle as O
from helpers import (OArr, oracle_binary, oracle_cast, oracle_filter, oracle_merge, oracle_scalar, oracle_take,
                     oracle_unary)
from test_gpu_parity import ALL_CLS, NAMES, assert_same, make, rand_vals

pytestmark = pytest.mark.gpu

INTS = [O.I8, O.U8, O.I16, O.U16, O.I32, O.U32]
LENGTHS = [1, 2, 31, 32, 33, 255, 1000, 4095, 4096, 4097, 20011]
CASTS = {O.I8: [O.U8, O.U16, O.U32, O.I16, O.I32, O.F32], O.I16: [O.I32, O.U16, O.U32, O.F32],
         O.U8: [O.U16, O.U32, O.I8, O.I16, O.I32, O.F32], O.U16: [O.U32, O.I16, O.I32, O.F32], O.F32: [O.U8]}


def make_bool(rng, n, nulls, device):
    flags = rng.random(n) < rng.choice([0.1, 0.5, 0.9])
    valid = rng.random(n) < 0.9 if nulls else None
    return (ag.BooleanArrayGPU.from_numpy(flags, valid, device),
            OArr(O.BOOL, O.pack_bits(flags), n, O.pack_bits(valid) if nulls else None))


def assert_same_bool(g, o, what):
    dev = g.gpu_device
    assert g.len == o.n, what
    assert np.array_equal(dev.retrive_data(g.data, O.words(o.n) * 4).view(np.uint32), o.data), what
    if o.valid is None:
        assert g.null_buffer is None, what
    else:
        assert np.array_equal(dev.retrive_data(g.null_buffer.bit_buffer, O.words(o.n) * 4).view(np.uint32), o.valid), \
            what + " (validity)"


def check(g, o, what):
    (assert_same_bool if o.dtype == O.BOOL else assert_same)(g, o, what)


def call(g, name, pipe, *args):
    """eager method, or the recording `_op` form on a pipeline"""
    if pipe is None:
        return getattr(g, name)(*args)
    return getattr(g, name + "_op")(*args, pipe)


def run_program(seed, device, fuse):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice(LENGTHS))
    pipe = ag.ArrowComputePipeline(device, fuse=True) if fuse else None
    pool = []  # (gpu array, oracle array)
    for dt in rng.choice(INTS + [O.F32, O.F32], size=3):
        pool.append(make(rng, int(dt), n, bool(rng.random() < 0.5), device))
    pool.append(make_bool(rng, n, bool(rng.random() < 0.5), device))
    trace = []
    pending = []

    def emit(g, o, what):
        trace.append(what)
        pending.append((g, o, " -> ".join(trace[-4:]) + f" [seed {seed} n {n}]"))
        pool.append((g, o))

    for _step in range(int(rng.integers(4, 9))):
        g, o = pool[int(rng.integers(len(pool)))]
        dt = o.dtype
        same = [(g2, o2) for g2, o2 in pool if o2.dtype == dt and o2.n == o.n]
        g2, o2 = same[int(rng.integers(len(same)))]
        if dt == O.BOOL:
            kind = rng.choice(["logic", "not", "cast", "merge", "take"])
            if kind == "logic":
                op = str(rng.choice(["bitwise_and", "bitwise_or", "bitwise_xor"]))
                emit(call(g, op, pipe, g2), oracle_binary(op, o, o2), f"bool.{op}")
            elif kind == "not":
                emit(call(g, "bitwise_not", pipe), oracle_unary("bitwise_not", o), "bool.not")
            elif kind == "cast":
                emit(call(g, "cast", pipe, ag.Float32ArrayGPU), oracle_cast(o, O.F32), "bool.cast f32")
            elif kind == "merge":
                gm, om = make_bool(rng, n, bool(rng.random() < 0.5), device)
                emit(call(g, "merge", pipe, g2, gm), oracle_merge(o, o2, om), "bool.merge")
            else:
                idx = rng.integers(0, n, n).astype(np.uint32)
                emit(call(g, "take", pipe, ag.UInt32ArrayGPU.from_numpy(idx, None, device)),
                     oracle_take(o, OArr(O.U32, idx, n)), "bool.take")
            continue
        name = NAMES[dt]
        kinds = ["binary", "binary", "scalar", "compare", "merge", "take", "filter"]
        kinds += ["cast"] if dt in CASTS else []
        kinds += ["not", "shift", "logic"] if dt != O.F32 else ["unary", "unary"]
        kind = rng.choice(kinds)
        if kind == "binary":
            op = str(rng.choice(["add", "sub", "mul", "min", "max"] + (["div"] if dt == O.F32 else [])))
            emit(call(g, op, pipe, g2), oracle_binary(op, o, o2), f"{name}.{op}")
        elif kind == "logic":
            op = str(rng.choice(["bitwise_and", "bitwise_or", "bitwise_xor"]))
            emit(call(g, op, pipe, g2), oracle_binary(op, o, o2), f"{name}.{op}")
        elif kind == "scalar":
            op = str(rng.choice(["add_scalar", "sub_scalar", "mul_scalar", "div_scalar", "rem_scalar"]))
            sv = rand_vals(rng, dt, 1, special=False)
            if rng.random() < 0.15:
                sv[0] = 0  # the divide-by-zero rules
            emit(call(g, op, pipe, ALL_CLS[dt].from_numpy(sv, None, device)), oracle_scalar(op, o, OArr(dt, sv, 1)),
                 f"{name}.{op}({sv[0]})")
        elif kind == "compare":
            op = str(rng.choice(["gt", "gteq", "lt", "lteq", "eq"]))
            emit(call(g, op, pipe, g2), oracle_binary(op, o, o2), f"{name}.{op}")
        elif kind == "not":
            emit(call(g, "bitwise_not", pipe), oracle_unary("bitwise_not", o), f"{name}.not")
        elif kind == "unary":
            op = str(rng.choice(["neg", "abs", "sqrt"]))
            emit(call(g, op, pipe), oracle_unary(op, o), f"{name}.{op}")
        elif kind == "shift":
            op = str(rng.choice(["bitwise_shl", "bitwise_shr"]))
            cnt = rng.integers(0, 40, n).astype(np.uint32)  # counts above the width too: `& 31` rule
            gc = ag.UInt32ArrayGPU.from_numpy(cnt, rng.random(n) < 0.9 if rng.random() < 0.3 else None, device)
            oc = OArr(O.U32, cnt, n, None if gc.null_buffer is None else
                      device.retrive_data(gc.null_buffer.bit_buffer, O.words(n) * 4).view(np.uint32).copy())
            emit(call(g, op, pipe, gc), oracle_binary(op, o, oc), f"{name}.{op}")
        elif kind == "cast":
            dst = int(rng.choice(CASTS[dt]))
            emit(call(g, "cast", pipe, ALL_CLS[dst]), oracle_cast(o, dst), f"{name}.cast {NAMES[dst]}")
        elif kind == "merge":
            gm, om = make_bool(rng, n, bool(rng.random() < 0.5), device)
            emit(call(g, "merge", pipe, g2, gm), oracle_merge(o, o2, om), f"{name}.merge")
        elif kind == "take":
            idx = rng.integers(0, n, n).astype(np.uint32)
            emit(call(g, "take", pipe, ag.UInt32ArrayGPU.from_numpy(idx, None, device)),
                 oracle_take(o, OArr(O.U32, idx, n)), f"{name}.take")
        else:  # filter changes the length: checked, not fed back
            gm, om = make_bool(rng, n, bool(rng.random() < 0.5), device)
            got = g.filter(gm)
            check(got, oracle_filter(o, om), f"{name}.filter [seed {seed} n {n}]")
    if pipe is not None:
        pipe.finish()
    for g, o, what in pending:
        check(g, o, what)


@pytest.mark.parametrize("seed", range(200))
def test_random_program_eager(seed, device):
    run_program(seed, device, fuse=False)


@pytest.mark.parametrize("seed", range(200))
def test_random_program_fusing_pipeline(seed, device):
    run_program(seed, device, fuse=True)
