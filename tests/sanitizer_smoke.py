"""Small end-to-end pass over every kernel family for `compute-sanitizer` (memcheck / racecheck /
initcheck / synccheck).  Not a pytest file: run as
    compute-sanitizer --tool racecheck python tests/sanitizer_smoke.py
Sizes are ragged and small so the instrumented run stays short; every result is still checked
against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import arrow_gpu_b200 as ag  # noqa: E402
from arrow_gpu_b200 import kernels as K  # noqa: E402
import oracle as O  # noqa: E402
from helpers import OArr, oracle_binary, oracle_filter, oracle_merge, oracle_take  # noqa: E402


def main():
    dev = ag.GPU_DEVICE()
    rng = np.random.default_rng(0)
    for n in (1, 33, 4097, 20011):
        x8 = rng.integers(-128, 128, n, dtype=np.int8)
        y8 = rng.integers(-128, 128, n, dtype=np.int8)
        va, vb = rng.random(n) < 0.8, rng.random(n) < 0.8
        a, b = ag.Int8ArrayGPU.from_numpy(x8, va, dev), ag.Int8ArrayGPU.from_numpy(y8, vb, dev)
        oa, ob = OArr(O.I8, x8, n, O.pack_bits(va)), OArr(O.I8, y8, n, O.pack_bits(vb))
        for op in ("add", "mul", "bitwise_xor", "min", "gt", "eq"):
            got, want = getattr(a, op)(b), oracle_binary(op, oa, ob)
            assert np.array_equal(got.raw_values(), want.raw_values()), (op, n)
        cnt = rng.integers(0, 8, n).astype(np.uint32)
        assert np.array_equal(a.bitwise_shr(ag.UInt32ArrayGPU.from_numpy(cnt, None, dev)).raw_values(), O.shift(O.SHR, O.I8, x8, cnt))
        assert np.array_equal(a.cast(ag.Float32ArrayGPU).raw_values(), O.cast(O.I8, O.F32, x8))
        assert np.array_equal(a.cast(ag.Int16ArrayGPU).raw_values(), O.cast(O.I8, O.I16, x8))
        f = rng.uniform(-50, 50, n).astype(np.float32)
        fa = ag.Float32ArrayGPU.from_numpy(f, va, dev)
        fa.sin(), fa.exp(), fa.sqrt(), a.sinh()
        assert np.array_equal(fa.sum().raw_values().view(np.uint32), np.array([O.sum(O.F32, f)], np.float32).view(np.uint32))
        K.fused_mul_add_gt(fa, fa, fa, fa)
        flags = rng.random(n) < 0.4
        m = ag.BooleanArrayGPU.from_numpy(flags, vb, dev)
        om = OArr(O.BOOL, O.pack_bits(flags), n, O.pack_bits(vb))
        i32 = rng.integers(-2**31, 2**31, n).astype(np.int32)
        ia, oia = ag.Int32ArrayGPU.from_numpy(i32, va, dev), OArr(O.I32, i32, n, O.pack_bits(va))
        assert np.array_equal(ia.merge(ia, m).raw_values(), oracle_merge(oia, oia, om).raw_values())
        idx = rng.integers(0, n, n).astype(np.uint32)
        got, want = ia.take(ag.UInt32ArrayGPU.from_numpy(idx, None, dev)), oracle_take(oia, OArr(O.U32, idx, n))
        assert np.array_equal(got.raw_values(), want.raw_values())
        assert np.array_equal(got.null_buffer.flags(), O.unpack_bits(want.valid, n))
        for arr, oarr in ((ia, oia), (a, oa)):
            got, want = arr.filter(m), oracle_filter(oarr, om)
            assert np.array_equal(got.raw_values(), want.raw_values())
            assert np.array_equal(got.null_buffer.flags(), O.unpack_bits(want.valid, want.n))
        dst = ag.Int32ArrayGPU.from_numpy(np.zeros(n, np.int32), None, dev)
        ag.Int32ArrayGPU.from_numpy(i32, None, dev).put(ag.UInt32ArrayGPU.from_numpy(idx, None, dev), dst,
                                                        ag.UInt32ArrayGPU.from_numpy(rng.permutation(n).astype(np.uint32), None, dev))
        m.any(), m.all(), m.bitwise_not(), m.take(ag.UInt32ArrayGPU.from_numpy(idx, None, dev))
        # fused chains: f32 interpreter, integer packed-word interpreter (value + predicate), u16 filter
        K.fused_chain(fa, [("abs",), ("sqrt",), ("mul", fa), ("lteq", fa)])
        # two results in one pass: the dedicated binop | compare kernel and the interpreter form
        fb = ag.Float32ArrayGPU.from_numpy(f[::-1].copy(), va, dev)
        pv, pp = K.fused_chain_pair(fa, [("add", fb)], [("gt", fb)])
        assert np.array_equal(pv.raw_values().view(np.uint32), fa.add(fb).raw_values().view(np.uint32))
        assert np.array_equal(pp.raw_values(), fa.gt(fb).raw_values())
        pv, pp = K.fused_chain_pair(fa, [("abs",), ("min", fb)], [("mul", 2.0), ("lteq", fb)])
        assert np.array_equal(pp.raw_values(), fa.mul_scalar(ag.Float32ArrayGPU.from_slice([2.0], dev)).lteq(fb).raw_values())
        s8 = ag.Int8ArrayGPU.from_slice([3], dev)
        got = K.fused_chain_int(a, [("add", b), ("bitwise_and", b), ("mul", K.DeviceScalar(s8))])
        assert np.array_equal(got.raw_values(), a.add(b).bitwise_and(b).mul_scalar(s8).raw_values())
        got = K.fused_chain_int(ia, [("bitwise_not",), ("div", ia), ("gt", ia)])
        assert np.array_equal(got.raw_values(), ia.bitwise_not().div(ia).gt(ia).raw_values())
        u16 = rng.integers(0, 65536, n).astype(np.uint16)
        assert np.array_equal(ag.UInt16ArrayGPU.from_numpy(u16, None, dev).filter(m).raw_values(), u16[flags & vb])
        # round 2: host-free filter (scatter launched before the count is read, output sized for the
        # shard), a captured pipeline replayed twice, put with out-of-range indices
        from arrow_gpu_b200 import sharded
        out, off, tot = sharded.sharded_filter_async(ia, m).result()
        assert off == 0 and tot == out.len and np.array_equal(out.raw_values(), i32[flags & vb])
        assert np.array_equal(out.null_buffer.flags(), va[flags & vb])
        p = ag.ArrowComputePipeline(dev, "smoke", capture=True)
        s1 = ia.add_op(ia, p)
        g1 = s1.gt_op(ia, p)
        p.finish()
        p.replay()
        p.replay()
        assert np.array_equal(g1.raw_values(), (i32 + i32).astype(np.int32) > i32)
        wild = idx.copy()
        wild[::3] += np.uint32(n)          # every third index out of range: reads zero / writes nothing
        ag.Int32ArrayGPU.from_numpy(i32, None, dev).put(ag.UInt32ArrayGPU.from_numpy(wild, None, dev), dst,
                                                        ag.UInt32ArrayGPU.from_numpy(wild, None, dev))
    dev.sync()
    print("sanitizer smoke ok,", dev.launch_count(), "launches")


if __name__ == "__main__":
    main()
