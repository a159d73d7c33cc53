"""CPU: the recording / auto-fusion logic of the host mirror (kernels.py) against a STUB of the C
ABI that only records calls.  No kernel runs and no result is computed here — the stub exists so
that the chain-building rules (what is recorded, where a chain ends, how many kernels a recorded
program becomes) are checked on every CPU run; the numerics of the same programs are checked on
the GPU (test_gpu_parity.py, test_gpu_fuzz.py)."""
import ctypes as C

import numpy as np
import pytest

import arrow_gpu_b200 as ag
from arrow_gpu_b200 import _ffi
from arrow_gpu_b200 import kernels as K

pytestmark = pytest.mark.timeout(120)

LAUNCHING = ("agpu_binary", "agpu_scalar", "agpu_unary", "agpu_compare", "agpu_shift", "agpu_cast", "agpu_fused_chain",
             "agpu_fused_chain_int", "agpu_fused_chain_pair", "agpu_fused_mul_add_gt", "agpu_bitmap_binary", "agpu_bitmap_not", "agpu_merge",
             "agpu_take")


class StubLib:
    """records (name, args); fills the out-parameters of the allocation / creation calls"""

    def __init__(self):
        self.calls, self._next = [], 0x10000

    def _ptr(self):
        self._next += 0x1000
        return self._next

    def __getattr__(self, name):
        def fn(*args):
            if name in ("agpu_device_create", "agpu_alloc", "agpu_event_create", "agpu_host_alloc"):
                args[-1]._obj.value = self._ptr()
            elif name == "agpu_launch_count":
                return sum(1 for n, _a in self.calls if n in LAUNCHING)
            if name == "agpu_h2d":                              # keep the uploaded bytes: (dst, bytes)
                self.calls.append((name, (args[1], C.string_at(args[2], args[3]))))
                return 0
            if name.startswith("agpu_fused_chain"):
                steps = [(st.kind, st.op, st.operand, st.scalar) for st in list(args[4])[: args[5]]]
                self.calls.append((name, steps))
            else:
                self.calls.append((name, args))
            return 0
        return fn

    def launched(self):
        return [(n, a) for n, a in self.calls if n in LAUNCHING]


@pytest.fixture
def stub(monkeypatch):
    lib = StubLib()
    monkeypatch.setattr(_ffi, "_lib", lib)
    dev = ag.GpuDevice(0)
    yield lib, dev
    dev.handle = None            # nothing to destroy


def f32(dev, n=64):
    return ag.Float32ArrayGPU.from_slice([1.0] * n, dev)


def test_plain_pipeline_launches_every_op(stub):
    lib, dev = stub
    a, b, s = f32(dev), f32(dev), f32(dev, 1)
    p = ag.ArrowComputePipeline(dev)
    K.gt_op_dyn(K.add_op_dyn(K.mul_scalar_op_dyn(a, s, p), b, p), b, p)
    p.finish()
    assert [n for n, _ in lib.launched()] == ["agpu_scalar", "agpu_binary", "agpu_compare"]


def test_fusing_pipeline_records_one_chain(stub):
    lib, dev = stub
    a, b, s = f32(dev), f32(dev), f32(dev, 1)
    p = ag.ArrowComputePipeline(dev, fuse=True)
    r = K.sqrt_op_dyn(K.add_op_dyn(K.mul_scalar_op_dyn(a, s, p), b, p), p)
    assert lib.launched() == []                               # nothing runs while recording
    p.finish()
    (name, steps), = lib.launched()
    assert name == "agpu_fused_chain"
    assert [(k, o) for k, o, _p, _s in steps] == [(_ffi.STEP_BINARY_DEVSCALAR, _ffi.MUL), (_ffi.STEP_BINARY_COLUMN, _ffi.ADD),
                                                  (_ffi.STEP_UNARY, _ffi.SQRT)]
    assert steps[0][2] == s.data.ptr and steps[1][2] == b.data.ptr
    assert r.len == a.len


def test_compare_ends_and_launches_the_chain_at_once(stub):
    lib, dev = stub
    a, b = f32(dev), f32(dev)
    p = ag.ArrowComputePipeline(dev, fuse=True)
    m = K.lteq_op_dyn(K.mul_op_dyn(a, b, p), b, p)
    (name, steps), = lib.launched()                           # before finish()
    assert name == "agpu_fused_chain" and steps[-1][:2] == (_ffi.STEP_COMPARE_COLUMN, _ffi.LTEQ)
    assert isinstance(m, ag.BooleanArrayGPU)
    p.finish()
    assert len(lib.launched()) == 1


def test_chain_limits_split_into_more_kernels(stub):
    lib, dev = stub
    a = f32(dev)
    p = ag.ArrowComputePipeline(dev, fuse=True)
    r = a
    for _ in range(_ffi.CHAIN_MAX_STEPS + 3):                 # 11 unary steps -> 8 + 3
        r = K.abs_op_dyn(r, p)
    p.finish()
    assert [len(steps) for _n, steps in lib.launched()] == [8, 3]
    lib.calls.clear()
    p = ag.ArrowComputePipeline(dev, fuse=True)
    r = a
    cols = [f32(dev) for _ in range(5)]
    for c in cols:                                            # 5 operand columns -> 3 + 2
        r = K.add_op_dyn(r, c, p)
    p.finish()
    assert [len(steps) for _n, steps in lib.launched()] == [3, 2]


def test_reading_a_recorded_array_launches_it_on_demand(stub):
    lib, dev = stub
    a, b = f32(dev), f32(dev)
    p = ag.ArrowComputePipeline(dev, fuse=True)
    r = K.add_op_dyn(a, b, p)
    assert lib.launched() == []
    assert r.data.ptr                                          # touching the buffer materialises the chain
    assert [n for n, _ in lib.launched()] == ["agpu_fused_chain"]
    p.finish()
    assert len(lib.launched()) == 1                            # not launched twice


def test_integer_chains_and_the_one_shift_rule(stub):
    lib, dev = stub
    a = ag.Int8ArrayGPU.from_slice([1] * 32, dev)
    b = ag.Int8ArrayGPU.from_slice([2] * 32, dev)
    s = ag.Int8ArrayGPU.from_slice([3], dev)
    cnt = ag.UInt32ArrayGPU.from_slice([1] * 32, dev)
    p = ag.ArrowComputePipeline(dev, fuse=True)
    r = K.bitwise_and_op_dyn(K.add_op_dyn(a, b, p), b, p)
    r = K.mul_scalar_op_dyn(r, s, p)
    r = K.bitwise_shl_op_dyn(r, cnt, p)
    r2 = K.bitwise_shr_op_dyn(r, cnt, p)                       # a second shift starts a new chain
    p.finish()
    launched = lib.launched()
    assert [n for n, _ in launched] == ["agpu_fused_chain_int", "agpu_fused_chain_int"]
    assert [k for k, *_ in launched[0][1]] == [_ffi.STEP_BINARY_COLUMN, _ffi.STEP_BINARY_COLUMN, _ffi.STEP_BINARY_DEVSCALAR,
                                               _ffi.STEP_SHIFT_COLUMN]
    assert [(k, o) for k, o, *_ in launched[1][1]] == [(_ffi.STEP_SHIFT_COLUMN, _ffi.SHR)]
    assert isinstance(r2, ag.Int8ArrayGPU)


def test_modes_do_not_mix_and_ineligible_ops_fall_back(stub):
    lib, dev = stub
    i8 = ag.Int8ArrayGPU.from_slice([1] * 32, dev)
    p = ag.ArrowComputePipeline(dev, fuse=True)
    t = K.bitwise_not_op_dyn(i8, p)                            # int-mode chain
    u = K.sin_op_dyn(t, p)                                     # f32-mode chain starting at the recorded int array
    p.finish()
    names = [n for n, _ in lib.launched()]
    assert names == ["agpu_fused_chain_int", "agpu_fused_chain"], names
    assert isinstance(u, ag.Float32ArrayGPU)
    lib.calls.clear()
    p = ag.ArrowComputePipeline(dev, fuse=True)
    other = ag.Int16ArrayGPU.from_slice([1] * 32, dev)
    with pytest.raises(ag.Panic):
        K.add_op_dyn(i8, other, p)                             # type mismatch is still the reference's panic
    m = ag.BooleanArrayGPU.from_slice([True] * 32, dev)
    K.merge_op_dyn(K.bitwise_not_op_dyn(i8, p), i8, m, p)      # merge is not fusable: its lazy operand is launched first
    assert [n for n, _ in lib.launched()] == ["agpu_fused_chain_int", "agpu_merge"]


def test_constructors_upload_the_reference_layout(stub):
    """primitive_array_gpu.rs:22-55, boolean_gpu.rs:23-50, null_bit_buffer.rs:10-62: dense
    little-endian values with T::default() in the null slots, LSB-first bitmaps padded to whole
    32-bit words with zero padding bits"""
    lib, dev = stub
    a = ag.Int16ArrayGPU.from_optional_slice([5, None, -2, None, 7], dev)
    uploads = {dst: data for n, (dst, data) in [(n, x) for n, x in lib.calls if n == "agpu_h2d"]}
    assert np.frombuffer(uploads[a.data.ptr], dtype="<i2").tolist() == [5, 0, -2, 0, 7]
    assert uploads[a.null_buffer.bit_buffer.ptr] == bytes([0b10101, 0, 0, 0])
    lib.calls.clear()
    flags = [True, False, None, True] + [True] * 30                      # 34 bits -> two words
    b = ag.BooleanArrayGPU.from_optional_slice(flags, dev)
    uploads = {dst: data for n, (dst, data) in [(n, x) for n, x in lib.calls if n == "agpu_h2d"]}
    data, valid = uploads[b.data.ptr], uploads[b.null_buffer.bit_buffer.ptr]
    assert len(data) == 8 and len(valid) == 8
    assert data[0] == 0b11111001 and data[4] == 0b11 and data[5:] == bytes(3)
    assert valid[0] == 0b11111011 and valid[4] == 0b11
    assert b.len == 34 and a.len == 5
    f = ag.Float32ArrayGPU.from_slice([1.0, -0.0], dev)
    assert f.null_buffer is None                                         # no nulls -> no bitmap (null_bit_buffer.rs:99-111)


def test_put_on_a_fusing_pipeline_launches_the_recorded_chains_first(stub):
    """put_op mutates `dst` at once; a chain recorded earlier on the same pipeline that
    reads `dst` must be launched BEFORE the put (the reference's encoder order), and ops recorded
    after it come after."""
    lib, dev = stub
    i32 = lambda n=64: ag.Int32ArrayGPU.from_slice([1] * n, dev)   # noqa: E731
    dst, c, src = i32(), i32(), i32()
    idx = ag.UInt32ArrayGPU.from_slice(list(range(8)), dev)
    p = ag.ArrowComputePipeline(dev, fuse=True)
    y = dst.add_op(c, p)                       # recorded, not launched
    assert lib.launched() == []
    src.put_op(idx, dst, idx, p)               # in place: flushes the recorded chain first
    order = [n for n, _a in lib.calls if n in LAUNCHING or n == "agpu_put"]
    assert order == ["agpu_fused_chain_int", "agpu_put"], order
    z = dst.add_op(c, p)                       # recorded after the put
    p.finish()
    order = [n for n, _a in lib.calls if n in LAUNCHING or n == "agpu_put"]
    assert order == ["agpu_fused_chain_int", "agpu_put", "agpu_fused_chain_int"], order
    put_args = [a for n, a in lib.calls if n == "agpu_put"][0]
    assert put_args[3] == src.len and put_args[6] == dst.len      # ABI v2: lengths for the bounds checks
    assert y.len == z.len == dst.len


def test_cross_handle_buffers_are_recorded_as_used(stub):
    """a column allocated through an upload handle and consumed on the compute handle: every op
    tells the allocator (agpu_buffer_record_use) before it launches"""
    lib, dev = stub
    up = ag.GpuDevice(0)
    a = ag.Int32ArrayGPU.from_slice([1] * 64, up)
    b = ag.Int32ArrayGPU.from_slice([2] * 64, dev)
    a.gpu_device = dev
    a.add(b)
    used = [args for n, args in lib.calls if n == "agpu_buffer_record_use"]
    assert len(used) == 1 and used[0][1] == a.data.ptr and used[0][0] == dev.handle
    names = [n for n, _a in lib.calls]
    assert names.index("agpu_buffer_record_use") < names.index("agpu_binary")
    up.handle = None


def test_captured_pipeline_brackets_the_ops_with_graph_begin_and_end(stub):
    lib, dev = stub
    a, b = f32(dev), f32(dev)
    p = ag.ArrowComputePipeline(dev, "prog", capture=True)
    a.add_op(b, p)
    a.gt_op(b, p)
    p.finish()
    p.replay()
    names = [n for n, _a in lib.calls if n.startswith("agpu_graph") or n in LAUNCHING]
    assert names == ["agpu_graph_begin", "agpu_binary", "agpu_compare", "agpu_graph_end", "agpu_graph_kernel_count",
                     "agpu_graph_launch", "agpu_graph_launch"], names


# ---- value chain + predicate chain over the same source: one agpu_fused_chain_pair launch ----
def f32v(dev, n=64):
    return ag.Float32ArrayGPU.from_numpy(np.ones(n, np.float32), np.arange(n) % 3 != 0, dev)


def test_add_then_gt_on_the_same_columns_is_one_kernel(stub):
    """BASELINE.json configs[0]: s = a + b; g = a > b"""
    lib, dev = stub
    for make in (f32, f32v):
        a, b = make(dev), make(dev)
        lib.calls.clear()
        p = ag.ArrowComputePipeline(dev, fuse=True)
        s = K.add_op_dyn(a, b, p)
        assert lib.launched() == []
        g = K.gt_op_dyn(a, b, p)
        (name, steps), = lib.launched()
        assert name == "agpu_fused_chain_pair"
        assert [(k, o) for k, o, _p, _s in steps] == [(_ffi.STEP_BINARY_COLUMN, _ffi.ADD), (_ffi.STEP_STORE, 0), (_ffi.STEP_RESET, 0),
                                                      (_ffi.STEP_COMPARE_COLUMN, _ffi.GT)]
        assert steps[0][2] == b.data.ptr and steps[3][2] == b.data.ptr
        assert s._lazy is None and s._data is not None and isinstance(g, ag.BooleanArrayGPU)
        if make is f32v:
            assert s.null_buffer.bit_buffer is g.null_buffer.bit_buffer        # one bitmap, written once
        else:
            assert s.null_buffer is None and g.null_buffer is None
        p.finish()
        assert len(lib.launched()) == 1                                         # nothing left to launch


def test_pairing_needs_the_same_validity_inputs_and_arithmetic_steps(stub):
    lib, dev = stub
    a, b, c = f32(dev), f32(dev), f32v(dev)
    # g depends on c's bitmap, s does not: two kernels
    p = ag.ArrowComputePipeline(dev, fuse=True)
    s = K.add_op_dyn(a, b, p)
    K.gt_op_dyn(a, c, p)
    p.finish()
    assert sorted(n for n, _ in lib.launched()) == ["agpu_fused_chain", "agpu_fused_chain"]
    # a transcendental in the value chain: the pair kernel is the arithmetic-only interpreter
    lib.calls.clear()
    p = ag.ArrowComputePipeline(dev, fuse=True)
    kept = K.sin_op_dyn(a, p)            # (an unreferenced recorded result is never launched)
    K.gt_op_dyn(a, b, p)
    p.finish()
    assert sorted(n for n, _ in lib.launched()) == ["agpu_fused_chain", "agpu_fused_chain"]
    # the predicate first: it is launched at once, there is nothing to pair the later value chain with
    lib.calls.clear()
    p = ag.ArrowComputePipeline(dev, fuse=True)
    K.gt_op_dyn(a, b, p)
    kept2 = K.add_op_dyn(a, b, p)
    p.finish()
    assert sorted(n for n, _ in lib.launched()) == ["agpu_fused_chain", "agpu_fused_chain"]
    assert s.len == a.len == kept.len == kept2.len


def test_absorbed_and_foreign_source_chains_are_not_paired(stub):
    lib, dev = stub
    a, b, c = f32(dev), f32(dev), f32(dev)
    p = ag.ArrowComputePipeline(dev, fuse=True)
    s = K.add_op_dyn(a, b, p)
    t = K.mul_op_dyn(s, c, p)            # absorbs s: the pending chain is [add b, mul c] over a
    g = K.lt_op_dyn(a, c, p)             # pairs with t's chain (same source a)
    (name, steps), = lib.launched()
    assert name == "agpu_fused_chain_pair" and len(steps) == 5 and t._lazy is None
    assert s._lazy is not None and s._lazy.consumed          # s itself stays unlaunched unless somebody reads it
    lib.calls.clear()
    K.eq_op_dyn(b, c, p)                 # source b: no pending chain starts there
    assert [n for n, _ in lib.launched()] == ["agpu_fused_chain"]
    p.finish()
    assert isinstance(g, ag.BooleanArrayGPU)


def test_pair_eligibility_rules(stub):
    _lib, dev = stub
    a, b, c, d, e = (f32(dev) for _ in range(5))
    assert K.pair_eligible(a, [("add", b)], [("gt", b)])
    assert K.pair_eligible(a, [("mul", b), ("add", c)], [("sub", d), ("lteq", b)])          # 3 distinct columns
    assert not K.pair_eligible(a, [("mul", b), ("add", c)], [("sub", d), ("lteq", e)])      # 4
    assert not K.pair_eligible(a, [("add", b)], [("add", b)])                                # no closing compare
    assert not K.pair_eligible(a, [("gt", b)], [("gt", b)])                                  # compare in the value chain
    assert not K.pair_eligible(a, [("power", b)], [("gt", b)])
    assert not K.pair_eligible(a, [("abs",)] * 4, [("neg",)] * 2 + [("eq", 1.0)])            # 4 + 3 + 2 > 8 steps
    assert K.pair_eligible(a, [("abs",)] * 3, [("neg",)] * 2 + [("eq", 1.0)])
    i = ag.Int8ArrayGPU.from_slice([1] * 64, dev)
    assert not K.pair_eligible(i, [("add", 1.0)], [("gt", 0.0)])                             # f32 sources only
    with pytest.raises(ag.Panic):
        K.fused_chain_pair(a, [("power", b)], [("gt", b)])


def test_a_predicate_that_reads_the_pending_value_is_not_paired_with_it(stub):
    lib, dev = stub
    a, b = f32(dev), f32(dev)
    p = ag.ArrowComputePipeline(dev, fuse=True)
    s = K.add_op_dyn(a, b, p)
    g = K.gt_op_dyn(a, s, p)             # needs s: s is launched first, then the compare
    p.finish()
    assert [n for n, _ in lib.launched()] == ["agpu_fused_chain", "agpu_fused_chain"]
    assert s._lazy is None and isinstance(g, ag.BooleanArrayGPU)
    # ... also when the dependency is indirect (an operand chain that has s as a column)
    lib.calls.clear()
    p = ag.ArrowComputePipeline(dev, fuse=True)
    s = K.add_op_dyn(a, b, p)
    t = K.mul_op_dyn(b, s, p)            # chain over b with s as operand column
    g = K.gt_op_dyn(a, t, p)
    p.finish()
    assert "agpu_fused_chain_pair" not in [n for n, _ in lib.launched()]
    assert len(lib.launched()) == 3
