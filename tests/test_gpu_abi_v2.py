"""GPU: the ABI-v2 pieces — bounds-checked put, allocator guards (double free, blocks used on
another handle's stream), filter output capacity, captured pipelines (CUDA graph replay with
programmatic dependent launches), in-place ops on a fusing pipeline, several device ordinals."""
import ctypes as C

import numpy as np
import pytest

import arrow_gpu_b200 as ag
from arrow_gpu_b200 import _ffi
from arrow_gpu_b200 import kernels as K
import oracle as O

pytestmark = pytest.mark.gpu


def test_put_out_of_range_indices_follow_robust_buffer_access(device):
    """routines/compute_shaders/32bit/put.wgsl relies on wgpu's robust buffer access: a source
    index past the end reads zero, a destination index past the end writes nothing — in
    particular not into the neighbouring block of the pool."""
    lib = _ffi.lib()
    for cls, npt in ((ag.Int32ArrayGPU, np.int32), (ag.UInt16ArrayGPU, np.uint16), (ag.Int8ArrayGPU, np.int8)):
        src_n, dst_n, m = 100, 64, 4096
        rng = np.random.default_rng(3)
        src = rng.integers(1, 100, src_n).astype(npt)
        # one allocation: [dst | guard]; the array only covers the first dst_n rows
        backing = ag.GpuDevice.create_gpu_buffer_with_data(device, np.full(dst_n + 4096, 77, dtype=npt))
        dst = cls(ag.ArrowGpuBuffer(device, backing.ptr, dst_n * np.dtype(npt).itemsize, owned=False), device, dst_n, None)
        si = rng.integers(0, src_n * 2, m).astype(np.uint32)         # half of them out of range
        di = rng.integers(dst_n, dst_n * 8, m).astype(np.uint32)     # out of range ...
        di[rng.permutation(m)[:40]] = rng.permutation(dst_n)[:40]    # ... except 40 unique in-range destinations
        cls.from_numpy(src, None, device).put(ag.UInt32ArrayGPU.from_numpy(si, None, device), dst,
                                              ag.UInt32ArrayGPU.from_numpy(di, None, device))
        got = device.retrive_data(backing).view(npt)
        want = np.full(dst_n + 4096, 77, dtype=npt)
        ok = di < dst_n
        want[di[ok]] = np.where(si[ok] < src_n, src[np.minimum(si[ok], src_n - 1)], 0)
        assert np.array_equal(got, want), cls.__name__
        # the oracle agrees on the in-range part
        want_o = O.put(cls.DTYPE, src, si[ok & (si < src_n)], np.full(dst_n, 77, dtype=npt), di[ok & (si < src_n)])
        m_in = ok & (si < src_n)
        assert np.array_equal(got[:dst_n][di[m_in]], want_o[di[m_in]])
    # bool put: bits past dst_len are left alone
    bits = ag.BooleanArrayGPU.from_numpy(np.ones(40, dtype=bool), None, device)
    dstb = ag.BooleanArrayGPU.from_numpy(np.zeros(40, dtype=bool), None, device)
    si = np.array([0, 1, 2, 999], dtype=np.uint32)
    di = np.array([3, 70, 5, 6], dtype=np.uint32)     # 70 >= 40: dropped;  source 999: reads 0
    bits.put(ag.UInt32ArrayGPU.from_numpy(si, None, device), dstb, ag.UInt32ArrayGPU.from_numpy(di, None, device))
    raw = device.retrive_data(dstb.data, 8).view(np.uint32)
    assert raw[0] == (1 << 3) | (1 << 5) and raw[1] == 0
    assert lib.agpu_abi_version() == 2


def test_double_free_is_rejected_and_blocks_are_not_shared(device):
    lib = _ffi.lib()
    p = C.c_void_p()
    _ffi.check(lib.agpu_alloc(device.handle, 1 << 20, C.byref(p)), "alloc")
    assert lib.agpu_free(device.handle, p) == 0
    assert lib.agpu_free(device.handle, p) == -4            # AGPU_EDOUBLEFREE, cache unchanged
    assert lib.agpu_free(device.handle, C.c_void_p(p.value + 256)) == -4   # never allocated here
    a, b = C.c_void_p(), C.c_void_p()
    _ffi.check(lib.agpu_alloc(device.handle, 1 << 20, C.byref(a)), "alloc")
    _ffi.check(lib.agpu_alloc(device.handle, 1 << 20, C.byref(b)), "alloc")
    assert a.value != b.value                                # a double free used to hand one block out twice
    # a block may be freed through ANOTHER handle of the same GPU: it goes back to its owner
    other = ag.GpuDevice(device.ordinal)
    assert lib.agpu_free(other.handle, a) == 0
    assert lib.agpu_free(device.handle, a) == -4
    other.sync()                                             # the block is parked until `other` got past the free
    c = C.c_void_p()
    _ffi.check(lib.agpu_alloc(device.handle, 1 << 20, C.byref(c)), "alloc")
    assert c.value == a.value                                # reused from the OWNER's cache
    for q in (b, c):
        assert lib.agpu_free(device.handle, q) == 0
    device.sync()
    other.sync()


def test_block_freed_on_upload_handle_waits_for_the_compute_handle(device):
    """bench.py's e2e pattern: columns are uploaded (allocated) on an upload handle and consumed on
    the compute handle.  Dropping the column while the compute stream is still far behind must not
    let the upload handle reuse the block for the next upload."""
    up = ag.GpuDevice(device.ordinal)
    n = 1 << 26                                       # 256 MiB per f32 column
    ones = device.pinned_empty(n, np.float32)
    ones[:] = 1.0
    twos = device.pinned_empty(n, np.float32)
    twos[:] = 2.0
    big = ag.Float32ArrayGPU.from_numpy(np.ones(1 << 28, dtype=np.float32), None, device)
    for trial in range(3):
        x = ag.Float32ArrayGPU.from_numpy(ones, None, up, wait=False)
        ready = up.record_event()
        # a deep queue on the compute stream: ~20 ms of work before it gets to x
        for _ in range(40):
            big = big.add(big)
        device.wait_event(ready)
        x.gpu_device = device                         # ops on this column run on the compute handle
        y = x.add(x)                                  # enqueued behind the queue above
        del x                                         # freed through the upload handle -> its cache
        z = ag.Float32ArrayGPU.from_numpy(twos, None, up, wait=False)   # same size: reuses the block
        up.sync()
        got = y.raw_values()
        assert got[0] == 2.0 and got[-1] == 2.0 and np.all(got == 2.0), f"trial {trial}: block reused while still read"
        assert np.all(z.raw_values() == 2.0)
        big = ag.Float32ArrayGPU.from_numpy(np.ones(1 << 28, dtype=np.float32), None, device)
    device.pinned_free(ones)
    device.pinned_free(twos)


def test_filter_scatter_respects_output_capacity(device):
    lib = _ffi.lib()
    n = 200_000
    rng = np.random.default_rng(9)
    vals = rng.integers(-2**31, 2**31, n, dtype=np.int64).astype(np.int32)
    valid = rng.random(n) < 0.9
    keep = rng.random(n) < 0.5
    a = ag.Int32ArrayGPU.from_numpy(vals, valid, device)
    m = ag.BooleanArrayGPU.from_numpy(keep, None, device)
    total = int(keep.sum())
    for cap in (total, total - 1, 4096 * 3 + 7, 1, n):
        plan = a.filter_count_op(m, None)
        guard = 1024
        out = device.create_gpu_buffer_with_data(np.full(min(cap, n) + guard, 0x5A5A5A5A, dtype=np.uint32))
        vwords = (cap + 31) // 32
        vout = device.create_gpu_buffer_with_data(np.full(vwords + guard, 0xFFFFFFFF, dtype=np.uint32))
        _ffi.check(lib.agpu_filter_scatter(device.handle, a.DTYPE, a.data.ptr, a.null_buffer.bit_buffer.ptr, m.data.ptr, None,
                                           n, plan.scratch.ptr, out.ptr, vout.ptr, cap), "filter_scatter")
        got = device.retrive_data(out).view(np.int32)
        k = min(cap, total)
        assert np.array_equal(got[:k], vals[keep][:k]), cap
        assert np.all(got[cap:].view(np.uint32) == 0x5A5A5A5A), f"cap={cap}: wrote past the capacity"
        gv = device.retrive_data(vout).view(np.uint32)
        assert np.all(gv[vwords:] == 0xFFFFFFFF), f"cap={cap}: validity written past the capacity"
        bits = O.unpack_bits(gv[:vwords].copy(), vwords * 32)
        assert np.array_equal(bits[:k], valid[keep][:k]), cap
        assert not bits[k:].any(), cap


def test_captured_pipeline_replays_with_one_submit(device):
    """ArrowComputePipeline(capture=True): record -> finish() submits the whole program once
    (compute_pipeline.rs:259-273); replay() submits it again.  Results identical to eager ops and
    to the oracle; a replay after the inputs changed in place sees the new data."""
    n = 1 << 20
    rng = np.random.default_rng(1)
    a_h = rng.uniform(-1000, 1000, n).astype(np.float32)
    b_h = rng.uniform(-1000, 1000, n).astype(np.float32)
    va, vb = rng.random(n) < 0.9, rng.random(n) < 0.9
    a = ag.Float32ArrayGPU.from_numpy(a_h, va, device)
    b = ag.Float32ArrayGPU.from_numpy(b_h, vb, device)
    l0 = device.launch_count()
    p = ag.ArrowComputePipeline(device, "cfg1", capture=True)
    s = a.add_op(b, p)
    t = s.mul_op(a, p)            # s is consumed inside the program
    g = t.gt_op(b, p)
    del t                         # a temporary freed inside the capture stays with the graph
    assert device.launch_count() - l0 == 3      # recorded, counted again per submit
    p.finish()
    assert p.graph.kernels == 3
    assert device.launch_count() - l0 == 3      # finish() = one submit of the three kernels

    def check(a_h, b_h):
        want_s = O.binary(O.ADD, O.F32, a_h, b_h)
        want_g = O.compare(O.GT, O.F32, O.binary(O.MUL, O.F32, want_s, a_h), b_h)
        assert np.array_equal(s.raw_values().view(np.uint32), want_s.view(np.uint32))
        assert np.array_equal(device.retrive_data(g.data, O.words(n) * 4).view(np.uint32), want_g)
        want_v = O.validity_and(O.pack_bits(va), O.pack_bits(vb), n)
        assert np.array_equal(device.retrive_data(g.null_buffer.bit_buffer, O.words(n) * 4).view(np.uint32), want_v)

    check(a_h, b_h)
    # unrelated eager work between replays must not disturb the program's buffers (or vice versa)
    junk = [ag.Float32ArrayGPU.from_numpy(b_h, None, device).add(b) for _ in range(4)]
    a2 = rng.uniform(-5, 5, n).astype(np.float32)
    _ffi.check(_ffi.lib().agpu_h2d(device.handle, a.data.ptr, a2.ctypes.data, a2.nbytes), "h2d")
    device.sync()
    for _ in range(3):
        p.replay()
    assert device.launch_count() - l0 == 3 + 4 + 9
    check(a2, b_h)
    for j in junk:
        assert np.array_equal(j.raw_values().view(np.uint32), O.binary(O.ADD, O.F32, b_h, b_h).view(np.uint32))
    # host synchronisation inside a capture is refused, the capture can be abandoned
    q = ag.ArrowComputePipeline(device, "bad", capture=True)
    assert _ffi.lib().agpu_sync(device.handle) == -2
    q.abort()
    assert np.array_equal(a.add(b).raw_values().view(np.uint32), O.binary(O.ADD, O.F32, a2, b_h).view(np.uint32))


def test_captured_fusing_pipeline_and_many_small_ops(device):
    """capture + fuse together, and a 40-op program replayed: the launch-bound regime the graph
    path exists for (1 Ki-row columns)"""
    n = 1024
    rng = np.random.default_rng(2)
    x_h = rng.integers(-100, 100, n).astype(np.int32)
    y_h = rng.integers(1, 50, n).astype(np.int32)
    x = ag.Int32ArrayGPU.from_numpy(x_h, None, device)
    y = ag.Int32ArrayGPU.from_numpy(y_h, None, device)
    p = ag.ArrowComputePipeline(device, "many", capture=True)
    acc = x
    for k in range(40):
        acc = acc.add_op(y, p) if k % 2 == 0 else acc.bitwise_xor_op(y, p)
    p.finish()
    want = x_h.copy()
    for k in range(40):
        want = (want + y_h).astype(np.int32) if k % 2 == 0 else want ^ y_h
    assert np.array_equal(acc.raw_values(), want)
    for _ in range(5):
        p.replay()
    assert np.array_equal(acc.raw_values(), want)
    pf = ag.ArrowComputePipeline(device, "fused", fuse=True, capture=True)
    r = x.add_op(y, pf).mul_op(y, pf).gt_op(x, pf)
    pf.finish()
    assert pf.graph.kernels == 1
    want_r = ((x_h + y_h) * y_h).astype(np.int32) > x_h
    assert np.array_equal(r.raw_values(), want_r)


def test_put_on_a_fusing_pipeline_keeps_the_recording_order(device):
    """y = dst.add_op(c, p); src.put_op(si, dst, di, p); p.finish() — the add was
    recorded before the put and must see dst as it was (the reference's encoder order)."""
    n = 5000
    rng = np.random.default_rng(4)
    d_h = rng.integers(-1000, 1000, n).astype(np.int32)
    c_h = rng.integers(-1000, 1000, n).astype(np.int32)
    s_h = rng.integers(5000, 9000, n).astype(np.int32)
    si = rng.integers(0, n, 700).astype(np.uint32)
    di = rng.permutation(n)[:700].astype(np.uint32)
    results = []
    for fuse in (False, True):
        dst = ag.Int32ArrayGPU.from_numpy(d_h, None, device)
        c = ag.Int32ArrayGPU.from_numpy(c_h, None, device)
        src = ag.Int32ArrayGPU.from_numpy(s_h, None, device)
        p = ag.ArrowComputePipeline(device, "order", fuse=fuse)
        y = dst.add_op(c, p)
        src.put_op(ag.UInt32ArrayGPU.from_numpy(si, None, device), dst, ag.UInt32ArrayGPU.from_numpy(di, None, device), p)
        z = dst.add_op(c, p)       # recorded after the put: sees the new dst
        p.finish()
        results.append((y.raw_values().copy(), z.raw_values().copy(), dst.raw_values().copy()))
    want_dst = d_h.copy()
    want_dst[di] = s_h[si]
    for y, z, d in results:
        assert np.array_equal(y, (d_h + c_h).astype(np.int32))
        assert np.array_equal(d, want_dst)
        assert np.array_equal(z, (want_dst + c_h).astype(np.int32))


def test_handles_of_several_ordinals_and_worker_threads(device):
    """every entry point selects the handle's device itself — a second ordinal in the
    same process and calls from a fresh thread (whose current device is 0) both work"""
    import threading
    n = 100_000
    x = np.arange(n, dtype=np.int32)
    devices = [ag.GpuDevice(o) for o in range(_ffi.device_count())]
    outs = {}

    def work(k, d):
        a = ag.Int32ArrayGPU.from_numpy(x, None, d)
        outs[k] = a.add(a).mul_scalar(ag.Int32ArrayGPU.from_slice([3], d)).raw_values()

    # interleaved on one thread ...
    for k, d in enumerate(devices):
        work(("main", k), d)
    # ... and each from its own thread, last ordinal first
    threads = [threading.Thread(target=work, args=(("thread", k), d)) for k, d in reversed(list(enumerate(devices)))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert len(outs) == 2 * len(devices)
    for v in outs.values():
        assert np.array_equal(v, x * 6)


def test_host_objects_dropped_in_the_middle_of_a_recording(device):
    """A pipeline variable rebound inside a loop drops the PREVIOUS graph after the next recording
    has begun; events and pinned buffers can go at any moment too (garbage collection).  None of
    that may invalidate the capture in progress (CUDA error 901 on the next launch)."""
    n = 50_000
    x_h = np.arange(n, dtype=np.int32)
    x = ag.Int32ArrayGPU.from_numpy(x_h, None, device)
    outs = []
    p = None
    for k in range(4):
        p = ag.ArrowComputePipeline(device, f"loop{k}", capture=True)   # rebinding drops graph k-1 here
        ev = ag.GpuEvent()
        pin = device.pinned_empty(1024, np.uint8)
        y = x.add_op(x, p)
        del ev                                                          # cudaEventDestroy while recording
        device.pinned_free(pin)
        z = y.bitwise_xor_op(x, p)
        p.finish()
        outs.append(z)
    for z in outs:
        assert np.array_equal(z.raw_values(), (x_h + x_h) ^ x_h)
    p.replay()
    assert np.array_equal(outs[-1].raw_values(), (x_h + x_h) ^ x_h)
