"""GPU: zero-copy exchange of device-resident columns through the Arrow C Device Data Interface
(arrow_gpu_b200/c_device.py): round trip between two handles with the ordering carried by
`sync_event`, a foreign producer (a torch CUDA tensor wrapped in an ArrowDeviceArray), bitmaps whose
padding bits are garbage, and the producer's `release` running only after the consumer is done."""
import ctypes as C
import gc

import numpy as np
import pytest

import arrow_gpu_b200 as ag
from arrow_gpu_b200 import c_device as cd
import oracle as O
from helpers import OArr, oracle_binary, oracle_filter

pytestmark = pytest.mark.gpu


def test_round_trip_between_handles_is_zero_copy_and_ordered(device):
    """the producer's kernel is still running when the column is handed over: the consumer's
    stream must wait for sync_event, not for the host"""
    producer = ag.GpuDevice(device.ordinal)
    n = (1 << 24) + 77
    rng = np.random.default_rng(5)
    x = rng.integers(-2**31, 2**31, n).astype(np.int32)
    y = rng.integers(-2**31, 2**31, n).astype(np.int32)
    va, vb = rng.random(n) < 0.9, rng.random(n) < 0.9
    a = ag.Int32ArrayGPU.from_numpy(x, va, producer)
    b = ag.Int32ArrayGPU.from_numpy(y, vb, producer)
    for _ in range(3):
        s = a.add(b)                                    # enqueued on the producer's stream, not waited for
        got = ag.from_arrow_device(s, device)           # consumer handle: another stream
        assert got.data.ptr == s.data.ptr and got.null_buffer.bit_buffer.ptr == s.null_buffer.bit_buffer.ptr
        doubled = got.add(got)                          # runs on `device`'s stream
        want = oracle_binary("add", OArr(O.I32, x, n, O.pack_bits(va)), OArr(O.I32, y, n, O.pack_bits(vb)))
        want2 = O.binary(O.ADD, O.I32, want.data, want.data)
        assert np.array_equal(doubled.raw_values(), want2)
        assert np.array_equal(device.retrive_data(doubled.null_buffer.bit_buffer).view(np.uint32)[: O.words(n)], want.valid)
        del s                                           # the producer's handle goes first ...
        assert np.array_equal(got.raw_values(), want.data)   # ... the exported buffers stay valid
        del got, doubled
    gc.collect()
    producer.sync()
    producer.destroy()


def test_release_runs_once_after_the_consumer(device):
    a = ag.Float32ArrayGPU.from_numpy(np.arange(4096, dtype=np.float32), np.arange(4096) % 5 != 0, device)
    before = len(cd._exports)
    got = ag.from_arrow_device(a, device)
    gc.collect()
    assert len(cd._exports) == before + 1
    r = got.sqrt()
    del got
    gc.collect()
    assert len(cd._exports) == before
    assert np.array_equal(r.raw_values(), np.sqrt(np.arange(4096, dtype=np.float32)))
    assert np.array_equal(a.raw_values(), np.arange(4096, dtype=np.float32))        # the producer still owns its column


def test_foreign_producer_torch_tensor(device):
    """another library on the same GPU: a torch tensor produced on torch's stream, handed over as
    an ArrowDeviceArray whose sync_event is a torch CUDA event"""
    torch = pytest.importorskip("torch")
    n = (1 << 22) + 5
    with torch.cuda.device(device.ordinal):
        t = torch.arange(n, device="cuda", dtype=torch.float32) * 0.5 + 1.0
        ev = torch.cuda.Event()
        ev.record()
        event_slot = C.c_void_p(ev.cuda_event)                          # a cudaEvent_t; sync_event points to it
        schema_ptr = cd._new_schema(b"f", False)
        array_ptr = cd._new_device_array(n, 0, [None, t.data_ptr()], (t, ev, event_slot), cd.ARROW_DEVICE_CUDA,
                                         device.ordinal, C.addressof(event_slot))
        caps = (cd._capsule_new(schema_ptr, cd._SCHEMA_NAME, C.cast(cd._destroy_schema_capsule, C.c_void_p)),
                cd._capsule_new(array_ptr, cd._DEVICE_ARRAY_NAME, C.cast(cd._destroy_device_array_capsule, C.c_void_p)))
        col = ag.from_arrow_device(caps, device)
        assert isinstance(col, ag.Float32ArrayGPU) and col.data.ptr == t.data_ptr() and col.null_buffer is None
        got = col.mul(col).raw_values()
        host = np.arange(n, dtype=np.float32) * np.float32(0.5) + np.float32(1.0)
        assert np.array_equal(got, host * host)
        del col, caps
        gc.collect()


@pytest.mark.parametrize("n", [1, 31, 33, 1000, 4097])
def test_imported_bitmaps_may_carry_garbage_padding(device, n):
    """Arrow leaves the bits past `length` unspecified; a zero-copy import cannot clear them, so
    every consumer of a bitmap has to ignore them"""
    rng = np.random.default_rng(n)
    flags, valid = rng.random(n) < 0.5, rng.random(n) < 0.8
    words = O.words(n)

    def dirty(bits):
        raw = O.pack_bits(bits).copy()
        if n % 32:
            raw[words - 1] |= np.uint32(0xFFFFFFFF) << np.uint32(n % 32)
        return raw

    data_buf = device.create_gpu_buffer_with_data(dirty(flags))
    valid_buf = device.create_gpu_buffer_with_data(dirty(valid))
    schema_ptr = cd._new_schema(b"b", True)
    array_ptr = cd._new_device_array(n, -1, [valid_buf.ptr, data_buf.ptr], (data_buf, valid_buf), cd.ARROW_DEVICE_CUDA,
                                     device.ordinal, None)
    caps = (cd._capsule_new(schema_ptr, cd._SCHEMA_NAME, C.cast(cd._destroy_schema_capsule, C.c_void_p)),
            cd._capsule_new(array_ptr, cd._DEVICE_ARRAY_NAME, C.cast(cd._destroy_device_array_capsule, C.c_void_p)))
    m = ag.from_arrow_device(caps, device)
    assert m.data.ptr == data_buf.ptr
    assert np.array_equal(m.raw_values(), flags) and np.array_equal(m.null_buffer.flags(), valid)
    assert m.any() == bool(flags.any()) and m.all() == bool(flags.all())
    inv = m.bitwise_not()
    assert np.array_equal(device.retrive_data(inv.data).view(np.uint32)[:words], O.pack_bits(~flags))
    both = m.bitwise_and(m)
    assert np.array_equal(device.retrive_data(both.data).view(np.uint32)[:words], O.pack_bits(flags))
    x = rng.integers(-2**31, 2**31, n).astype(np.int32)
    col = ag.Int32ArrayGPU.from_numpy(x, None, device)
    got, want = col.filter(m), oracle_filter(OArr(O.I32, x, n), OArr(O.BOOL, O.pack_bits(flags), n, O.pack_bits(valid)))
    assert got.len == want.n and np.array_equal(got.raw_values(), want.data[: want.n])
    twice = (x + x).astype(np.int32)
    merged = col.merge(col.add(col), m)
    assert np.array_equal(merged.raw_values(), O.merge(O.I32, x, twice, O.pack_bits(flags), n))


def test_cpp_mirror_exchanges_device_columns_zero_copy(device):
    """Python exporter -> C++ importer (arrow_gpu_b200/cpp/c_data_interface.hpp) -> device compute
    -> C++ exporter -> Python importer: two independent implementations of the structures, no
    host copy anywhere, ordering carried by the two sync events"""
    import os
    import subprocess
    cpp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "arrow_gpu_b200", "cpp")
    so = os.path.join(cpp, "libagpu_cdata.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", cpp], check=True)
    lib = C.CDLL(so)
    lib.agpu_cdata_device_apply.argtypes = [C.c_char_p] + [C.c_void_p] * 4
    lib.agpu_cdata_device_apply.restype = C.c_int

    def apply(op, column):
        schema_capsule, array_capsule = column.__arrow_c_device_array__()
        sp = cd._capsule_ptr(schema_capsule, cd._SCHEMA_NAME)
        ap = cd._capsule_ptr(array_capsule, cd._DEVICE_ARRAY_NAME)
        out_schema, out_array = cd._libc.malloc(C.sizeof(cd.ArrowSchema)), cd._libc.malloc(C.sizeof(cd.ArrowDeviceArray))
        C.memset(out_schema, 0, C.sizeof(cd.ArrowSchema)), C.memset(out_array, 0, C.sizeof(cd.ArrowDeviceArray))
        rc = lib.agpu_cdata_device_apply(op, sp, ap, out_schema, out_array)
        assert rc == 0, (op, rc)
        caps = (cd._capsule_new(out_schema, cd._SCHEMA_NAME, C.cast(cd._destroy_schema_capsule, C.c_void_p)),
                cd._capsule_new(out_array, cd._DEVICE_ARRAY_NAME, C.cast(cd._destroy_device_array_capsule, C.c_void_p)))
        return ag.from_arrow_device(caps, device)

    n = (1 << 22) + 13
    rng = np.random.default_rng(11)
    x = rng.integers(-2**31, 2**31, n).astype(np.int32)
    valid = rng.random(n) < 0.85
    before = len(cd._exports)
    a = ag.Int32ArrayGPU.from_numpy(x, valid, device)
    same = apply(b"identity", a)
    assert same.data.ptr == a.data.ptr and same.null_buffer.bit_buffer.ptr == a.null_buffer.bit_buffer.ptr
    assert np.array_equal(same.raw_values(), x)
    twice = apply(b"add", a)
    assert isinstance(twice, ag.Int32ArrayGPU) and twice.data.ptr != a.data.ptr
    assert np.array_equal(twice.raw_values(), O.binary(O.ADD, O.I32, x, x))
    assert np.array_equal(twice.null_buffer.flags(), valid)
    pred = apply(b"gt", a)
    assert isinstance(pred, ag.BooleanArrayGPU)
    want = O.compare(O.GT, O.I32, O.binary(O.ADD, O.I32, x, x), x)
    assert np.array_equal(device.retrive_data(pred.data).view(np.uint32)[: O.words(n)], want)
    for f32 in (np.float32(1.5), ):
        col = ag.Float32ArrayGPU.from_numpy(np.full(1000, f32), None, device)
        out = apply(b"add", col)
        assert out.null_buffer is None and np.array_equal(out.raw_values(), np.full(1000, f32 * 2))
    del same, twice, pred, out, col
    gc.collect()
    assert len(cd._exports) == before            # every structure we exported was released by the C++ side
    assert np.array_equal(a.raw_values(), x)
