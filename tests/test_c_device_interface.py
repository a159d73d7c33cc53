"""CPU: the Arrow C Device Data Interface structures and ownership rules of arrow_gpu_b200/c_device.py.
The struct layouts are checked against pyarrow's own importer / exporter (device type CPU — this
pyarrow has no CUDA support); the zero-copy path and its release / move rules run against the
call-recording stub of the C ABI (no kernel runs; GPU behaviour: tests/test_gpu_c_device.py)."""
import ctypes as C
import gc

import numpy as np
import pyarrow as pa
import pytest

import arrow_gpu_b200 as ag
from arrow_gpu_b200 import _ffi, c_device as cd
from test_host_fusion_logic import StubLib

pytestmark = pytest.mark.timeout(120)


@pytest.fixture
def stub(monkeypatch):
    lib = StubLib()
    monkeypatch.setattr(_ffi, "_lib", lib)
    dev = ag.GpuDevice(0)
    yield lib, dev
    gc.collect()
    dev.handle = None


def test_struct_sizes_match_the_specification():
    assert C.sizeof(cd.ArrowSchema) == 72 and C.sizeof(cd.ArrowArray) == 80
    assert C.sizeof(cd.ArrowDeviceArray) == 80 + 8 + 8 + 8 + 24      # array, device_id, type (+pad), sync_event, reserved
    assert cd.ArrowDeviceArray.device_id.offset == 80 and cd.ArrowDeviceArray.sync_event.offset == 96


@pytest.mark.parametrize("ty,np_ty", [(pa.int8(), np.int8), (pa.uint16(), np.uint16), (pa.int32(), np.int32),
                                      (pa.float32(), np.float32), (pa.date32(), np.int32)])
def test_pyarrow_imports_what_the_exporter_builds(ty, np_ty):
    """an ArrowDeviceArray built by _new_device_array over HOST memory (device type CPU) is accepted
    by pyarrow's C-device importer and carries the same values and nulls"""
    rng = np.random.default_rng(3)
    n = 77
    values = rng.integers(-100, 100, n).astype(np_ty)
    valid = rng.random(n) < 0.7
    bitmap = ag.array.pack_bits(valid)
    at = {v: k for k, v in ag.interop._types().items()}
    fmt = cd._FORMATS[ag.interop._types()[ty]]
    schema_ptr = cd._new_schema(fmt, True)
    array_ptr = cd._new_device_array(n, int((~valid).sum()), [bitmap.ctypes.data, values.ctypes.data], (bitmap, values),
                                     cd.ARROW_DEVICE_CPU, -1, None)
    got = pa.Array._import_from_c_device(array_ptr, schema_ptr)   # moves both structures
    assert got.type == ty and len(got) == n
    want = pa.array(values, mask=~valid).cast(ty) if ty != pa.date32() else pa.array(values, mask=~valid).cast(pa.date32())
    assert got.equals(want)
    keys_before = len(cd._exports)
    del got
    gc.collect()
    assert len(cd._exports) < keys_before                          # pyarrow called our release callbacks
    cd._libc.free(schema_ptr), cd._libc.free(array_ptr)


def test_parses_what_pyarrow_exports():
    arr = pa.array([1, None, 3, 4, None], type=pa.int16())
    schema_capsule, array_capsule = arr.__arrow_c_device_array__()
    fmt, dptr = cd.parse_device_capsules(schema_capsule, array_capsule)
    d = dptr.contents
    assert fmt == b"s" and d.device_type == cd.ARROW_DEVICE_CPU
    assert d.array.length == 5 and d.array.null_count == 2 and d.array.n_buffers == 2
    assert d.array.buffers[1] == arr.buffers()[1].address and d.array.buffers[0] == arr.buffers()[0].address


def test_host_column_is_copied_in(stub):
    lib, dev = stub
    arr = pa.array([5, None, 7, 8, None, 10, 11], type=pa.uint8()).slice(1, 5)      # offset 1
    got = ag.from_arrow_device(arr, dev)
    assert isinstance(got, ag.UInt8ArrayGPU) and got.len == 5 and got.null_buffer is not None
    uploads = [a for n, a in lib.calls if n == "agpu_h2d"]
    assert np.frombuffer(uploads[0][1], np.uint8)[0] == 0b10110                      # rows: null, 7, 8, null, 10 (LSB first)
    assert list(np.frombuffer(uploads[1][1], np.uint8)[[1, 2, 4]]) == [7, 8, 10]


def test_device_round_trip_is_zero_copy_and_releases_once(stub):
    lib, dev = stub
    a = ag.Float32ArrayGPU.from_numpy(np.arange(64, dtype=np.float32), np.arange(64) % 3 != 0, dev)
    before = len(cd._exports)
    lib.calls.clear()
    b = ag.from_arrow_device(a, dev)                    # arrays are protocol objects
    assert type(b) is ag.Float32ArrayGPU and b.len == 64
    assert b.data.ptr == a.data.ptr and b.null_buffer.bit_buffer.ptr == a.null_buffer.bit_buffer.ptr
    names = [n for n, _ in lib.calls]
    assert "agpu_event_record" in names and "agpu_stream_wait_event" in names        # producer event, consumer wait
    assert "agpu_h2d" not in names and "agpu_alloc" not in names and "agpu_d2d" not in names
    gc.collect()
    assert len(cd._exports) == before + 1               # the array structure is alive (schema capsule already gone)
    data_ptr = a.data.ptr
    del a
    gc.collect()
    assert not any(n == "agpu_free" and args[1] == data_ptr for n, args in lib.calls), "exported buffers outlive the producer's handle"
    lib.calls.clear()
    del b
    gc.collect()
    assert len(cd._exports) == before                   # release ran once, after a sync of the consumer
    names = [n for n, _ in lib.calls]
    assert names.index("agpu_sync") < names.index("agpu_free")
    assert sum(1 for n, args in lib.calls if n == "agpu_free" and args[1] == data_ptr) == 1


def test_unconsumed_capsules_release_themselves(stub):
    _lib, dev = stub
    a = ag.Int32ArrayGPU.from_slice([1, 2, 3], dev)
    before = len(cd._exports)
    caps = ag.export_device(a).__arrow_c_device_array__()
    assert len(cd._exports) == before + 2
    del caps
    gc.collect()
    assert len(cd._exports) == before


def test_wrong_device_and_bad_offsets_are_refused(stub):
    _lib, dev = stub
    a = ag.BooleanArrayGPU.from_slice([True, False] * 40, dev)
    schema_capsule, array_capsule = a.__arrow_c_device_array__()
    _fmt, dptr = cd.parse_device_capsules(schema_capsule, array_capsule)
    dptr.contents.device_id = 5
    with pytest.raises(ValueError, match="CUDA device 5"):
        ag.from_arrow_device((schema_capsule, array_capsule), dev)
    dptr.contents.device_id = 0
    dptr.contents.array.offset = 8
    dptr.contents.array.length = 40
    with pytest.raises(ValueError, match="multiple of 32"):
        ag.from_arrow_device((schema_capsule, array_capsule), dev)
    dptr.contents.array.offset = 32
    got = ag.from_arrow_device((schema_capsule, array_capsule), dev)
    assert got.len == 40 and got.data.ptr == a.data.ptr + 4
    with pytest.raises(ValueError, match="already released"):
        ag.from_arrow_device((schema_capsule, array_capsule), dev)      # moved: the capsule's structure is spent
