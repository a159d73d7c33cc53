"""Arrow C Device Data Interface (`ArrowDeviceArray`, ARROW_DEVICE_CUDA) for the device arrays:
ZERO-COPY exchange of device-resident columns with any other library on the same GPU
(SURVEY.md §8f rank 4 — "lets pyarrow feed columns zero-copy"; `interop.from_arrow/to_arrow` are
the copying host-memory variants).

The device layout of a column IS the Arrow layout (dense little-endian values, LSB-first validity
bitmap, 1 = valid, buffers padded to whole 32-bit words), so

  export   hands out the two device pointers of an array as they are; the exported structure keeps
           the buffers alive until the consumer calls `release`; `sync_event` is a CUDA event
           recorded on the producing handle's stream (an `agpu_event*` is layout-compatible with
           the `cudaEvent_t*` the specification asks for);
  import   wraps the producer's device pointers in non-owning `ArrowGpuBuffer`s, makes the
           consuming handle's stream wait for `sync_event`, and calls the producer's `release`
           when the last imported buffer is dropped — after the consuming stream has drained.

Both directions speak the Arrow PyCapsule protocol (`__arrow_c_device_array__`), so any producer or
consumer of that protocol can be on the other side.  A column in host memory (ARROW_DEVICE_CPU /
CUDA_HOST, what a CPU-only pyarrow exports) is imported by copying, like `from_arrow`.

The struct definitions follow the Arrow specification ("The Arrow C data interface", "The Arrow C
Device data interface"); the CPU tests check them against pyarrow's own importer and exporter."""
from __future__ import annotations

import ctypes as C
import itertools
import threading
from typing import Optional

import numpy as np

from .array import (ARRAY_TYPES, ArrowGpuBuffer, ArrowType, BooleanArrayGPU, GpuDevice, GpuEvent, NullBitBufferGpu,
                    PrimitiveArrayGpu, bitmap_words)

ARROW_DEVICE_CPU, ARROW_DEVICE_CUDA, ARROW_DEVICE_CUDA_HOST = 1, 2, 3
ARROW_FLAG_NULLABLE = 2


class ArrowSchema(C.Structure):
    pass


class ArrowArray(C.Structure):
    pass


_SchemaRelease = C.CFUNCTYPE(None, C.POINTER(ArrowSchema))
_ArrayRelease = C.CFUNCTYPE(None, C.POINTER(ArrowArray))

ArrowSchema._fields_ = [("format", C.c_char_p), ("name", C.c_char_p), ("metadata", C.c_char_p), ("flags", C.c_int64),
                        ("n_children", C.c_int64), ("children", C.POINTER(C.POINTER(ArrowSchema))),
                        ("dictionary", C.POINTER(ArrowSchema)), ("release", _SchemaRelease), ("private_data", C.c_void_p)]
ArrowArray._fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64), ("n_buffers", C.c_int64),
                       ("n_children", C.c_int64), ("buffers", C.POINTER(C.c_void_p)),
                       ("children", C.POINTER(C.POINTER(ArrowArray))), ("dictionary", C.POINTER(ArrowArray)),
                       ("release", _ArrayRelease), ("private_data", C.c_void_p)]


class ArrowDeviceArray(C.Structure):
    _fields_ = [("array", ArrowArray), ("device_id", C.c_int64), ("device_type", C.c_int32), ("sync_event", C.c_void_p),
                ("reserved", C.c_int64 * 3)]


# Arrow format strings (specification, "Data type description — format strings")
_FORMATS = {ArrowType.Int8Type: b"c", ArrowType.UInt8Type: b"C", ArrowType.Int16Type: b"s", ArrowType.UInt16Type: b"S",
            ArrowType.Int32Type: b"i", ArrowType.UInt32Type: b"I", ArrowType.Float32Type: b"f",
            ArrowType.Date32Type: b"tdD", ArrowType.BooleanType: b"b"}
_BY_FORMAT = {v: k for k, v in _FORMATS.items()}

_libc = C.CDLL(None)
_libc.malloc.restype, _libc.malloc.argtypes = C.c_void_p, [C.c_size_t]
_libc.free.restype, _libc.free.argtypes = None, [C.c_void_p]

# PyCapsule calls on RAW object addresses (a destructor runs while its capsule is being destroyed:
# it must not be turned into a counted Python reference again)
_capsule_new = C.PYFUNCTYPE(C.py_object, C.c_void_p, C.c_char_p, C.c_void_p)(("PyCapsule_New", C.pythonapi))
_capsule_ptr_raw = C.PYFUNCTYPE(C.c_void_p, C.c_void_p, C.c_char_p)(("PyCapsule_GetPointer", C.pythonapi))
_capsule_ptr = C.PYFUNCTYPE(C.c_void_p, C.py_object, C.c_char_p)(("PyCapsule_GetPointer", C.pythonapi))
_CapsuleDestructor = C.CFUNCTYPE(None, C.c_void_p)
_SCHEMA_NAME, _DEVICE_ARRAY_NAME = b"arrow_schema", b"arrow_device_array"

# what an exported structure keeps alive, keyed by the integer stored in its private_data
_exports: dict = {}
_export_ids = itertools.count(1)
_exports_lock = threading.Lock()


@_ArrayRelease
def _release_exported_array(array_ptr):
    a = array_ptr.contents
    with _exports_lock:
        _exports.pop(a.private_data, None)       # drops the buffers (stream-ordered free) and the event
    a.release = _ArrayRelease()


@_SchemaRelease
def _release_exported_schema(schema_ptr):
    s = schema_ptr.contents
    with _exports_lock:
        _exports.pop(s.private_data, None)
    s.release = _SchemaRelease()


@_CapsuleDestructor
def _destroy_schema_capsule(capsule):
    ptr = _capsule_ptr_raw(capsule, _SCHEMA_NAME)
    if ptr:
        s = C.cast(ptr, C.POINTER(ArrowSchema))
        if s.contents.release:
            s.contents.release(s)
        _libc.free(ptr)


@_CapsuleDestructor
def _destroy_device_array_capsule(capsule):
    ptr = _capsule_ptr_raw(capsule, _DEVICE_ARRAY_NAME)
    if ptr:
        d = C.cast(ptr, C.POINTER(ArrowDeviceArray))
        if d.contents.array.release:
            d.contents.array.release(C.pointer(d.contents.array))
        _libc.free(ptr)


def _keep(objects) -> int:
    with _exports_lock:
        key = next(_export_ids)
        _exports[key] = objects
    return key


def _new_schema(fmt: bytes, nullable: bool) -> int:
    """malloc'ed ArrowSchema for a primitive column; returns its address"""
    ptr = _libc.malloc(C.sizeof(ArrowSchema))
    s = C.cast(ptr, C.POINTER(ArrowSchema)).contents
    C.memset(ptr, 0, C.sizeof(ArrowSchema))
    fmt_buf, name_buf = C.create_string_buffer(fmt), C.create_string_buffer(b"")
    s.format = C.cast(fmt_buf, C.c_char_p)
    s.name = C.cast(name_buf, C.c_char_p)
    s.flags = ARROW_FLAG_NULLABLE if nullable else 0
    s.release = _release_exported_schema
    s.private_data = _keep((fmt_buf, name_buf))
    return ptr


def _new_device_array(length: int, null_count: int, pointers, keepalive, device_type: int, device_id: int,
                      sync_event: Optional[int]) -> int:
    """malloc'ed ArrowDeviceArray over [validity, data] pointers; `keepalive` lives until release"""
    ptr = _libc.malloc(C.sizeof(ArrowDeviceArray))
    C.memset(ptr, 0, C.sizeof(ArrowDeviceArray))
    d = C.cast(ptr, C.POINTER(ArrowDeviceArray)).contents
    buffers = (C.c_void_p * 2)(*pointers)
    d.array.length, d.array.null_count, d.array.offset = length, null_count, 0
    d.array.n_buffers, d.array.n_children = 2, 0
    d.array.buffers = C.cast(buffers, C.POINTER(C.c_void_p))
    d.array.release = _release_exported_array
    d.array.private_data = _keep((buffers, keepalive))
    d.device_id, d.device_type, d.sync_event = device_id, device_type, sync_event
    return ptr


class DeviceArrayExport:
    """what `export_device` returns: speaks `__arrow_c_device_array__` (and `__arrow_c_schema__`)"""

    def __init__(self, array):
        self._array = array

    def __arrow_c_schema__(self):
        a = self._array
        return _capsule_new(_new_schema(_FORMATS[a.get_dtype()], a.null_buffer is not None), _SCHEMA_NAME,
                            C.cast(_destroy_schema_capsule, C.c_void_p))

    def __arrow_c_device_array__(self, requested_schema=None, **kwargs):
        a = self._array
        dev = a.gpu_device
        nb = a.null_buffer                     # (launches a pending recorded chain)
        data = a.data
        event = dev.record_event()             # everything enqueued so far, i.e. the producer of both buffers
        # null_count = -1: "not computed" — counting would need a kernel and a read-back
        ptr = _new_device_array(a.len, -1 if nb is not None else 0, [nb.bit_buffer.ptr if nb is not None else None, data.ptr],
                                (data, nb, event), ARROW_DEVICE_CUDA, dev.ordinal, event.handle.value)
        return (self.__arrow_c_schema__(),
                _capsule_new(ptr, _DEVICE_ARRAY_NAME, C.cast(_destroy_device_array_capsule, C.c_void_p)))


def export_device(array) -> DeviceArrayExport:
    """device array -> object implementing the Arrow PyCapsule device protocol; zero-copy: the
    consumer sees this array's own device buffers and must not write to them"""
    if not isinstance(array, (PrimitiveArrayGpu, BooleanArrayGPU)):
        raise TypeError(f"cannot export {type(array).__name__}")
    return DeviceArrayExport(array)


class _Imported:
    """the consumer's side of one imported ArrowDeviceArray: the moved structure, released once"""

    def __init__(self, moved_ptr: int, device: GpuDevice):
        self.ptr, self.device = moved_ptr, device

    def __del__(self):
        ptr, self.ptr = self.ptr, None
        if not ptr:
            return
        d = C.cast(ptr, C.POINTER(ArrowDeviceArray)).contents
        if d.array.release:
            try:
                if self.device.handle:
                    self.device.sync()          # kernels of this handle may still read the producer's memory
            except Exception as exc:            # a failed wait must not leak the producer's buffers
                import warnings
                warnings.warn(f"arrow_gpu_b200: sync before releasing an imported ArrowDeviceArray failed: {exc}")
            d.array.release(C.pointer(d.array))
        _libc.free(ptr)


class _ForeignBuffer(ArrowGpuBuffer):
    """device memory owned by the producer of an imported array"""
    __slots__ = ("_import",)

    def __init__(self, device: GpuDevice, ptr: int, size: int, imported: _Imported):
        super().__init__(device, ptr, size, owned=False, kind="foreign")
        self._import = imported


def parse_device_capsules(schema_capsule, array_capsule):
    """-> (format bytes, pointer to the ArrowDeviceArray inside the capsule)"""
    sp = _capsule_ptr(schema_capsule, _SCHEMA_NAME)
    ap = _capsule_ptr(array_capsule, _DEVICE_ARRAY_NAME)
    if not sp or not ap:
        raise ValueError("not an (arrow_schema, arrow_device_array) capsule pair")
    schema = C.cast(sp, C.POINTER(ArrowSchema)).contents
    return schema.format, C.cast(ap, C.POINTER(ArrowDeviceArray))


def from_arrow_device(obj, device: GpuDevice):
    """any object with `__arrow_c_device_array__` (or a (schema capsule, device array capsule) pair)
    -> device array.  CUDA memory of `device`'s GPU is adopted without a copy; host memory is copied."""
    schema_capsule, array_capsule = obj if isinstance(obj, tuple) else obj.__arrow_c_device_array__()
    fmt, src = parse_device_capsules(schema_capsule, array_capsule)
    at = _BY_FORMAT.get(fmt)
    if at is None:
        raise TypeError(f"unsupported Arrow format {fmt!r}")
    cls = ARRAY_TYPES[at]
    d = src.contents
    a = d.array
    if not a.release:
        raise ValueError("the ArrowDeviceArray was already released")
    if a.n_children or a.dictionary or a.n_buffers != 2:
        raise TypeError("only primitive and boolean columns are supported")
    n, off = a.length, a.offset
    item_bits = 1 if cls is BooleanArrayGPU else cls.ITEMSIZE * 8
    validity_ptr, data_ptr = a.buffers[0], a.buffers[1]
    has_nulls = bool(validity_ptr) and a.null_count != 0

    if d.device_type in (ARROW_DEVICE_CPU, ARROW_DEVICE_CUDA_HOST):
        # host memory: copy (like interop.from_arrow); the structure stays with the caller's capsule
        def bits(ptr):
            raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=((off + n + 7) // 8,)) if n else np.zeros(0, np.uint8)
            flags = np.unpackbits(raw, bitorder="little")[off: off + n].astype(bool)
            from .array import pack_bits
            return pack_bits(flags)
        nb = NullBitBufferGpu(device.create_gpu_buffer_with_data(bits(validity_ptr)), n, device) if has_nulls else None
        if cls is BooleanArrayGPU:
            return cls(device.create_gpu_buffer_with_data(bits(data_ptr)), device, n, nb)
        nbytes = n * cls.ITEMSIZE
        raw = np.ctypeslib.as_array(C.cast(data_ptr + off * cls.ITEMSIZE, C.POINTER(C.c_uint8)), shape=(nbytes,)) if n else np.zeros(0, np.uint8)
        return cls(device.create_gpu_buffer_with_data(raw.view(cls.NP)), device, n, nb)

    if d.device_type != ARROW_DEVICE_CUDA:
        raise TypeError(f"unsupported ArrowDeviceType {d.device_type}")
    if d.device_id != device.ordinal:
        raise ValueError(f"the column lives on CUDA device {d.device_id}, the handle is on {device.ordinal}")
    # bitmaps are read as whole 32-bit words: an offset must keep them word-aligned
    if (has_nulls or cls is BooleanArrayGPU) and off % 32:
        raise ValueError("zero-copy import needs a bitmap offset that is a multiple of 32 rows")
    if (off * item_bits) % 8:
        raise ValueError("unaligned offset")

    # move the structure (specification: "moving an array"): copy it, mark the source released
    moved = _libc.malloc(C.sizeof(ArrowDeviceArray))
    C.memmove(moved, src, C.sizeof(ArrowDeviceArray))
    a.release = _ArrayRelease()
    imported = _Imported(moved, device)
    if d.sync_event:
        # an agpu_event* is a pointer to a cudaEvent_t, which is what sync_event points to
        from ._ffi import check, lib
        check(lib().agpu_stream_wait_event(device.handle, C.c_void_p(d.sync_event)), "agpu_stream_wait_event")
    nb = None
    if has_nulls:
        nb = NullBitBufferGpu(_ForeignBuffer(device, validity_ptr + off // 8, bitmap_words(n) * 4, imported), n, device)
    if cls is BooleanArrayGPU:
        return cls(_ForeignBuffer(device, data_ptr + off // 8, bitmap_words(n) * 4, imported), device, n, nb)
    return cls(_ForeignBuffer(device, data_ptr + off * cls.ITEMSIZE, n * cls.ITEMSIZE, imported), device, n, nb)


def _array_device_capsules(self, requested_schema=None, **kwargs):
    return DeviceArrayExport(self).__arrow_c_device_array__(requested_schema, **kwargs)


def _array_schema_capsule(self):
    return DeviceArrayExport(self).__arrow_c_schema__()


# the arrays themselves are protocol objects: any consumer of `__arrow_c_device_array__` takes them
for _cls in (PrimitiveArrayGpu, BooleanArrayGPU):
    _cls.__arrow_c_device_array__ = _array_device_capsules
    _cls.__arrow_c_schema__ = _array_schema_capsule
