"""pyarrow <-> device arrays (the in-memory format adjacent to the path).

The device layout IS the Arrow columnar layout (dense little-endian values, LSB-first validity
bitmap, 1 = valid), so import/export are plain buffer copies: host Arrow buffers -> H2D, D2H ->
`pa.Array.from_buffers`.  Replaces the reference's `python_wgarrow` stub, which exposes dtypes only
(crates/python_wgarrow/src/lib.rs:7-11)."""
from __future__ import annotations

import numpy as np

from .array import (ARRAY_TYPES, ArrowType, BooleanArrayGPU, GpuDevice, NullBitBufferGpu, PrimitiveArrayGpu,
                    bitmap_words)


def _types():
    import pyarrow as pa
    return {pa.int8(): ArrowType.Int8Type, pa.int16(): ArrowType.Int16Type, pa.int32(): ArrowType.Int32Type,
            pa.uint8(): ArrowType.UInt8Type, pa.uint16(): ArrowType.UInt16Type, pa.uint32(): ArrowType.UInt32Type,
            pa.float32(): ArrowType.Float32Type, pa.date32(): ArrowType.Date32Type, pa.bool_(): ArrowType.BooleanType}


def _padded_bitmap(buf, n_bits: int) -> np.ndarray:
    out = np.zeros(bitmap_words(n_bits) * 4, dtype=np.uint8)
    raw = np.frombuffer(buf, dtype=np.uint8)[: (n_bits + 7) // 8]
    out[: len(raw)] = raw
    if n_bits % 8:  # Arrow leaves padding bits unspecified: clear them
        out[(n_bits - 1) // 8] &= (1 << (n_bits % 8)) - 1
    return out


def from_arrow(arr, device: GpuDevice):
    """pyarrow.Array (int8..uint32, float32, date32, bool) -> device array of the matching type"""
    import pyarrow as pa
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks()
    if arr.offset != 0:
        arr = pa.concat_arrays([arr])  # re-base to offset 0 (bitmaps cannot be sliced on a byte boundary)
    at = _types().get(arr.type)
    if at is None:
        raise TypeError(f"unsupported Arrow type {arr.type}")
    cls = ARRAY_TYPES[at]
    n = len(arr)
    validity, data = arr.buffers()[0], arr.buffers()[1]
    nb = None
    if validity is not None and arr.null_count:
        nb = NullBitBufferGpu(device.create_gpu_buffer_with_data(_padded_bitmap(validity, n)), n, device)
    if cls is BooleanArrayGPU:
        return cls(device.create_gpu_buffer_with_data(_padded_bitmap(data, n)), device, n, nb)
    values = np.frombuffer(data, dtype=cls.NP, count=n) if n else np.zeros(0, cls.NP)
    return cls(device.create_gpu_buffer_with_data(values), device, n, nb)


def to_arrow(array):
    """device array -> pyarrow.Array (one D2H per buffer)"""
    import pyarrow as pa
    rev = {v: k for k, v in _types().items()}
    ty = rev[array.get_dtype()]
    dev = array.gpu_device
    validity = None
    if array.null_buffer is not None:
        validity = pa.py_buffer(dev.retrive_data(array.null_buffer.bit_buffer, bitmap_words(array.len) * 4).tobytes())
    if isinstance(array, BooleanArrayGPU):
        data = pa.py_buffer(dev.retrive_data(array.data, bitmap_words(array.len) * 4).tobytes())
    else:
        assert isinstance(array, PrimitiveArrayGpu)
        data = pa.py_buffer(dev.retrive_data(array.data, array.len * array.NP.itemsize).tobytes())
    return pa.Array.from_buffers(ty, array.len, [validity, data])
