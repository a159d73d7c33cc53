"""Array / buffer layer: the host-side mirror of the reference's `crates/array`.

Same type and method names as the reference so its tests read the same here:
`GpuDevice`, `ArrowComputePipeline`, `ArrowGpuBuffer`, `BooleanBufferBuilder`,
`NullBitBufferGpu`, `PrimitiveArrayGpu` (+ the eight typed aliases), `BooleanArrayGPU`,
`ArrowType`, `ScalarValue`, `ArrowErrorGPU`, `GPU_DEVICE`, `broadcast_dyn`.

What changed underneath (BASELINE.json north_star (1)): buffers are stream-ordered CUDA
allocations owned through the C ABI (include/agpu.h) instead of `Arc<wgpu::Buffer>`; a
`GpuDevice` is `{ordinal, cudaStream_t, cudaMemPool_t}`; `ArrowComputePipeline` is a
stream scope whose `finish()` has nothing left to submit.
"""
from __future__ import annotations

import atexit
import ctypes as C
import enum
from typing import Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import check, lib

_byref = C.byref


class ArrowErrorGPU(Exception):
    """crates/array/src/lib.rs:10-13"""


class OperationNotSupported(ArrowErrorGPU):
    pass


class CastingNotSupported(ArrowErrorGPU):
    pass


class Panic(RuntimeError):
    """What the reference does with `panic!` on unsupported dtype pairs
    (e.g. crates/arithmetic/src/arithmetic_kernels.rs:92-97)."""


class ArrowType(enum.Enum):
    """crates/array/src/array/mod.rs:40-50"""
    BooleanType = _ffi.BOOL
    Float32Type = _ffi.F32
    UInt32Type = _ffi.U32
    UInt16Type = _ffi.U16
    UInt8Type = _ffi.U8
    Int32Type = _ffi.I32
    Int16Type = _ffi.I16
    Int8Type = _ffi.I8
    Date32Type = _ffi.DATE32


# At interpreter exit the CUDA runtime(s) of the process are being torn down in an order we do
# not control (ours is linked statically, torch brings its own): objects that are still alive
# then must not call back into the library.  The OS reclaims device memory with the process.
_SHUTDOWN = False


def _mark_shutdown() -> None:
    global _SHUTDOWN
    _SHUTDOWN = True


atexit.register(_mark_shutdown)


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


# ------------------------------------------------------------------------------------------
# device + buffers  (crates/array/src/gpu_utils/gpu_device.rs, array/buffer.rs)
# ------------------------------------------------------------------------------------------
class GpuDevice:
    """gpu_device.rs:29-33 — here: one CUDA device ordinal + stream + memory pool."""

    def __init__(self, ordinal: int = 0):
        h = C.c_void_p()
        check(lib().agpu_device_create(ordinal, C.byref(h)), "agpu_device_create")
        self.handle = h
        self.ordinal = ordinal

    @classmethod
    def new(cls) -> "GpuDevice":
        return cls(0)

    def destroy(self) -> None:
        """Release the stream.  Never implicit: buffers, events and (when torch interop is used)
        torch-side objects may still refer to the stream when the Python handle is collected, so a
        device handle lives until the process ends unless the owner ends it explicitly."""
        if getattr(self, "handle", None) and not _SHUTDOWN:
            lib().agpu_device_destroy(self.handle)
        self.handle = None

    # --- buffers
    def create_empty_buffer(self, size: int) -> "ArrowGpuBuffer":
        """gpu_device.rs:183-192 (not zero-filled: every kernel writes its whole output)"""
        p = C.c_void_p()
        rc = (_ffi._lib or lib()).agpu_alloc(self.handle, (size + 15) & -16 or 16, _byref(p))   # hot: one call per output
        if rc:
            check(rc, "agpu_alloc")
        return ArrowGpuBuffer(self, p.value, size)

    def create_gpu_buffer_with_data(self, data: np.ndarray, wait: bool = True) -> "ArrowGpuBuffer":
        """gpu_device.rs:171-181.  `wait=False` is for pinned host arrays the caller keeps alive
        (see `pinned_empty`): the copy is then fully asynchronous on the device's stream."""
        data = np.ascontiguousarray(data)
        buf = self.create_empty_buffer(data.nbytes)
        if data.nbytes:
            check(lib().agpu_h2d(self.handle, buf.ptr, data.ctypes.data, data.nbytes), "agpu_h2d")
            if wait:
                # the host array may be a temporary: wait so it can be dropped (pageable copies
                # are staged synchronously by the driver anyway)
                check(lib().agpu_sync(self.handle), "agpu_sync")
        return buf

    @staticmethod
    def pinned_empty(n: int, dtype) -> np.ndarray:
        """page-locked host array for from_numpy(..., wait=False) / raw_values(out=...)"""
        dt = np.dtype(dtype)
        p = C.c_void_p()
        check(lib().agpu_host_alloc(max(n * dt.itemsize, 1), C.byref(p)), "agpu_host_alloc")
        raw = (C.c_uint8 * max(n * dt.itemsize, 1)).from_address(p.value)
        arr = np.frombuffer(raw, dtype=dt, count=n)
        _PINNED[arr.ctypes.data] = (p.value, raw)   # keep the mapping alive; freed by pinned_free
        return arr

    @staticmethod
    def pinned_free(arr: np.ndarray) -> None:
        ent = _PINNED.pop(arr.ctypes.data, None)
        if ent:
            lib().agpu_host_free(ent[0])

    def read_into(self, buffer: "ArrowGpuBuffer", out: np.ndarray, wait: bool = True) -> None:
        """device -> caller-provided (ideally pinned) host array"""
        fn = lib().agpu_d2h if wait else lib().agpu_d2h_async
        check(fn(self.handle, out.ctypes.data, buffer.ptr, out.nbytes), "agpu_d2h")

    def create_scalar_buffer(self, value) -> "ArrowGpuBuffer":
        """gpu_device.rs:203-210"""
        return self.create_gpu_buffer_with_data(np.asarray([value]))

    def create_ipc_buffer_with_data(self, data: np.ndarray) -> "ArrowGpuBuffer":
        """like create_gpu_buffer_with_data but in memory other GPUs of the box can map (CUDA IPC)"""
        data = np.ascontiguousarray(data)
        p = C.c_void_p()
        check(lib().agpu_ipc_alloc(self.handle, _round_up(max(data.nbytes, 1), 16), C.byref(p)), "agpu_ipc_alloc")
        buf = ArrowGpuBuffer(self, p.value, data.nbytes, kind="ipc")
        if data.nbytes:
            check(lib().agpu_h2d(self.handle, buf.ptr, data.ctypes.data, data.nbytes), "agpu_h2d")
            check(lib().agpu_sync(self.handle), "agpu_sync")
        return buf

    def ipc_export(self, buffer: "ArrowGpuBuffer") -> bytes:
        h = C.create_string_buffer(64)
        check(lib().agpu_ipc_export(self.handle, buffer.ptr, h), "agpu_ipc_export")
        return h.raw

    def ipc_open(self, handle: bytes, size: int) -> "ArrowGpuBuffer":
        p = C.c_void_p()
        check(lib().agpu_ipc_open(self.handle, handle, C.byref(p)), "agpu_ipc_open")
        return ArrowGpuBuffer(self, p.value, size, kind="peer")

    def clone_buffer(self, buffer: "ArrowGpuBuffer") -> "ArrowGpuBuffer":
        """gpu_device.rs:212-222"""
        out = self.create_empty_buffer(buffer.size)
        check(lib().agpu_d2d(self.handle, out.ptr, buffer.ptr, buffer.size), "agpu_d2d")
        return out

    def retrive_data(self, buffer: "ArrowGpuBuffer", nbytes: Optional[int] = None) -> np.ndarray:
        """gpu_device.rs:232-265 — the only host synchronisation point"""
        n = buffer.size if nbytes is None else nbytes
        if 0 < n <= 256:
            # counts, flags, one-element results: through a pinned staging word (a pageable
            # destination makes the driver stage the copy, ~15 us on top of the synchronisation)
            stage = getattr(self, "_small_pinned", None)
            if stage is None:
                stage = self._small_pinned = self.pinned_empty(256, np.uint8)
            check(lib().agpu_d2h(self.handle, stage.ctypes.data, buffer.ptr, n), "agpu_d2h")
            return stage[:n].copy()
        out = np.empty(n, dtype=np.uint8)
        check(lib().agpu_d2h(self.handle, out.ctypes.data, buffer.ptr, n), "agpu_d2h")
        return out

    def sync(self) -> None:
        check(lib().agpu_sync(self.handle), "agpu_sync")

    # --- events (CmpQuery of the reference, gpu_utils/compute_query.rs) and cross-stream order
    def record_event(self, event: Optional["GpuEvent"] = None) -> "GpuEvent":
        event = event or GpuEvent()
        check(lib().agpu_event_record(self.handle, event.handle), "agpu_event_record")
        return event

    def wait_event(self, event: "GpuEvent") -> None:
        """later work of this device handle waits on the GPU for `event` (recorded by another
        handle of the same GPU, e.g. an upload stream)"""
        check(lib().agpu_stream_wait_event(self.handle, event.handle), "agpu_stream_wait_event")

    @property
    def stream_ptr(self) -> int:
        """the cudaStream_t of this handle (e.g. for torch.cuda.ExternalStream)"""
        return int(lib().agpu_device_stream(self.handle) or 0)

    def launch_count(self) -> int:
        return int(lib().agpu_launch_count(self.handle))

    def record_use(self, buffer: "ArrowGpuBuffer") -> None:
        """this handle has enqueued (or is about to enqueue) work on a buffer that ANOTHER handle of
        the same GPU allocated: when the buffer is dropped, its block only returns to the owner's
        cache after this handle's stream got there (agpu_buffer_record_use; wgpu keeps a buffer
        alive until submitted work is done, buffer.rs:5-7)"""
        check(lib().agpu_buffer_record_use(self.handle, buffer.ptr), "agpu_buffer_record_use")


class GpuEvent:
    def __init__(self):
        h = C.c_void_p()
        check(lib().agpu_event_create(C.byref(h)), "agpu_event_create")
        self.handle = h

    def elapsed_ms(self, later: "GpuEvent") -> float:
        ms = C.c_float(0)
        check(lib().agpu_event_elapsed_ms(self.handle, later.handle, C.byref(ms)), "agpu_event_elapsed_ms")
        return ms.value

    def __del__(self):
        try:
            if self.handle and not _SHUTDOWN:
                lib().agpu_event_destroy(self.handle)
        except Exception:
            pass
        self.handle = None


class ArrowGpuBuffer:
    """array/buffer.rs:5-7 — owns one device allocation; dropping it frees stream-ordered."""

    __slots__ = ("device", "ptr", "_size", "_owned", "_kind", "__weakref__")

    def __init__(self, device: GpuDevice, ptr: int, size: int, owned: bool = True, kind: str = "pool"):
        # kind: "pool" = stream-ordered pool allocation; "ipc" = cudaMalloc'd, exportable to the
        # other GPUs of the box; "peer" = another process's buffer opened through CUDA IPC
        self.device, self.ptr, self._size, self._owned, self._kind = device, ptr, size, owned, kind

    @property
    def size(self) -> int:
        """bytes (buffer.rs:22-24)"""
        return self._size

    def free(self) -> None:
        """explicit drop (idempotent; __del__ does the same)"""
        self.__del__()

    def __del__(self):
        try:
            if self._owned and self.ptr and self.device.handle and not _SHUTDOWN:
                if self._kind == "ipc":
                    lib().agpu_ipc_free(self.device.handle, self.ptr)
                elif self._kind == "peer":
                    lib().agpu_ipc_close(self.device.handle, self.ptr)
                else:
                    (_ffi._lib or lib()).agpu_free(self.device.handle, self.ptr)
        except Exception:
            pass
        self.ptr = None


class ArrowComputePipeline:
    """gpu_utils/compute_pipeline.rs:8-22.  The reference records compute passes into one
    wgpu CommandEncoder and submits them in `finish()`.  CUDA streams are already ordered
    queues, so by default every `*_op` enqueues its kernel immediately on the device's stream and
    `finish()` only marks the end of the scope (it never waits, like the reference).

    capture=True is the literal analogue of record-then-submit: the ops recorded between the
    constructor and `finish()` are captured into a CUDA graph (agpu_graph_begin/end) and `finish()`
    submits the whole program with ONE driver call; `replay()` submits the same recorded program
    again (same input buffers, outputs overwritten in place) for the cost of one graph launch
    instead of one launch per op.  Inside a capture nothing may read back to the host."""

    def __init__(self, device: GpuDevice, label: Optional[str] = None, profile: bool = False, fuse: bool = False,
                 capture: bool = False):
        self.device = device
        self.label = label
        self.finished = False
        # fuse=True: f32 op chains recorded on this pipeline (cast -> unary/binary/scalar ops ->
        # optional compare) are not launched one by one; each maximal linear chain becomes ONE
        # agpu_fused_chain kernel at finish() (or earlier, when a result is read).  Same results
        # bit for bit as fuse=False (kernels.py, "auto-fusion").
        self.fuse = fuse
        self._lazies: list = []
        # the reference's `profile` cargo feature wraps every compute pass in timestamp queries
        # (gpu_utils/compute_query.rs:3-90); here: a CUDA event pair per recorded op
        self.profile = profile
        self.queries: list = []
        self.capture = capture
        self.graph: Optional["GpuGraph"] = None
        if capture:
            if profile:
                raise ValueError("profile=True records events between ops; it cannot be combined with capture=True")
            check(lib().agpu_graph_begin(device.handle), "agpu_graph_begin")

    @classmethod
    def new(cls, device: GpuDevice, label: Optional[str] = None) -> "ArrowComputePipeline":
        return cls(device, label)

    def wait_for_results(self) -> list:
        """compute_query.rs:54-75: [(op name, milliseconds)] of every op recorded with profile=True"""
        return [(name, start.elapsed_ms(stop)) for name, start, stop in self.queries]

    def clone_buffer(self, buffer: ArrowGpuBuffer) -> ArrowGpuBuffer:
        return self.device.clone_buffer(buffer)

    def flush_recorded(self) -> None:
        """launch every chain recorded with fuse=True that nothing absorbed (their results must exist
        before an in-place op such as put_op touches the buffers they read)"""
        for ref in self._lazies:
            arr = ref()
            if arr is not None and arr._lazy is not None and not arr._lazy.consumed:
                arr._materialize()
        self._lazies = []

    def finish(self) -> None:
        """compute_pipeline.rs:259-273.  With fuse=True this launches one fused kernel per recorded
        chain whose result is still referenced and was not absorbed into a longer chain.  With
        capture=True it ends the capture and submits the recorded program once."""
        self.flush_recorded()
        if self.capture and not self.finished:
            h = C.c_void_p()
            check(lib().agpu_graph_end(self.device.handle, C.byref(h)), "agpu_graph_end")
            self.graph = GpuGraph(self.device, h)
            self.graph.launch()
        self.finished = True

    def replay(self) -> None:
        """submit the recorded program again (capture=True pipelines, after finish())"""
        if self.graph is None:
            raise RuntimeError("replay() needs a pipeline created with capture=True and finished")
        self.graph.launch()

    def abort(self) -> None:
        """leave a capture without submitting (error paths)"""
        if self.capture and not self.finished:
            h = C.c_void_p()
            if lib().agpu_graph_end(self.device.handle, C.byref(h)) == 0 and h:
                lib().agpu_graph_destroy(h)
            self.finished = True


class GpuGraph:
    """a captured pipeline: agpu_graph (cudaGraphExec_t + the temporaries its kernels write)"""

    def __init__(self, device: GpuDevice, handle):
        self.device, self.handle = device, handle
        self.kernels = int(lib().agpu_graph_kernel_count(handle))

    def launch(self) -> None:
        check(lib().agpu_graph_launch(self.device.handle, self.handle), "agpu_graph_launch")

    def __del__(self):
        try:
            if self.handle and self.device.handle and not _SHUTDOWN:
                lib().agpu_graph_destroy(self.handle)
        except Exception:
            pass
        self.handle = None


_PINNED: dict = {}
_GPU_DEVICE: Optional[GpuDevice] = None


def GPU_DEVICE() -> GpuDevice:
    """crates/array/src/lib.rs:16-17 (LazyLock<Arc<GpuDevice>>)"""
    global _GPU_DEVICE
    if _GPU_DEVICE is None:
        _GPU_DEVICE = GpuDevice(0)
    return _GPU_DEVICE


# ------------------------------------------------------------------------------------------
# bitmaps  (crates/array/src/array/null_bit_buffer.rs)
# ------------------------------------------------------------------------------------------
def bitmap_words(n_bits: int) -> int:
    return (n_bits + 31) // 32


def pack_bits(flags: np.ndarray) -> np.ndarray:
    """bool[n] -> LSB-first bitmap padded to whole u32 words (null_bit_buffer.rs:47-49)."""
    n = len(flags)
    out = np.zeros(bitmap_words(n) * 4, dtype=np.uint8)
    if n:
        packed = np.packbits(np.asarray(flags, dtype=bool), bitorder="little")
        out[: len(packed)] = packed
    return out


def unpack_bits(raw: np.ndarray, n_bits: int) -> np.ndarray:
    return np.unpackbits(np.asarray(raw, dtype=np.uint8), bitorder="little", count=None)[:n_bits].astype(bool)


class BooleanBufferBuilder:
    """null_bit_buffer.rs:10-62"""

    def __init__(self, size: int = 1024, set_all: bool = False):
        self.len = size
        self.data = np.zeros(_round_up(size, 8) // 8, dtype=np.uint8)
        self.contains_nulls = True
        if set_all:
            self.data[:] = 0xFF
            if size % 8:
                self.data[-1] = 0xFF >> (8 - size % 8)
            self.contains_nulls = False

    @classmethod
    def new_with_capacity(cls, size: int) -> "BooleanBufferBuilder":
        return cls(size)

    @classmethod
    def new_set_with_capacity(cls, size: int) -> "BooleanBufferBuilder":
        return cls(size, set_all=True)

    def set_bit(self, pos: int) -> None:
        self.data[pos // 8] |= 1 << (pos % 8)

    def unset_bit(self, pos: int) -> None:
        self.data[pos // 8] &= ~(1 << (pos % 8)) & 0xFF

    def is_set(self, pos: int) -> bool:
        return bool(self.data[pos // 8] & (1 << (pos % 8)))

    @staticmethod
    def is_set_in_slice(data, pos: int) -> bool:
        return bool(data[pos // 8] & (1 << (pos % 8)))


class NullBitBufferGpu:
    """null_bit_buffer.rs:91-96 — validity bitmap on the device (1 = valid)."""

    def __init__(self, bit_buffer: ArrowGpuBuffer, length: int, gpu_device: GpuDevice):
        self.bit_buffer, self.len, self.gpu_device = bit_buffer, length, gpu_device

    @classmethod
    def new(cls, gpu_device: GpuDevice, builder: BooleanBufferBuilder) -> Optional["NullBitBufferGpu"]:
        if not builder.contains_nulls:
            return None
        return cls.from_flags(gpu_device, unpack_bits(builder.data, builder.len))

    @classmethod
    def from_flags(cls, gpu_device: GpuDevice, valid: np.ndarray) -> "NullBitBufferGpu":
        return cls(gpu_device.create_gpu_buffer_with_data(pack_bits(valid)), len(valid), gpu_device)

    @classmethod
    def new_set_with_capacity(cls, gpu_device: GpuDevice, size: int) -> "NullBitBufferGpu":
        return cls.from_flags(gpu_device, np.ones(size, dtype=bool))

    def raw_values(self) -> np.ndarray:
        """bytes of the bitmap, ceil(len/8) of them (null_bit_buffer.rs:124-128)"""
        return self.gpu_device.retrive_data(self.bit_buffer, bitmap_words(self.len) * 4)[: _round_up(self.len, 8) // 8]

    def flags(self) -> np.ndarray:
        return unpack_bits(self.gpu_device.retrive_data(self.bit_buffer, bitmap_words(self.len) * 4), self.len)

    @staticmethod
    def clone_null_bit_buffer(data: Optional["NullBitBufferGpu"]) -> Optional["NullBitBufferGpu"]:
        if data is None:
            return None
        return NullBitBufferGpu(data.gpu_device.clone_buffer(data.bit_buffer), data.len, data.gpu_device)

    clone_null_bit_buffer_pass = clone_null_bit_buffer

    @staticmethod
    def clone_null_bit_buffer_op(data, pipeline) -> Optional["NullBitBufferGpu"]:
        return NullBitBufferGpu.clone_null_bit_buffer(data)

    @staticmethod
    def merge_null_bit_buffer(left: Optional["NullBitBufferGpu"],
                              right: Optional["NullBitBufferGpu"]) -> Optional["NullBitBufferGpu"]:
        """null_bit_buffer.rs:168-204: AND of both bitmaps; one-sided -> copy; none -> None"""
        if left is None and right is None:
            return None
        ref = left if left is not None else right
        if left is not None and right is not None:
            assert left.len == right.len
            assert left.gpu_device is right.gpu_device
        dev = ref.gpu_device
        out = dev.create_empty_buffer(bitmap_words(ref.len) * 4)
        check(lib().agpu_validity_and(dev.handle, left.bit_buffer.ptr if left else None,
                                      right.bit_buffer.ptr if right else None, out.ptr, ref.len),
              "agpu_validity_and")
        return NullBitBufferGpu(out, ref.len, dev)

    @staticmethod
    def merge_null_bit_buffer_op(left, right, pipeline) -> Optional["NullBitBufferGpu"]:
        return NullBitBufferGpu.merge_null_bit_buffer(left, right)


def _vptr(nb: Optional[NullBitBufferGpu]):
    return nb.bit_buffer.ptr if nb is not None else None


def _new_validity(dev: GpuDevice, length: int, *inputs: Optional[NullBitBufferGpu]):
    """Allocate the output bitmap of an op iff at least one input has one."""
    for x in inputs:
        if x is not None:
            return NullBitBufferGpu(dev.create_empty_buffer(((length + 31) >> 5) * 4), length, dev)
    return None


# ------------------------------------------------------------------------------------------
# arrays  (crates/array/src/array/primitive_array_gpu.rs, boolean_gpu.rs, *_gpu.rs)
# ------------------------------------------------------------------------------------------
class PrimitiveArrayGpu:
    """primitive_array_gpu.rs:12-19 — fields `data, gpu_device, len, null_buffer` as in the
    reference.  Concrete element types are the subclasses below (the reference's type aliases)."""

    DTYPE: int = -1                 # agpu dtype id
    NP: np.dtype = np.dtype("u1")   # numpy element type
    ITEMSIZE: int = 1               # bytes per row
    ARROW_TYPE: ArrowType

    def __init__(self, data: ArrowGpuBuffer, gpu_device: GpuDevice, length: int,
                 null_buffer: Optional[NullBitBufferGpu] = None):
        self._lazy = None
        self._data, self.gpu_device, self.len, self._null_buffer = data, gpu_device, length, null_buffer

    # `data` and `null_buffer` are the reference's public fields.  On a fusing pipeline an array may
    # be a recorded-but-not-yet-launched chain (`_lazy`); touching either field launches it.
    @property
    def data(self) -> ArrowGpuBuffer:
        if self._lazy is not None:
            self._materialize()
        return self._data

    @data.setter
    def data(self, value) -> None:
        self._data = value

    @property
    def null_buffer(self) -> Optional[NullBitBufferGpu]:
        if self._lazy is not None:
            self._materialize()
        return self._null_buffer

    @null_buffer.setter
    def null_buffer(self, value) -> None:
        self._null_buffer = value

    def _materialize(self) -> None:
        lazy, self._lazy = self._lazy, None
        out = lazy.evaluate()
        self._data, self._null_buffer = out._data, out._null_buffer

    # --- constructors
    @classmethod
    def from_slice(cls, value: Sequence, gpu_device: GpuDevice):
        arr = np.ascontiguousarray(np.asarray(value).astype(cls.NP, copy=False))
        return cls(gpu_device.create_gpu_buffer_with_data(arr), gpu_device, len(arr), None)

    @classmethod
    def from_optional_slice(cls, value: Sequence, gpu_device: GpuDevice):
        """nulls store T::default() (primitive_array_gpu.rs:39-41)"""
        valid = np.array([v is not None for v in value], dtype=bool)
        dense = np.array([0 if v is None else v for v in value]).astype(cls.NP) if len(value) else np.zeros(0, cls.NP)
        nb = NullBitBufferGpu.from_flags(gpu_device, valid)
        return cls(gpu_device.create_gpu_buffer_with_data(dense), gpu_device, len(value), nb)

    @classmethod
    def from_numpy(cls, values: np.ndarray, valid: Optional[np.ndarray], gpu_device: GpuDevice, wait: bool = True):
        """bulk constructor: dense values + optional bool validity flags"""
        nb = NullBitBufferGpu.from_flags(gpu_device, valid) if valid is not None else None
        arr = np.ascontiguousarray(values.astype(cls.NP, copy=False))
        return cls(gpu_device.create_gpu_buffer_with_data(arr, wait), gpu_device, len(arr), nb)

    @classmethod
    def empty(cls, length: int, gpu_device: GpuDevice, null_buffer=None):
        return cls(gpu_device.create_empty_buffer(length * cls.NP.itemsize), gpu_device, length, null_buffer)

    # --- readback
    def raw_values(self, out: Optional[np.ndarray] = None, wait: bool = True) -> np.ndarray:
        """primitive_array_gpu.rs:70-74; `out` = caller-provided (pinned) host array"""
        if out is not None:
            self.gpu_device.read_into(self.data, out[: self.len], wait)
            return out
        raw = self.gpu_device.retrive_data(self.data, self.len * self.NP.itemsize)
        return raw.view(self.NP)[: self.len]

    def values(self) -> list:
        raw = self.raw_values()
        if self.null_buffer is None:
            return [v.item() for v in raw]
        flags = self.null_buffer.flags()
        return [raw[i].item() if flags[i] else None for i in range(self.len)]

    def clone_array(self):
        return type(self)(self.gpu_device.clone_buffer(self.data), self.gpu_device, self.len,
                          NullBitBufferGpu.clone_null_bit_buffer(self.null_buffer))

    def get_gpu_device(self) -> GpuDevice:
        return self.gpu_device

    def get_dtype(self) -> ArrowType:
        return self.ARROW_TYPE

    def get_raw_values(self) -> np.ndarray:
        return self.raw_values()

    def __len__(self):
        return self.len

    # --- broadcast (array/src/kernels/broadcast.rs:6-17)
    @classmethod
    def broadcast(cls, value, length: int, gpu_device: GpuDevice):
        pipeline = ArrowComputePipeline(gpu_device, "broadcast")
        arr = cls.broadcast_op(value, length, pipeline)
        pipeline.finish()
        return arr

    @classmethod
    def broadcast_op(cls, value, length: int, pipeline: ArrowComputePipeline):
        dev = pipeline.device
        out = cls.empty(length, dev)
        scalar = np.asarray([value]).astype(cls.NP)
        check(lib().agpu_broadcast(dev.handle, cls.DTYPE, scalar.ctypes.data, out.data.ptr, length), "agpu_broadcast")
        return out

    def __repr__(self):
        return f"{type(self).__name__}(len={self.len}, values={self.values()[:16]}{'...' if self.len > 16 else ''})"


def _prim(name: str, dtype: int, np_dtype: str, arrow_type: ArrowType):
    return type(name, (PrimitiveArrayGpu,), {"DTYPE": dtype, "NP": np.dtype(np_dtype), "ARROW_TYPE": arrow_type,
                                             "ITEMSIZE": np.dtype(np_dtype).itemsize})


Float32ArrayGPU = _prim("Float32ArrayGPU", _ffi.F32, "<f4", ArrowType.Float32Type)
UInt32ArrayGPU = _prim("UInt32ArrayGPU", _ffi.U32, "<u4", ArrowType.UInt32Type)
UInt16ArrayGPU = _prim("UInt16ArrayGPU", _ffi.U16, "<u2", ArrowType.UInt16Type)
UInt8ArrayGPU = _prim("UInt8ArrayGPU", _ffi.U8, "u1", ArrowType.UInt8Type)
Int32ArrayGPU = _prim("Int32ArrayGPU", _ffi.I32, "<i4", ArrowType.Int32Type)
Int16ArrayGPU = _prim("Int16ArrayGPU", _ffi.I16, "<i2", ArrowType.Int16Type)
Int8ArrayGPU = _prim("Int8ArrayGPU", _ffi.I8, "i1", ArrowType.Int8Type)
Date32ArrayGPU = _prim("Date32ArrayGPU", _ffi.DATE32, "<i4", ArrowType.Date32Type)


class BooleanArrayGPU:
    """boolean_gpu.rs:15-21 — `data` is a packed LSB-first bitmap."""

    DTYPE = _ffi.BOOL
    ARROW_TYPE = ArrowType.BooleanType

    def __init__(self, data: ArrowGpuBuffer, gpu_device: GpuDevice, length: int,
                 null_buffer: Optional[NullBitBufferGpu] = None):
        self.data, self.gpu_device, self.len, self.null_buffer = data, gpu_device, length, null_buffer

    @classmethod
    def from_slice(cls, value: Sequence, gpu_device: GpuDevice):
        flags = np.asarray(value, dtype=bool)
        return cls(gpu_device.create_gpu_buffer_with_data(pack_bits(flags)), gpu_device, len(flags), None)

    @classmethod
    def from_optional_slice(cls, value: Sequence, gpu_device: GpuDevice):
        flags = np.array([bool(v) if v is not None else False for v in value], dtype=bool)
        valid = np.array([v is not None for v in value], dtype=bool)
        return cls(gpu_device.create_gpu_buffer_with_data(pack_bits(flags)), gpu_device, len(value),
                   NullBitBufferGpu.from_flags(gpu_device, valid))

    @classmethod
    def from_numpy(cls, flags: np.ndarray, valid: Optional[np.ndarray], gpu_device: GpuDevice):
        nb = NullBitBufferGpu.from_flags(gpu_device, valid) if valid is not None else None
        return cls(gpu_device.create_gpu_buffer_with_data(pack_bits(flags)), gpu_device, len(flags), nb)

    @classmethod
    def from_bytes_slice(cls, value: Sequence, gpu_device: GpuDevice, length: Optional[int] = None):
        """boolean_gpu.rs:72-82.  The reference sets len = number of BYTES (Q10); pass `length`
        for a bit count, the default keeps 8 * bytes bits."""
        raw = np.asarray(value, dtype=np.uint8)
        padded = np.zeros(_round_up(len(raw), 4), dtype=np.uint8)
        padded[: len(raw)] = raw
        return cls(gpu_device.create_gpu_buffer_with_data(padded), gpu_device,
                   len(raw) * 8 if length is None else length, None)

    @classmethod
    def empty(cls, length: int, gpu_device: GpuDevice, null_buffer=None):
        return cls(gpu_device.create_empty_buffer(bitmap_words(length) * 4), gpu_device, length, null_buffer)

    def raw_bytes(self) -> np.ndarray:
        return self.gpu_device.retrive_data(self.data, bitmap_words(self.len) * 4)

    def raw_values(self) -> np.ndarray:
        return unpack_bits(self.raw_bytes(), self.len)

    def values(self) -> list:
        raw = self.raw_values()
        if self.null_buffer is None:
            return [bool(v) for v in raw]
        flags = self.null_buffer.flags()
        return [bool(raw[i]) if flags[i] else None for i in range(self.len)]

    def get_gpu_device(self) -> GpuDevice:
        return self.gpu_device

    def get_dtype(self) -> ArrowType:
        return self.ARROW_TYPE

    def __len__(self):
        return self.len

    @classmethod
    def broadcast(cls, value: bool, length: int, gpu_device: GpuDevice):
        return cls.broadcast_op(value, length, ArrowComputePipeline(gpu_device, "broadcast"))

    @classmethod
    def broadcast_op(cls, value: bool, length: int, pipeline: ArrowComputePipeline):
        """boolean_gpu.rs:119-135: built on the host like the reference"""
        builder = BooleanBufferBuilder(length, set_all=bool(value))
        padded = np.zeros(bitmap_words(length) * 4, dtype=np.uint8)
        padded[: len(builder.data)] = builder.data
        return cls(pipeline.device.create_gpu_buffer_with_data(padded), pipeline.device, length, None)

    def __repr__(self):
        return f"BooleanArrayGPU(len={self.len}, values={self.values()[:16]}{'...' if self.len > 16 else ''})"


ARRAY_TYPES = {
    ArrowType.Float32Type: Float32ArrayGPU, ArrowType.UInt32Type: UInt32ArrayGPU,
    ArrowType.UInt16Type: UInt16ArrayGPU, ArrowType.UInt8Type: UInt8ArrayGPU,
    ArrowType.Int32Type: Int32ArrayGPU, ArrowType.Int16Type: Int16ArrayGPU,
    ArrowType.Int8Type: Int8ArrayGPU, ArrowType.Date32Type: Date32ArrayGPU,
    ArrowType.BooleanType: BooleanArrayGPU,
}
ARRAY_BY_NAME = {cls.__name__: cls for cls in ARRAY_TYPES.values()}

# `ArrowArrayGPU` is a Rust enum over the array types (array/mod.rs:104-114); in Python any of
# the classes above plays that role.
ArrowArrayGPU = (PrimitiveArrayGpu, BooleanArrayGPU)


class ScalarValue:
    """kernels/mod.rs:7-17: ScalarValue::F32(x) ... -> ScalarValue.F32(x)"""

    def __init__(self, cls, value):
        self.cls, self.value = cls, value

    F32 = classmethod(lambda c, x: c(Float32ArrayGPU, x))
    U32 = classmethod(lambda c, x: c(UInt32ArrayGPU, x))
    U16 = classmethod(lambda c, x: c(UInt16ArrayGPU, x))
    U8 = classmethod(lambda c, x: c(UInt8ArrayGPU, x))
    I32 = classmethod(lambda c, x: c(Int32ArrayGPU, x))
    I16 = classmethod(lambda c, x: c(Int16ArrayGPU, x))
    I8 = classmethod(lambda c, x: c(Int8ArrayGPU, x))
    BOOL = classmethod(lambda c, x: c(BooleanArrayGPU, x))


def broadcast_dyn(value: ScalarValue, length: int, device: GpuDevice):
    """array/mod.rs:181-192"""
    return value.cls.broadcast(value.value, length, device)


def broadcast_op_dyn(value: ScalarValue, length: int, pipeline: ArrowComputePipeline):
    """array/mod.rs:196-211"""
    return value.cls.broadcast_op(value.value, length, pipeline)
