"""arrow_gpu_b200 — B200-native (sm_100a) implementation of psvri/arrow-gpu's columnar compute
hot path behind the reference's own array/operator surface.

Layout:  csrc/  hand-written CUDA kernels + the C ABI (include/agpu.h) -> lib/libagpu.so
         array.py / kernels.py  host-side mirror of the reference's crates over that C ABI
         sharded.py  row-range sharding across the GPUs of one box (one process per GPU)
"""
from . import _ffi  # noqa: F401
from .array import *  # noqa: F401,F403
from .array import GPU_DEVICE, ARRAY_BY_NAME, ARRAY_TYPES  # noqa: F401
from .kernels import *  # noqa: F401,F403
from . import kernels  # noqa: F401
from .interop import from_arrow, to_arrow  # noqa: F401,E402
from .c_device import export_device, from_arrow_device  # noqa: F401,E402
