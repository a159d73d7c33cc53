"""GPU: the C++ host-side mirror (arrow_gpu_b200/cpp/arrow_gpu.hpp) runs the transcribed
reference unit tests through the C ABI.  The binary links only libagpu.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "arrow_gpu_b200", "cpp")


@pytest.mark.gpu
def test_cpp_reference_cases():
    exe = os.path.join(CPP, "test_reference_cases")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", CPP], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(res.stdout, res.stderr)
    assert res.returncode == 0, res.stdout + res.stderr
    assert " 0 failures" in res.stdout


def test_cpp_mirror_links_only_the_c_abi():
    exe = os.path.join(CPP, "test_reference_cases")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", CPP], check=True)
    deps = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libagpu.so" in deps
    assert "oracle" not in deps and "torch" not in deps


@pytest.mark.gpu
def test_arrow_c_data_interface_roundtrip():
    """pyarrow -> Arrow C Data Interface -> C++ mirror (import, compute on the device, export) -> pyarrow"""
    import ctypes as C
    pa = pytest.importorskip("pyarrow")
    import pyarrow.compute as pc
    from pyarrow.cffi import ffi
    so = os.path.join(CPP, "libagpu_cdata.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", CPP], check=True)
    lib = C.CDLL(so)
    lib.agpu_cdata_apply.argtypes = [C.c_char_p] + [C.c_void_p] * 6
    lib.agpu_cdata_apply.restype = C.c_int

    def exported(arr):
        s, a = ffi.new("struct ArrowSchema*"), ffi.new("struct ArrowArray*")
        arr._export_to_c(int(ffi.cast("uintptr_t", a)), int(ffi.cast("uintptr_t", s)))
        return s, a

    def apply(op, x, y=None):
        sx, ax = exported(x)
        sy, ay = exported(y) if y is not None else (ffi.NULL, ffi.NULL)
        so_, ao = ffi.new("struct ArrowSchema*"), ffi.new("struct ArrowArray*")
        ptr = lambda p: int(ffi.cast("uintptr_t", p)) if p != ffi.NULL else None  # noqa: E731
        rc = lib.agpu_cdata_apply(op.encode(), ptr(sx), ptr(ax), ptr(sy), ptr(ay), ptr(so_), ptr(ao))
        assert rc == 0, (op, rc)
        for s, a in ((sx, ax), (sy, ay)):   # we are the consumer of what we exported: release it
            if a != ffi.NULL and a.release != ffi.NULL:
                a.release(a)
            if s != ffi.NULL and s.release != ffi.NULL:
                s.release(s)
        return pa.Array._import_from_c(ptr(ao), ptr(so_))

    cases = [pa.array([1, None, -3, 4, None, 100, -128, 127], pa.int8()), pa.array([1.5, None, 2.25, -0.0], pa.float32()),
             pa.array([True, None, False, True] * 9, pa.bool_()), pa.array(list(range(1000)), pa.uint16())[7:900],
             pa.array([None, 7, 19000], pa.date32()), pa.array([], pa.int32())]
    for a in cases:
        back = apply("identity", a)
        assert back.equals(pa.concat_arrays([a])), a.type
    x = pa.array([100, 127, -128, None, 5], pa.int8())
    y = pa.array([100, 1, -1, 3, None], pa.int8())
    assert apply("add", x, y).equals(pc.add(x, y))
    assert apply("gt", x, y).equals(pc.greater(x, y))
    f = pa.array([0.0, 1.0, 4.0, None, 9.0], pa.float32())
    assert apply("sqrt", f).equals(pc.sqrt(f))
    keep = pa.array([True, False, None, True, True])
    assert apply("filter", x, keep).equals(pc.filter(x, keep, null_selection_behavior="drop"))


@pytest.mark.gpu
def test_examples_run():
    """examples/scalar_ops.{py,cpp}: the usage styles of the reference's own example on both host mirrors"""
    import sys
    ex = os.path.join(ROOT, "examples")
    res = subprocess.run([sys.executable, os.path.join(ex, "scalar_ops.py")], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "scalar_ops ok: 2 kernels recorded one by one, 1 on a fusing pipeline" in res.stdout
    exe = os.path.join(ex, "scalar_ops")
    subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + CPP, os.path.join(ex, "scalar_ops.cpp"),
                    "-L" + os.path.join(ROOT, "arrow_gpu_b200", "lib"), "-lagpu",
                    "-Wl,-rpath," + os.path.join(ROOT, "arrow_gpu_b200", "lib"), "-o", exe], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "scalar_ops ok" in res.stdout, res.stdout + res.stderr
