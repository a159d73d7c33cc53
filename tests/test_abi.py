"""CPU: the C-ABI library builds, loads, exports every symbol include/agpu.h declares, keeps the
ids of the header, fails loudly without a GPU — and the product never touches the oracle."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ffi():
    from arrow_gpu_b200 import _ffi
    if not os.path.exists(_ffi.LIB_PATH):
        _ffi.build()
    return _ffi


def test_library_exports_every_header_symbol(ffi):
    lib = ffi.lib()
    declared = ffi.header_functions()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/agpu.h but not exported by libagpu.so"
    # and every declared function has a ctypes signature (so tests call it with checked types)
    assert sorted(ffi.SIGNATURES) == declared
    assert lib.agpu_abi_version() == 2


def test_enum_ids_match_header(ffi):
    text = open(ffi.HEADER).read()
    found = re.findall(r"\b(AGPU_[A-Z0-9_]+)\s*=\s*(\d+)", text)
    assert len(found) >= 45
    for name, value in found:
        short = name[len("AGPU_"):]
        assert getattr(ffi, short) == int(value), name
    import oracle
    for short in ("BOOL I8 I16 I32 U8 U16 U32 F32 DATE32 ADD SUB MUL DIV REM MIN MAX AND OR XOR POW NEG ABS NOT "
                  "SQRT CBRT EXP EXP2 LOG LOG2 SIN COS ACOS SINH GT GTEQ LT LTEQ EQ SHL SHR").split():
        assert getattr(oracle, short) == getattr(ffi, short), short


def test_only_sm100a_code_in_library(ffi):
    out = subprocess.run(["cuobjdump", "--list-elf", ffi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device(ffi):
    """On a box without a GPU every compute entry point must fail, never compute on the host."""
    n = C.c_int(0)
    ffi.lib().agpu_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is visible here")
    h = C.c_void_p()
    assert ffi.lib().agpu_device_create(0, C.byref(h)) == -3  # AGPU_ENODEVICE
    assert ffi.lib().agpu_binary(None, 0, ffi.F32, None, None, None, 16, None, None, None) == -3
    assert ffi.lib().agpu_compare(None, 0, ffi.F32, None, None, None, 16, None, None, None) == -3
    import arrow_gpu_b200 as ag
    with pytest.raises(ag._ffi.AgpuError):
        ag.GpuDevice(0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under arrow_gpu_b200/ may reference it."""
    pkg = os.path.join(ROOT, "arrow_gpu_b200")
    for dirpath, _d, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, fn)).read()
                code = "\n".join(l for l in text.splitlines() if not l.strip().startswith(("#", "//", "*", "/*")))
                assert not re.search(r"^\s*(import|from)\s+oracle|liboracle|oracle_[a-z]+\s*\(", code, re.M), fn
    code = "import sys; import arrow_gpu_b200; assert 'oracle' not in sys.modules"
    subprocess.run([sys.executable, "-c", code], cwd=ROOT, check=True)
    deps = subprocess.run(["ldd", os.path.join(pkg, "lib", "libagpu.so")], capture_output=True, text=True).stdout
    assert "oracle" not in deps


def test_rust_ffi_module_matches_header(ffi):
    """include/agpu_ffi.rs (the extern "C" module of INTEGRATION.md) is generated from the header:
    it must be current and declare every function with the argument count ctypes binds"""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(root, "include", "gen_rust_ffi.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    committed = open(os.path.join(root, "include", "agpu_ffi.rs")).read()
    assert committed == gen.generate(), "run python include/gen_rust_ffi.py"
    declared = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (agpu_\w+)\((.*?)\)(?: ->|;)", committed)}
    assert set(declared) == set(ffi.SIGNATURES)
    for name, (_ret, args) in ffi.SIGNATURES.items():
        n_args = len([a for a in declared[name].split(",") if a.strip()])
        assert n_args == len(args), name
