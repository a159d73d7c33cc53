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
