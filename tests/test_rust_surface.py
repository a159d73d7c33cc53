"""CPU: the Rust drop-in tree (crates/) keeps the reference's public surface.

There is no cargo/rustc in the build image, so crates/ cannot be compiled here.  It is kept honest
at text level: every public trait (with each method's arity), every `*_dyn` / `*_op_dyn`
function and every public type of the reference's hot-path crates — recorded in
tests/golden/reference_surface.json by tests/golden/extract_reference_surface.py — must appear in
crates/ with the same arity, and every `agpu_*` function the Rust files call must be declared in
include/agpu_ffi.rs with the number of arguments used."""
import importlib.util
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# items of the reference that ARE the wgpu seam this build replaces (they name wgpu types or carry
# WGSL text): not part of the surface a caller of the operator crates sees
WGPU_SEAM_FUNCTIONS = {"apply_boolean_unary_function"}
WGPU_SEAM_ARITY_DIFFERS = {"merge_null_buffers_op"}   # takes `&wgpu::Buffer`s there, NullBitBufferGpu options here
WGPU_SEAM_TYPES = {"CmpQuery"}                         # timestamp queries -> agpu_event_* (include/agpu.h)


def _extractor():
    spec = importlib.util.spec_from_file_location("extract_reference_surface", os.path.join(GOLDEN, "extract_reference_surface.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def surfaces():
    ex = _extractor()
    want = json.load(open(os.path.join(GOLDEN, "reference_surface.json")))
    have = ex.surface_of_tree(os.path.join(ROOT, "crates"))
    return ex, want, have


def test_fixture_is_current_when_the_reference_is_present(surfaces):
    ex, want, _have = surfaces
    ref = "/root/reference/crates"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present (GPU box)")
    fresh = ex.surface_of_tree(ref)
    for t in ex.SHADER_HELPERS:
        if t in fresh["traits"]:
            fresh["traits"][t] = {}
    assert fresh == want, "run python tests/golden/extract_reference_surface.py /root/reference"


def test_every_reference_trait_and_method_is_present(surfaces):
    _ex, want, have = surfaces
    assert len(want["traits"]) >= 40
    for trait, methods in want["traits"].items():
        assert trait in have["traits"], f"trait {trait} missing from crates/"
        for name, n_args in methods.items():
            assert name in have["traits"][trait], f"{trait}::{name} missing"
            assert have["traits"][trait][name] == n_args, f"{trait}::{name}: arity {have['traits'][trait][name]} != {n_args}"


def test_every_dyn_function_is_present_with_the_same_arity(surfaces):
    _ex, want, have = surfaces
    assert len(want["dyn_functions"]) >= 90
    for name, n_args in want["dyn_functions"].items():
        assert name in have["dyn_functions"], f"{name} missing from crates/"
        assert have["dyn_functions"][name] == n_args, f"{name}: arity {have['dyn_functions'][name]} != {n_args}"
    for name, n_args in want["functions"].items():
        if name in WGPU_SEAM_FUNCTIONS:
            continue
        assert name in have["functions"], f"pub fn {name} missing"
        if name not in WGPU_SEAM_ARITY_DIFFERS:
            assert have["functions"][name] == n_args, name


def test_every_public_type_is_present(surfaces):
    _ex, want, have = surfaces
    for name, kind in want["types"].items():
        if name in WGPU_SEAM_TYPES:
            continue
        assert have["types"].get(name) == kind, f"pub {kind} {name} missing"
    text = open(os.path.join(ROOT, "crates", "array", "src", "array", "primitive_array_gpu.rs")).read()
    for field in ("pub data: ArrowGpuBuffer", "pub gpu_device: Arc<GpuDevice>", "pub phantom: PhantomData<T>", "pub len: usize",
                  "pub null_buffer: Option<NullBitBufferGpu>"):   # public fields of the reference (primitive_array_gpu.rs:12-19)
        assert field in text, field


def test_ffi_calls_match_the_generated_declarations():
    ffi = open(os.path.join(ROOT, "include", "agpu_ffi.rs")).read()
    declared = {m.group(1): len([a for a in m.group(2).split(",") if a.strip()])
                for m in re.finditer(r"pub fn (agpu_\w+)\((.*?)\)(?: ->|;)", ffi)}
    ex = _extractor()
    used = {}
    for dirpath, _d, files in os.walk(os.path.join(ROOT, "crates")):
        for fn in files:
            if not fn.endswith(".rs"):
                continue
            text = ex.strip_comments(open(os.path.join(dirpath, fn)).read())
            for m in re.finditer(r"\b(agpu_\w+)\s*\(", text):
                end = ex.matching(text, m.end() - 1, "(", ")")
                used.setdefault(m.group(1), set()).add((ex.arity(text[m.end():end]), os.path.relpath(os.path.join(dirpath, fn), ROOT)))
    assert len(used) >= 25, sorted(used)
    for name, calls in used.items():
        assert name in declared, f"{name} is called in crates/ but not declared in include/agpu_ffi.rs"
        for n_args, where in calls:
            assert n_args == declared[name], f"{where}: {name} called with {n_args} arguments, declared with {declared[name]}"
