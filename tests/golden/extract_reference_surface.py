#!/usr/bin/env python3
"""Extract the public Rust surface of the reference's hot-path crates (traits with the arity of
every method, `*_dyn` / `*_op_dyn` functions with their arity, free `pub fn`s, public enums and
structs) into tests/golden/reference_surface.json.

    python tests/golden/extract_reference_surface.py /root/reference

The same parser (`surface_of_tree`) is applied to this repo's crates/ by tests/test_rust_surface.py,
which requires every reference item to be present with the same arity.  No Rust toolchain is
needed: it is a text-level check (brace matching + parameter counting)."""
from __future__ import annotations

import json
import os
import re
import sys

CRATES = ["array", "arithmetic", "compare", "logical", "cast", "math", "trigonometry", "routines"]


def strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def matching(text: str, start: int, open_ch: str, close_ch: str) -> int:
    """index of the bracket closing the one at `start`"""
    depth = 0
    for i in range(start, len(text)):
        c = text[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i
    return len(text) - 1


def arity(params: str) -> int:
    """number of parameters other than self"""
    parts, depth, cur = [], 0, ""
    for c in params:
        if c in "(<[{":
            depth += 1
        elif c in ")>]}":
            depth -= 1
        if c == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += c
    parts.append(cur)
    parts = [p.strip() for p in parts if p.strip()]
    return len([p for p in parts if not re.fullmatch(r"&?\s*(mut\s+)?self", p)])


def functions_in(body: str):
    """[(name, arity)] of every `fn name(...)` in a block (names may be macro metavariables)"""
    out = []
    for m in re.finditer(r"\bfn\s+(\$?\w+)\s*(<[^>(]*>)?\s*\(", body):
        end = matching(body, m.end() - 1, "(", ")")
        out.append((m.group(1), arity(body[m.end():end])))
    return out


def surface_of_text(text: str, surface: dict) -> None:
    text = strip_comments(text)
    # macro definitions: remember the arities of the `pub fn $x` they generate, then cut them out
    macros = {}
    for m in list(re.finditer(r"macro_rules!\s*(\w+)\s*\{", text)):
        end = matching(text, m.end() - 1, "{", "}")
        body = text[m.end():end]
        ar = sorted({a for n, a in functions_in(body) if n.startswith("$")})
        if ar:
            macros[m.group(1)] = ar
    cut = text
    for m in reversed(list(re.finditer(r"macro_rules!\s*(\w+)\s*\{", text))):
        end = matching(text, m.end() - 1, "{", "}")
        cut = cut[:m.start()] + cut[end + 1:]
    # macro invocations that generate *_dyn functions: `_op_dyn` names take the pipeline as well
    for m in re.finditer(r"\b(\w+)!\s*\(", cut):
        if m.group(1) not in macros:
            continue
        end = matching(cut, m.end() - 1, "(", ")")
        lo, hi = min(macros[m.group(1)]), max(macros[m.group(1)])
        for name in re.findall(r"\b(\w+_dyn)\b", cut[m.end():end]):
            surface["dyn_functions"][name] = hi if name.endswith("_op_dyn") else lo
    # traits
    for m in re.finditer(r"\bpub\s+trait\s+(\w+)", cut):
        brace = cut.find("{", m.end())
        semi = cut.find(";", m.end())
        if brace < 0 or (0 <= semi < brace):
            continue
        end = matching(cut, brace, "{", "}")
        methods = {n: a for n, a in functions_in(cut[brace:end]) if not n.startswith("$")}
        surface["traits"].setdefault(m.group(1), {}).update(methods)
    # free functions
    for m in re.finditer(r"^pub\s+fn\s+(\w+)\s*(<[^>(]*>)?\s*\(", cut, flags=re.M):
        end = matching(cut, m.end() - 1, "(", ")")
        key = "dyn_functions" if m.group(1).endswith("_dyn") else "functions"
        surface[key][m.group(1)] = arity(cut[m.end():end])
    for m in re.finditer(r"^\s*pub\s+(enum|struct)\s+(\w+)", cut, flags=re.M):
        surface["types"].setdefault(m.group(2), m.group(1))


def surface_of_tree(crates_root: str) -> dict:
    surface = {"traits": {}, "dyn_functions": {}, "functions": {}, "types": {}}
    for crate in CRATES:
        src = os.path.join(crates_root, crate, "src")
        for dirpath, _dirs, files in sorted(os.walk(src)):
            for fn in sorted(files):
                if fn.endswith(".rs"):
                    surface_of_text(open(os.path.join(dirpath, fn)).read(), surface)
    return surface


# helper traits whose only content was the WGSL text of the wgpu path (`const SHADER`, `create_new`)
# keep their NAMES here as marker traits; their methods are implementation details of that path
SHADER_HELPERS = {"NegUnaryType", "Sum32Bit", "LogicalType", "CompareType", "SwizzleType", "MathUnaryType", "MathBinaryType",
                  "FloatMathUnaryType", "HyperbolicType", "TrigonometricType"}


def main() -> None:
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    surface = surface_of_tree(os.path.join(ref, "crates"))
    for t in SHADER_HELPERS:
        if t in surface["traits"]:
            surface["traits"][t] = {}
    # crate-private trait of the array crate (`pub(crate) trait ArrowArray`) is not public surface
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_surface.json")
    with open(out, "w") as f:
        json.dump(surface, f, indent=1, sort_keys=True)
        f.write("\n")
    print(f"wrote {out}: {len(surface['traits'])} traits, {len(surface['dyn_functions'])} dyn functions, "
          f"{len(surface['functions'])} functions, {len(surface['types'])} types")


if __name__ == "__main__":
    main()
