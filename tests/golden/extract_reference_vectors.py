#!/usr/bin/env python3
"""Extract the golden vectors held by the reference's own unit tests.

The reference (psvri/arrow-gpu, mounted read-only at /root/reference) keeps all of
its known-answer tests as `macro_rules!` invocations with literal Rust arrays
(SURVEY.md section 4):

    test_unary_op! / test_unary_op_float!   crates/test_macros/src/lib.rs:2,158
    test_scalar_op! / test_float_scalar_op! crates/test_macros/src/lib.rs:32,120
    test_array_op! / test_float_array_op!   crates/test_macros/src/lib.rs:55,197
    test_cast_op! / test_bitcast_op!        crates/cast/src/lib.rs:196,220
    test_broadcast!                         crates/array/src/array/primitive_array_gpu.rs:143
    test_sum!                               crates/arithmetic/src/lib.rs:100
    test_merge_op!                          crates/routines/src/merge.rs:148
    test_take_op!                           crates/routines/src/take.rs:100
    test_put_op!                            crates/routines/src/put.rs:113

This script parses every invocation, evaluates the Rust literal expressions with a
small recursive-descent evaluator (no Rust toolchain exists in this image) and writes
`reference_vectors.json` beside itself.  The reference tree does not travel to the GPU
box, so the JSON is committed; re-run this script only when the reference changes.

    python tests/golden/extract_reference_vectors.py [/root/reference]

Nothing is copied from the reference except the literal test data (inputs and
expected outputs), which is exactly what a golden-vector fixture is.
"""
from __future__ import annotations

import json
import math
import os
import re
import sys

import numpy as np

MACROS = {
    "test_unary_op", "test_unary_op_float", "test_scalar_op", "test_float_scalar_op",
    "test_array_op", "test_float_array_op", "test_cast_op", "test_bitcast_op",
    "test_broadcast", "test_sum", "test_merge_op", "test_take_op", "test_put_op",
}

INT_RANGES = {
    "i8": (-(1 << 7), (1 << 7) - 1), "i16": (-(1 << 15), (1 << 15) - 1),
    "i32": (-(1 << 31), (1 << 31) - 1), "i64": (-(1 << 63), (1 << 63) - 1),
    "u8": (0, (1 << 8) - 1), "u16": (0, (1 << 16) - 1), "u32": (0, (1 << 32) - 1),
    "u64": (0, (1 << 64) - 1), "usize": (0, (1 << 64) - 1),
}


# --------------------------------------------------------------------------- tokenizer
TOKEN_RE = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<num>0b[01_]+|0x[0-9a-fA-F_]+|\d[\d_]*(?:\.\d[\d_]*)?(?:[eE][+-]?\d+)?)
  | (?P<ident>[A-Za-z_][A-Za-z_0-9]*!?)
  | (?P<op>::|<<|>>|[\[\]\(\),;\.\-\+\*/%&|^!])
""", re.X | re.S)


def tokenize(src: str):
    pos, out = 0, []
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at {src[pos:pos + 30]!r}")
        pos = m.end()
        if m.lastgroup == "ws":
            continue
        out.append((m.lastgroup, m.group()))
    return out


# --------------------------------------------------------------------------- evaluator
class F32(float):
    """Marks a value as an f32 (so method calls round like Rust's f32 would)."""


def f32(x) -> F32:
    return F32(float(np.float32(x)))


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else (None, None)

    def eat(self, val=None):
        kind, tok = self.peek()
        if val is not None and tok != val:
            raise SyntaxError(f"expected {val!r} got {tok!r}")
        self.i += 1
        return tok

    # precedence climbing: (| ) < (&) < (<< >>) < (+ -) < (* / %) < as < unary < postfix
    def expr(self):
        return self.bin_or()

    # Rust precedence: | lowest, then ^, then &
    def bin_or(self):
        v = self.bin_xor_lvl()
        while self.peek()[1] == "|":
            self.eat()
            v = v | self.bin_xor_lvl()
        return v

    def bin_xor_lvl(self):
        v = self.bin_and()
        while self.peek()[1] == "^":
            self.eat()
            v = v ^ self.bin_and()
        return v

    def bin_and(self):
        v = self.shift()
        while self.peek()[1] == "&":
            self.eat()
            v = v & self.shift()
        return v

    def shift(self):
        v = self.add()
        while self.peek()[1] in ("<<", ">>"):
            op = self.eat()
            r = self.add()
            v = v << r if op == "<<" else v >> r
        return v

    def add(self):
        v = self.mul()
        while self.peek()[1] in ("+", "-"):
            op = self.eat()
            r = self.mul()
            v = self._arith(v, r, op)
        return v

    def mul(self):
        v = self.cast()
        while self.peek()[1] in ("*", "/", "%"):
            op = self.eat()
            r = self.cast()
            v = self._arith(v, r, op)
        return v

    @staticmethod
    def _arith(a, b, op):
        isf = isinstance(a, float) or isinstance(b, float)
        if op == "+":
            r = a + b
        elif op == "-":
            r = a - b
        elif op == "*":
            r = a * b
        elif op == "/":
            r = a / b if isf else int(a / b)  # Rust int division truncates
        else:
            r = math.fmod(a, b) if isf else int(math.fmod(a, b))
        if isinstance(a, F32) or isinstance(b, F32):
            return f32(r)
        return r

    def cast(self):
        v = self.unary()
        while self.peek()[1] == "as":
            self.eat()
            ty = self.eat()
            v = self._as(v, ty)
        return v

    @staticmethod
    def _as(v, ty):
        if ty in ("f32",):
            return f32(v)
        if ty == "f64":
            return float(v)
        lo, hi = INT_RANGES[ty]
        if isinstance(v, float):  # Rust float->int `as` saturates, NaN -> 0
            if math.isnan(v):
                return 0
            return int(min(max(math.trunc(v), lo), hi))
        span = hi - lo + 1
        return (int(v) - lo) % span + lo  # int->int `as` wraps

    def unary(self):
        if self.peek()[1] == "-":
            self.eat()
            v = self.unary()
            return f32(-v) if isinstance(v, F32) else -v
        if self.peek()[1] == "!":
            # bitwise not on an integer whose width comes from context: evaluate in
            # unbounded two's complement, `wrap_to_type` narrows it afterwards
            self.eat()
            v = self.unary()
            return (not v) if isinstance(v, bool) else ~v
        return self.postfix()

    def postfix(self):
        v = self.atom()
        while self.peek()[1] == ".":
            self.eat()
            name = self.eat()
            self.eat("(")
            args = []
            while self.peek()[1] != ")":
                args.append(self.expr())
                if self.peek()[1] == ",":
                    self.eat()
            self.eat(")")
            v = self._method(v, name, args)
        return v

    @staticmethod
    def _method(v, name, args):
        if name in ("to_vec", "clone", "into"):
            return v
        x = np.float32(v)
        with np.errstate(all="ignore"):
            table = {
                "sin": np.sin, "cos": np.cos, "acos": np.arccos, "sinh": np.sinh,
                "sqrt": np.sqrt, "cbrt": np.cbrt, "exp": np.exp, "exp2": np.exp2,
                "ln": np.log, "log2": np.log2, "abs": np.abs,
            }
            if name in table:
                return f32(table[name](x))
            if name == "powf":
                return f32(np.power(x, np.float32(args[0])))
            if name == "powi":
                return f32(np.power(x, np.float32(args[0])))
        raise SyntaxError(f"unknown method .{name}()")

    def atom(self):
        kind, tok = self.peek()
        if tok == "(":
            self.eat()
            v = self.expr()
            self.eat(")")
            return v
        if tok == "[":
            return self.array()
        if kind == "num":
            self.eat()
            return self.number(tok)
        if kind == "ident":
            self.eat()
            if tok == "vec!":
                return self.array()
            if tok == "true":
                return True
            if tok == "false":
                return False
            if tok == "None":
                return None
            if tok == "Some":
                self.eat("(")
                v = self.expr()
                self.eat(")")
                return v
            if self.peek()[1] == "::":
                self.eat()
                name = self.eat()
                if name == "from_bits":  # f32::from_bits(u32) keeps the exact pattern
                    self.eat("(")
                    bits = self.expr()
                    self.eat(")")
                    return F32Bits(int(bits) & 0xFFFFFFFF)
                return self.constant(tok, name)
            return Ident(tok)
        raise SyntaxError(f"unexpected token {tok!r}")

    def number(self, tok):
        # a type suffix is tokenized as a following identifier glued to the number:
        # "100i32" -> num "100", ident "i32";  "1.0f32" -> num "1.0", ident "f32"
        clean = tok.replace("_", "")
        suffix = None
        k, nxt = self.peek()
        if k == "ident" and nxt in (*INT_RANGES, "f32", "f64"):
            suffix = self.eat()
        if clean.startswith("0b"):
            v = int(clean[2:], 2)
        elif clean.startswith("0x"):
            v = int(clean[2:], 16)
        elif any(c in clean for c in ".eE"):
            v = float(clean)
        else:
            v = int(clean)
        if suffix == "f32":
            return f32(v)
        if suffix == "f64":
            return float(v)
        return v

    @staticmethod
    def constant(ty, name):
        if ty in ("f32", "f64"):
            val = {"NAN": math.nan, "INFINITY": math.inf, "NEG_INFINITY": -math.inf,
                   "MAX": float(np.finfo(np.float32).max), "MIN": float(np.finfo(np.float32).min),
                   "EPSILON": float(np.finfo(np.float32).eps)}[name]
            return f32(val)
        lo, hi = INT_RANGES[ty]
        return {"MAX": hi, "MIN": lo}[name]

    def array(self):
        self.eat("[")
        items = []
        if self.peek()[1] == "]":
            self.eat()
            return items
        first = self.expr()
        if self.peek()[1] == ";":  # [value; count]
            self.eat()
            count = self.expr()
            self.eat("]")
            return [first] * int(count)
        items.append(first)
        while self.peek()[1] == ",":
            self.eat()
            if self.peek()[1] == "]":
                break
            items.append(self.expr())
        self.eat("]")
        return items


class Ident(str):
    pass


class F32Bits(int):
    """An f32 given by its bit pattern (NaN payloads must survive the JSON)."""


def evaluate(src: str):
    p = Parser(tokenize(src))
    v = p.expr()
    if p.i != len(p.t):
        raise SyntaxError(f"trailing tokens in {src!r}")
    return v


# --------------------------------------------------------------------------- macro scan
def split_top_level(body: str):
    """Split macro arguments on top-level commas."""
    args, depth, cur, i = [], 0, [], 0
    while i < len(body):
        c = body[i]
        if body.startswith("//", i):
            j = body.find("\n", i)
            i = len(body) if j < 0 else j
            continue
        if body.startswith("/*", i):
            i = body.find("*/", i) + 2
            continue
        if c == '"':
            j = body.find('"', i + 1)
            cur.append(body[i:j + 1])
            i = j + 1
            continue
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            args.append("".join(cur).strip())
            cur = []
        else:
            cur.append(c)
        i += 1
    tail = "".join(cur).strip()
    if tail:
        args.append(tail)
    return args


def strip_block_comments(text: str) -> str:
    """Blank out /* ... */ regions (the reference keeps disabled tests in them) but keep
    newlines so that line numbers survive."""
    def blank(m):
        return re.sub(r"[^\n]", " ", m.group())
    return re.sub(r"/\*.*?\*/", blank, text, flags=re.S)


def find_invocations(text: str):
    text = strip_block_comments(text)
    for m in re.finditer(r"\b(test_[a-z_0-9]+)!\s*\(", text):
        name = m.group(1)
        if name not in MACROS:
            continue
        line_start = text.rfind("\n", 0, m.start()) + 1
        prefix = text[line_start:m.start()]
        if "macro_rules" in prefix or prefix.strip().startswith("//"):
            continue
        depth, i = 1, m.end()
        while depth:
            c = text[i]
            if c == "(":
                depth += 1
            elif c == ")":
                depth -= 1
            i += 1
        yield name, text.count("\n", 0, m.start()) + 1, text[m.end():i - 1]


def jsonable(v):
    if isinstance(v, list):
        return [jsonable(x) for x in v]
    if isinstance(v, bool) or v is None:
        return v
    if isinstance(v, F32Bits):
        return {"f32_bits": int(v)}
    if isinstance(v, float):
        if math.isnan(v):
            return "NaN"
        if math.isinf(v):
            return "Infinity" if v > 0 else "-Infinity"
        return float(v)
    if isinstance(v, str):
        return str(v)
    return int(v)


NATIVE = {
    "Float32ArrayGPU": "f32", "Int32ArrayGPU": "i32", "UInt32ArrayGPU": "u32",
    "Int16ArrayGPU": "i16", "UInt16ArrayGPU": "u16", "Int8ArrayGPU": "i8",
    "UInt8ArrayGPU": "u8", "Date32ArrayGPU": "i32", "BooleanArrayGPU": "bool",
}


def wrap_to_type(v, array_type):
    """Narrow untyped integer literals to the element type of the array they feed
    (what rustc's inference does for `[!0, 256 * 256 * -5, ...]`)."""
    ty = NATIVE.get(array_type)
    if isinstance(v, list):
        return [wrap_to_type(x, array_type) for x in v]
    if ty is None or v is None or isinstance(v, (bool, F32Bits, str)):
        return v
    if ty == "f32":
        return f32(v)
    if ty == "bool":
        return v
    if isinstance(v, float):
        return v
    lo, hi = INT_RANGES[ty]
    return (int(v) - lo) % (hi - lo + 1) + lo


# which array type gives each data field its element type
FIELD_TYPES = {
    "input": "input_type", "scalar": "scalar_type", "lhs": "lhs_type", "rhs": "rhs_type",
    "expected": "output_type", "value": "output_type", "base": "input_type",
    "src": "array_type", "dst": "array_type",
}


def parse_case(macro, args):
    attrs = []
    while args and args[0].lstrip().startswith("#["):
        # an attribute is glued to the first real argument: "#[...] test_name"
        m = re.match(r"\s*(#\[(?:[^\[\]]|\[[^\]]*\])*\])\s*(.*)", args[0], re.S)
        if not m:
            break
        attrs.append(" ".join(m.group(1).split()))
        rest = m.group(2).strip()
        if rest:
            args[0] = rest
        else:
            args.pop(0)
    ev = evaluate
    n = len(args)
    c = {"name": args[0], "ignored_in_reference_ci": attrs}
    if macro in ("test_unary_op", "test_unary_op_float"):
        c.update(input_type=args[1], output_type=args[2], input=ev(args[3]), op=args[4])
        if n == 7:
            c.update(op_dyn=args[5], expected=ev(args[6]))
        else:
            c.update(op_dyn=None, expected=ev(args[5]))
    elif macro in ("test_scalar_op", "test_float_scalar_op"):
        c.update(input_type=args[1], scalar_type=args[2], output_type=args[3], input=ev(args[4]),
                 op=args[5], op_dyn=args[6], scalar=ev(args[7]), expected=ev(args[8]))
    elif macro in ("test_array_op", "test_float_array_op", "test_take_op"):
        c.update(lhs_type=args[1], rhs_type=args[2], output_type=args[3], op=args[4])
        if n == 9:
            c.update(op_dyn=args[5], lhs=ev(args[6]), rhs=ev(args[7]), expected=ev(args[8]))
        else:
            c.update(op_dyn=None, lhs=ev(args[5]), rhs=ev(args[6]), expected=ev(args[7]))
        if macro == "test_take_op":
            # 8-arg arm builds the source with from_slice, the 9-arg arm with
            # from_optional_slice (routines/src/take.rs:102-129)
            c["lhs_optional"] = n == 9
    elif macro in ("test_cast_op", "test_bitcast_op"):
        c.update(input_type=args[1], output_type=args[2], input=ev(args[3]),
                 cast_type=args[4], expected=ev(args[5]))
    elif macro == "test_broadcast":
        c.update(output_type=args[1], value=ev(args[2]), length=100)
    elif macro == "test_sum":
        c.update(input_type=args[1], base=ev(args[2]), size=ev(args[3]), expected=ev(args[4]))
    elif macro == "test_merge_op":
        c.update(lhs_type=args[1], rhs_type=args[2], output_type=args[3], op=args[4])
        if n == 10:
            c.update(op_dyn=args[5], lhs=ev(args[6]), rhs=ev(args[7]), mask=ev(args[8]),
                     expected=ev(args[9]))
        else:
            c.update(op_dyn=None, lhs=ev(args[5]), rhs=ev(args[6]), mask=ev(args[7]),
                     expected=ev(args[8]))
    elif macro == "test_put_op":
        c.update(array_type=args[1], op=args[2])
        if n == 9:
            c.update(op_dyn=args[3], src=ev(args[4]), dst=ev(args[5]), src_indexes=ev(args[6]),
                     dst_indexes=ev(args[7]), expected=ev(args[8]))
        else:
            c.update(op_dyn=None, src=ev(args[3]), dst=ev(args[4]), src_indexes=ev(args[5]),
                     dst_indexes=ev(args[6]), expected=ev(args[7]))
    for field, tyfield in FIELD_TYPES.items():
        if field in c:
            aty = c.get(tyfield)
            if field == "expected":
                aty = c.get("output_type") or c.get("array_type") or c.get("input_type")
            c[field] = wrap_to_type(c[field], aty)
    for field in ("src_indexes", "dst_indexes"):
        if field in c:
            c[field] = wrap_to_type(c[field], "UInt32ArrayGPU")
    if "mask" in c:
        c["mask"] = wrap_to_type(c["mask"], "BooleanArrayGPU")
    return {k: jsonable(v) for k, v in c.items()}


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    cases = []
    for dirpath, _dirs, files in sorted(os.walk(os.path.join(root, "crates"))):
        for fn in sorted(files):
            if not fn.endswith(".rs"):
                continue
            path = os.path.join(dirpath, fn)
            rel = os.path.relpath(path, root)
            text = open(path, encoding="utf-8").read()
            for macro, line, body in find_invocations(text):
                case = parse_case(macro, split_top_level(body))
                case["macro"] = macro
                case["source"] = f"{rel}:{line}"
                cases.append(case)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(out, "w") as f:
        json.dump({"reference": "psvri/arrow-gpu", "generator": "tests/golden/extract_reference_vectors.py",
                   "float_tolerance": "abs 0.01 on |x| (crates/test_macros/src/lib.rs:88-109)",
                   "cases": cases}, f, indent=1)
    by = {}
    for c in cases:
        by[c["macro"]] = by.get(c["macro"], 0) + 1
    print(f"wrote {len(cases)} cases to {out}: {by}")


if __name__ == "__main__":
    main()
