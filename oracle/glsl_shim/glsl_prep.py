#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY (oracle/) -- lexical GLSL -> C++ adapter for the reference's shaders.

The reference's GPU half (main.glsl, brdfs.glsl, progressive_rendering.glsl, temporal_reprojection.glsl
under project/addons/jar_path_tracing/src/shaders/) cannot be executed here (no Vulkan, no glslang), so
oracle/Makefile compiles the shader TEXT itself as C++ against oracle/glsl_shim/glsl.hpp.  A C++ compiler
cannot read six GLSL spellings; this script rewrites exactly those and nothing else.  Every rule is
lexical -- no expression, statement, constant or identifier of the algorithm is touched -- so whatever
the shader says is what runs.  Output goes to a scratch directory the Makefile deletes after compiling;
no reference source is stored in this repository.

Rules (applied to comment-stripped text; line structure is preserved so compiler messages cite shader lines):
  R1  `#[compute]` and `#version N` lines are dropped (not C preprocessor directives).
  R2  a floating literal without suffix gets `f`: GLSL literals are 32-bit floats, C++ ones are doubles.
  R3  interface blocks  `layout(..) [restrict] buffer Name { members } [instance];`
          with an instance name  ->  `struct Name { members } instance;`
          without                ->  the members themselves (GLSL puts them in global scope)
      and a member  `T name[];`  ->  `glsl::ssbo<T> name;`  (std430 runtime array, stride given at bind time).
  R4  `layout(.., FORMAT) [restrict] uniform [readonly|writeonly] image2D name;` -> `glsl::image2D<glsl::FORMAT> name;`
      `layout(..) uniform sampler2DArray name;`                                  -> `glsl::sampler2DArray name;`
  R5  `layout(local_size_x = ..) in;` is dropped (work-group shape; the bridge loops over pixels).
  R6  parameter qualifiers: `inout T x` / `in out T x` / `out T x` -> `T &x`;  `const in T x` -> `const T x`.
  R7  (only with --segments NAME) the literal bound of `for (int i = 0; i < 5; i++)` in path_trace
      (main.glsl:377) is replaced by the identifier NAME: BASELINE.json's configs ask for depth 4 and 8,
      the shader hard-codes 5.  NAME = 5 reproduces the text exactly; tests run 4, 5 and 8.
"""
import argparse
import os
import re
import sys

TOKEN = re.compile(r"""
    (?P<hex>0[xX][0-9a-fA-F]+[uU]?)
  | (?P<float>(?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?P<fsuf>[fF]|lf|LF)?
  | (?P<int>\d+[uU]?)
  | (?P<ident>[A-Za-z_]\w*)
  | (?P<other>.)
""", re.X | re.S)


def strip_comments(text):
    def blank(m):
        return re.sub(r"[^\n]", " ", m.group(0))
    text = re.sub(r"/\*.*?\*/", blank, text, flags=re.S)
    return re.sub(r"//[^\n]*", blank, text)


def float_suffix(text):  # R2
    out = []
    for m in TOKEN.finditer(text):
        if m.group("float") is not None and m.group("fsuf") is None:
            out.append(m.group("float") + "f")
        else:
            out.append(m.group(0))
    return "".join(out)


def keep_lines(original, replacement):
    """Pad the replacement with the newlines the original had, so line numbers stay those of the shader."""
    return replacement + "\n" * (original.count("\n") - replacement.count("\n"))


def interface_blocks(text):  # R3
    pat = re.compile(r"layout\s*\([^)]*\)\s*(?:restrict\s+)?buffer\s+(\w+)\s*\{([^}]*)\}\s*(\w*)\s*;")

    def repl(m):
        name, body, inst = m.group(1), m.group(2), m.group(3)
        body = re.sub(r"\b(\w+)\s+(\w+)\s*\[\s*\]\s*;", r"glsl::ssbo<\1> \2;", body)
        new = f"struct {name} {{{body}}} {inst};" if inst else body
        return keep_lines(m.group(0), new)
    return pat.sub(repl, text)


def resources(text):  # R4, R5
    text = re.sub(r"layout\s*\(\s*local_size_x[^)]*\)\s*in\s*;", "", text)
    img = re.compile(r"layout\s*\([^)]*?,\s*(\w+)\s*\)\s*(?:restrict\s+)?uniform\s+(?:(?:readonly|writeonly|restrict)\s+)*image2D\s+(\w+)\s*;")
    text = img.sub(lambda m: keep_lines(m.group(0), f"glsl::image2D<glsl::{m.group(1)}> {m.group(2)};"), text)
    smp = re.compile(r"layout\s*\([^)]*\)\s*uniform\s+sampler2DArray\s+(\w+)\s*;")
    return smp.sub(lambda m: keep_lines(m.group(0), f"glsl::sampler2DArray {m.group(1)};"), text)


def parameter_qualifiers(text):  # R6
    text = re.sub(r"\b(?:inout|in\s+out|out)\s+(\w+)\s+(\w+)", r"\1 &\2", text)
    return re.sub(r"\bconst\s+in\s+", "const ", text)


def segments(text, name):  # R7
    pat = re.compile(r"for\s*\(\s*int\s+i\s*=\s*0\s*;\s*i\s*<\s*5\s*;\s*i\+\+\s*\)")
    text, n = pat.subn(f"for (int i = 0; i < {name}; i++)", text)
    if n != 1:
        raise SystemExit(f"glsl_prep: expected exactly one path-segment loop, found {n}")
    return text


def convert(text, seg_name=None):
    text = strip_comments(text)
    text = re.sub(r"^\s*#\s*(\[compute\]|version\b)[^\n]*", "", text, flags=re.M)  # R1
    text = interface_blocks(text)
    text = resources(text)
    text = parameter_qualifiers(text)
    if seg_name:
        text = segments(text, seg_name)
    text = float_suffix(text)
    if re.search(r"\blayout\s*\(", text):
        raise SystemExit("glsl_prep: a layout(...) declaration was not recognised")
    return text


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("src_dir")
    ap.add_argument("out_dir")
    ap.add_argument("--segments", default=None)
    a = ap.parse_args()
    os.makedirs(a.out_dir, exist_ok=True)
    for name in sorted(os.listdir(a.src_dir)):
        if not name.endswith(".glsl"):
            continue
        with open(os.path.join(a.src_dir, name)) as f:
            text = f.read()
        with open(os.path.join(a.out_dir, name), "w") as f:
            f.write(convert(text, a.segments if name == "main.glsl" else None))
    return 0


if __name__ == "__main__":
    sys.exit(main())
