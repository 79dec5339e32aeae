#!/usr/bin/env python3
"""Rewrite the reference's GLSL compute-shader text into something a C++ compiler accepts.

TEST INFRASTRUCTURE ONLY.  Reads resources/*.comp and *.glsl where they lie under the
reference checkout and writes the rewritten text to oracle/_ref/gen/ (git-ignored; the
reference's sources are never copied into the repository).  The rewrite is purely
syntactic -- the algorithm text is untouched:

  * `#version`, `layout(local_size...) in;` lines are dropped;
  * interface blocks  `layout(...) [qualifiers] uniform|buffer Name {` -> `struct Name {`
  * opaque uniforms   `layout(...) [qualifiers] uniform T name;`       -> `T name;`
  * runtime array     `Node nodes[];`                                   -> `const Node* nodes;`
  * `out T name` parameters                                             -> `T& name`
  * multi-component swizzles `.xyz` -> `.xyz()` (single components stay fields)
  * `void main()` -> `void shader_main()`
  * `#include "x.glsl"` -> `#include "x.glsl.inc"`
  * the one semantic patch: `*_stack[scale]` -> `*_stack[xn_guard(scale)]`, because
    esvo.comp:119-123 reads the stack with an underflowed index right before its loop
    exits (undefined in GLSL, a segfault in C++; the value is never used).
"""
import os
import re
import sys

SWIZZLES = r"\.(xyz|yzx|zxy|rgb|xx|yz|xy)\b(?!\()"
QUALS = r"(?:(?:readonly|writeonly|restrict|coherent)\s+)*"


def rewrite(text: str) -> str:
    text = re.sub(r"^\s*#version.*$", "", text, flags=re.M)
    text = re.sub(r"^\s*layout\([^)]*\)\s*in\s*;\s*$", "", text, flags=re.M)
    text = re.sub(r"layout\([^)]*\)\s*" + QUALS + r"(?:uniform|buffer)\s+(\w+)\s*\{", r"struct \1 {", text)
    text = re.sub(r"layout\([^)]*\)\s*" + QUALS + r"uniform\s+(\w+)\s+(\w+)\s*;", r"\1 \2;", text)
    text = re.sub(r"\bNode\s+nodes\[\]\s*;", "const Node* nodes;", text)
    text = re.sub(r"\bout\s+(\w+)\s+(\w+)", r"\1& \2", text)
    text = re.sub(SWIZZLES, r".\1()", text)
    text = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", text)
    text = re.sub(r'#include\s+"(\w+\.glsl)"', r'#include "\1.inc"', text)
    text = re.sub(r"(\w+_stack)\[scale\]", r"\1[xn_guard(scale)]", text)
    return text


def main():
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    res = os.path.join(ref, "resources")
    for name in sorted(os.listdir(res)):
        if name.endswith((".comp", ".glsl")):
            with open(os.path.join(res, name)) as f:
                src = f.read()
            with open(os.path.join(out, name + ".inc"), "w") as f:
                f.write("// GENERATED from the reference's resources/%s by glsl2cpp.py -- do not commit\n" % name)
                f.write(rewrite(src))


if __name__ == "__main__":
    main()
