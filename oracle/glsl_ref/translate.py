#!/usr/bin/env python
"""Turn the reference's shader TEXT into something a C++20 compiler accepts — TEST INFRASTRUCTURE ONLY.

    python oracle/glsl_ref/translate.py <reference checkout> <out dir>

Reads assets/shaders/{primary.comp,secondary.comp,blit.fragment}.glsl where they lie in the reference checkout
(never copied into this repository; the outputs go to oracle/_ref/, which is git-ignored), expands their `#include`
lines the way the reference's own loader does (src/engine/graphics/shader.zig:14-40: the line is replaced by the
contents of the named file) and applies the rewrites below.  Every rewrite is syntactic: none changes an expression,
a constant, an operand order or the control flow.  The GLSL library (types, operators, built-ins) is glsl_shim.h.

  R1  `#version` / `#extension` lines and `layout(local_size_x = ..) in;` are dropped (no C++ meaning).
  R2  resource declarations become plain variables of the shim's types:
        layout(fmt, binding = N) uniform [readonly] image2D|image3D|sampler2D name;  ->  static <type> name;
        layout(binding = 8) uniform u_Camera { members };                             ->  static <member>; ...
        layout(binding = N) buffer X { uint name[]; };                                ->  static uint *name;
        layout(location = 0) in vec2 texPos;  /  out vec4 fragColor;                   ->  static thread_local ...;
  R3  parameter qualifiers: `in T x` -> `T x` (by value: GLSL `in` parameters are copies), `inout T x` -> `T &x`.
  R4  floating literals get an `f` suffix (GLSL literals are 32-bit floats; unsuffixed C++ ones are doubles).
  R5  multi-component swizzles become calls: `.xyz` -> `.xyz()`, `.xy` -> `.xy()`.
  R6  `void main()` -> `void shader_main()`.
  R7  `#define MAP_DIMENSION 512` -> `#define MAP_DIMENSION ref_map_dimension`, a variable the harness sets (default 512, the
      text's value): BASELINE configs 3-5 run the 4x world, for which the reference itself would need this line edited.
  variant "entities" only (the dead code of traceEntities made live; never part of the default library):
  E1  the early `return HitInfo(0xFFFFFFFF, positions[id], vec3(0.));` of map.glsl:199 is deleted, so the sub-model
      DDA behind it runs;
  E2  the entity composite primary.comp.glsl keeps commented out (:47-54) is uncommented.
"""
import os
import re
import sys

SHADERS = {"primary": "assets/shaders/primary.comp.glsl", "secondary": "assets/shaders/secondary.comp.glsl", "blit": "assets/shaders/blit.fragment.glsl"}


def expand_includes(root, rel):
    out = []
    for line in open(os.path.join(root, rel)).read().split("\n"):
        i = line.find("#include")
        if i >= 0:
            out.append(open(os.path.join(root, line[i + 9:].strip())).read())  # shader.zig:30-37
        else:
            out.append(line)
    return "\n".join(out)


def translate(text, entities=False):
    if entities and "traceEntities" in text:
        n_before = len(text)
        text = re.sub(r"^[ \t]*return HitInfo\(0xFFFFFFFF, positions\[id\], vec3\(0\.\)\);[ \t]*\n", "", text, count=1, flags=re.M)  # E1
        assert len(text) != n_before, "E1: the early return of traceEntities was not found"
        # E2: the commented block between the two banner comments of primary main()
        m = re.search(r"(// -+ Entity intersection -+\n)(.*?)(\n\s*// -+ Terrain intersection -+)", text, flags=re.S)
        if m:
            body = re.sub(r"^(\s*)// ?", r"\1", m.group(2), flags=re.M)
            text = text[:m.start(2)] + body + text[m.end(2):]
    text = re.sub(r"^\s*#(version|extension)\b.*$", "", text, flags=re.M)                                                      # R1
    text = re.sub(r"layout\s*\(\s*local_size_x[^)]*\)\s*in\s*;", "", text)                                                      # R1
    text = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+(?:readonly\s+)?(image2D|image3D|sampler2D)\s+(\w+)\s*;", r"static \1 \2;", text)  # R2
    text = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+\w+\s*\{([^}]*)\}\s*;",
                  lambda m: "\n".join("static " + ln.strip() for ln in m.group(1).split("\n") if ln.strip()), text)             # R2
    text = re.sub(r"layout\s*\([^)]*\)\s*buffer\s+\w+\s*\{\s*uint\s+(\w+)\s*\[\s*\]\s*;\s*\}\s*;", r"static uint *\1;", text)   # R2
    text = re.sub(r"layout\s*\(\s*location\s*=\s*\d+\s*\)\s*in\s+(\w+)\s+(\w+)\s*;", r"static thread_local \1 \2;", text)      # R2
    text = re.sub(r"^out\s+(\w+)\s+(\w+)\s*;", r"static thread_local \1 \2;", text, flags=re.M)                                 # R2
    assert "layout" not in text, "an unhandled layout() declaration: " + re.search(r".*layout.*", text).group(0)
    text = re.sub(r"\binout\s+(\w+)\s+(\w+)", r"\1 &\2", text)                                                                  # R3
    text = re.sub(r"(?<=[(,])(\s*)in\s+(?=\w)", r"\1", text)                                                                    # R3
    text = re.sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?)(?![\w.])", r"\1f", text)                                    # R4
    text = re.sub(r"\.(xyz|xy)\b(?!\s*\()", r".\1()", text)                                                                     # R5
    text = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", text)                                                       # R6
    text, n = re.subn(r"^#define\s+MAP_DIMENSION\s+512\s*$", "#define MAP_DIMENSION ref_map_dimension", text, flags=re.M)       # R7
    return text


def main():
    root, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    for variant in ("", "_entities"):
        for name, rel in SHADERS.items():
            src = expand_includes(root, rel)
            with open(os.path.join(out, f"{name}{variant}.gen.inc"), "w") as f:
                f.write(f"// GENERATED by oracle/glsl_ref/translate.py from {os.path.join(root, rel)} (+ its #includes); do not commit.\n")
                f.write(translate(src, entities=bool(variant)))
                f.write("\n")


if __name__ == "__main__":
    main()
