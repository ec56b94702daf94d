#!/usr/bin/env python3
"""ORACLE build recipe — test infrastructure only.

Rewrites one of the reference's GLSL compute programs into a C++ translation unit that compiles
against oracle/ref_shim/glsl_shim.h.  The shader TEXT is read from the reference tree where it
lies (never copied into this repository); the generated file goes under oracle/_ref/gen/
(git-ignored).  The rewriting is purely lexical — no statement of the reference is restated by
hand — and follows the reference's own preprocessing (src/core/Shader.cpp:180-259):

  @type compute|lib   dropped (compute: the work-group size the host passes is recorded)
  @include f.glsl     the file's text inline, once per program (include record as in Shader.cpp)
  layout(fmt, binding = n) uniform [readonly|writeonly] image2D u;  ->  image2D u;  + registry entry
  uniform T u;        ->  T u;  + registry entry (name, type, address), set by name like glUniform*
  inout T x / out T x ->  T& x
  1.0, 1e-4, .5       ->  1.0f ... (GLSL literals are binary32; C++ ones would be binary64)
  v.xyz / v.rgb / ... ->  v.xyz() ... (multi-component swizzles; single components are fields)
  vecN(f(), f(), ..)  ->  vecN{f(), f(), ..} where the arguments advance the sampler / RNG state
                          (GLSL orders argument evaluation left to right, C++ only inside braces)
  file-scope mutable variables (randSeed, sampleOffset, ...) -> thread_local (one per invocation)
  void main()         ->  void shader_main()

usage: ref_glsl2cpp.py <shader_dir> <program.glsl> <local_x> <local_y> <local_z> <out.cpp> [--kat]
"""
import os
import re
import sys

TYPES = r"(?:bool|int|uint|float|vec2|vec3|vec4|ivec2|ivec3|ivec4|uvec2|uvec3|uvec4|mat3|Ray|SurfaceInfo|HitInfo)"
UNIFORM_KIND = {
    "int": "U_INT", "uint": "U_UINT", "float": "U_FLOAT", "bool": "U_BOOL", "vec2": "U_VEC2", "vec3": "U_VEC3",
    "vec4": "U_VEC4", "ivec2": "U_IVEC2", "mat3": "U_MAT3", "samplerBuffer": "U_SAMPLER_BUFFER",
    "isamplerBuffer": "U_ISAMPLER_BUFFER", "usamplerBuffer": "U_USAMPLER_BUFFER", "sampler2D": "U_SAMPLER_2D",
    "isampler2D": "U_ISAMPLER_2D", "sampler2DArray": "U_SAMPLER_2D_ARRAY", "image2D": "U_IMAGE_2D",
}
FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")
SWIZZLE = re.compile(r"\.((?:[xyzw]{2,4})|(?:[rgba]{2,4}))\b(?!\s*\()")
SIDE_EFFECT_CTOR = re.compile(r"\b(vec[234])\(((?:\s*(?:rand\(\)|sample1D\(s\))\s*,)+\s*(?:rand\(\)|sample1D\(s\))\s*)\)")


def load(shader_dir, name, seen, out):
    """Shader::loadShader: lines starting with '@' are directives; an already recorded include is skipped."""
    with open(os.path.join(shader_dir, name), encoding="utf-8", errors="replace") as f:
        for lineno, line in enumerate(f.read().splitlines(), 1):
            if line.startswith("@"):
                parts = line.split()
                if parts[0] == "@include":
                    if parts[1] not in seen:
                        seen.add(parts[1])
                        load(shader_dir, parts[1], seen, out)
                continue
            out.append((name, lineno, line))


def strip_comments(line, state):
    """remove // and /* */ comments (their text may hold non-ASCII bytes and literal-looking tokens)"""
    res = []
    i = 0
    while i < len(line):
        if state["block"]:
            j = line.find("*/", i)
            if j < 0:
                return "".join(res)
            state["block"] = False
            i = j + 2
        elif line.startswith("//", i):
            break
        elif line.startswith("/*", i):
            state["block"] = True
            i += 2
        else:
            res.append(line[i])
            i += 1
    return "".join(res)


def rewrite(lines):
    uniforms = []
    body = []
    depth = 0
    cstate = {"block": False}
    for name, lineno, raw in lines:
        line = strip_comments(raw, cstate).rstrip()
        if not line.strip():
            continue
        if line.lstrip().startswith("#extension") or line.lstrip().startswith("#version"):
            continue
        m = re.match(r"\s*(?:layout\s*\(([^)]*)\)\s*)?uniform\s+(?:(?:readonly|writeonly)\s+)*(\w+)\s+(\w+)\s*;", line)
        if m and depth == 0:
            layout, typ, uname = m.groups()
            binding, fmt = -1, ""
            if layout:
                for item in layout.split(","):
                    item = item.strip()
                    if item.startswith("binding"):
                        binding = int(item.split("=")[1])
                    else:
                        fmt = item
            uniforms.append((uname, typ, binding, fmt))
            body.append(f"{typ} {uname};   // {name}:{lineno}")
            continue
        line = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", line)
        line = re.sub(r"\bin\s+(" + TYPES + r")\s+(\w+)", r"\1 \2", line)
        line = FLOAT_LIT.sub(lambda mm: mm.group(1) + "f", line)
        line = SWIZZLE.sub(lambda mm: "." + mm.group(1) + "()", line)
        line = SIDE_EFFECT_CTOR.sub(lambda mm: mm.group(1) + "{" + mm.group(2) + "}", line)
        line = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", line)
        if depth == 0 and not line.lstrip().startswith(("#", "const ", "struct ", "}")):
            # a file-scope variable definition (not a function: no '(' before the first '=' or ';')
            mm = re.match(r"(\s*)(" + TYPES + r")\s+(\w+)\s*(=[^;]*)?;\s*$", line)
            if mm:
                line = f"{mm.group(1)}thread_local {line.strip()}"
        body.append(f"{line}   // {name}:{lineno}" if not line.lstrip().startswith("#") else line)
        depth += line.count("{") - line.count("}")
    return uniforms, body


def main():
    shader_dir, prog, lx, ly, lz, out_path = sys.argv[1:7]
    with_kat = "--kat" in sys.argv[7:]
    lines = []
    load(shader_dir, prog, set(), lines)
    uniforms, body = rewrite(lines)
    ident = os.path.splitext(os.path.basename(prog))[0]
    o = []
    o.append(f"// GENERATED by oracle/ref_glsl2cpp.py from {prog} of the reference tree — do not edit, do not commit.")
    o.append('#include "glsl_shim.h"')
    o.append("namespace glsl {")
    o.append(f"namespace prog_{ident} {{")
    o.append(f"static const uvec3 gl_WorkGroupSize = uvec3({lx}u, {ly}u, {lz}u);")
    o.append("static thread_local uvec3 gl_GlobalInvocationID;")
    o.extend(body)
    o.append("static const UniformEntry kUniforms[] = {")
    for uname, typ, binding, fmt in uniforms:
        o.append(f'    {{"{uname}", {UNIFORM_KIND[typ]}, (void*)&{uname}, {binding}, "{fmt}"}},')
    o.append("};")
    o.append("static void invoke(uint gx, uint gy, uint gz) { gl_GlobalInvocationID = uvec3(gx, gy, gz); shader_main(); }")
    if with_kat:
        o.append('#include "ref_kat.inc"')
        kat, trace = "refKat", "refTrace"
    else:
        kat, trace = "nullptr", "nullptr"
    o.append(f'static Program kProgram = {{"{prog}", kUniforms, (int)(sizeof(kUniforms) / sizeof(kUniforms[0])), '
             f"{{{lx}, {ly}, {lz}}}, invoke, {kat}, {trace}, nullptr}};")
    o.append("static struct Reg { Reg() { registerProgram(&kProgram); } } kReg;")
    o.append("}  // namespace prog")
    o.append("}  // namespace glsl")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        f.write("\n".join(o) + "\n")


if __name__ == "__main__":
    main()
