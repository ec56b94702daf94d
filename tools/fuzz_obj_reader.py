#!/usr/bin/env python3
"""Mutation fuzzing of the OBJ reader (zillumgl_b200/host/Model.cpp) under AddressSanitizer + UBSan.

    python tools/fuzz_obj_reader.py [cases=4000] [seed=1]

Builds the host sources with -fsanitize=address,undefined around a main() that opens every file given on its command line,
mutates a seed OBJ that uses every corner form (insert tokens such as `//`, `-`, huge integers, NUL bytes; delete; flip; truncate)
and reports anything the sanitizers print.  Last run (round 2): 4000 cases, no report.
"""
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "zillumgl_b200", "host")
SOURCES = ["BVH", "Sampler", "EnvironmentMap", "Camera", "MaterialLoader", "Xml", "ImageIO", "ImageDecode", "Model", "ProceduralMeshes", "Scene",
           "SceneBuiltin", "Integrator", "HostApi", "NoDevice"]
MAIN = r'''
#include "Model.h"
#include <cstdio>
int main(int argc, char** argv) {
    size_t tris = 0;
    for (int i = 1; i < argc; i++) {
        auto m = zillum::Resource::openModelInstance(argv[i]);
        if (m) for (auto& mi : m->meshInstances()) tris += mi->meshData->indices.size() / 3;
    }
    std::printf("triangles %zu\n", tris);
}
'''
SEED = (b"mtllib t.mtl\r\n# comment\r\n\r\nv 0 0 0\r\nv 1 0 0\r\nv 1 1 0\r\nv 0 1 0\r\nv  0.5\t0.5 1e0\r\nv +2 -0.0 .5\r\nvt 0 0\r\nvt 1 0\r\nvt 1 1\r\nvt 0 1\r\nvn 0 0 1\r\n"
        b"usemtl red\r\nf 1/1/1 2/2/1 3/3/1 4/4/1\r\nusemtl blue\r\nf 1 2 5\r\nf -5//1 -4//1 -2//1\r\nf 2/2 3/3 5/1\r\no obj2\r\ng grp\r\ns off\r\n"
        b"usemtl red\r\nf 1/1/1 3/3/1 6/2/1\r\nf 1/1/ 2/2/ 99/1/1\r\nf 3 4 5 6 1\r\nusemtl nomat\r\nf 1//1 2//1 6//1\r\n")
TOKENS = [b"f", b"v", b"vt", b"vn", b"/", b"//", b"-", b"+", b"1e999", b"nan", b"inf", b"-0", b"2147483648", b"-2147483649", b"99999999999999999999", b"\n", b"\r",
          b" ", b"\t", b"usemtl", b"mtllib", b"\x00", b"\xff", b".", b"f 1/", b"f -", b"f +", b"f 1//", b"f /1", b"\nf"]


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    work = tempfile.mkdtemp(prefix="zl_fuzz_obj_")
    open(os.path.join(work, "main.cpp"), "w").write(MAIN)
    exe = os.path.join(work, "objasan")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fopenmp", "-w", "-I" + HOST,
                           "-I" + os.path.join(ROOT, "include"), os.path.join(work, "main.cpp")] + [os.path.join(HOST, s + ".cpp") for s in SOURCES] + ["-o", exe])
    open(os.path.join(work, "t.mtl"), "w").write("newmtl red\nKd 1 0 0\nmap_Kd missing.png\nnewmtl blue\nKd 0 0 1\n")
    files = []
    for it in range(cases):
        b = bytearray(SEED)
        for _ in range(rnd.randint(1, 10)):
            k, pos = rnd.randint(0, 3), rnd.randint(0, len(b))
            if k == 0:
                b[pos:pos] = rnd.choice(TOKENS)
            elif k == 1 and len(b) > 2:
                del b[pos:pos + rnd.randint(1, 6)]
            elif k == 2 and pos < len(b):
                b[pos] = rnd.randint(0, 255)
            else:
                b = b[:pos]
        files.append(os.path.join(work, f"m{it}.obj"))
        open(files[-1], "wb").write(bytes(b))
    reports = 0
    for i in range(0, cases, 500):
        out = subprocess.run([exe] + files[i:i + 500], capture_output=True, text=True, errors="replace")
        text = out.stdout + out.stderr
        bad = [ln for ln in text.splitlines() if "Sanitizer" in ln or "runtime error" in ln]
        if out.returncode != 0 or bad:
            reports += 1
            print("\n".join(bad[:40]))
    print(f"{cases} cases, {reports} batches with sanitizer reports; work dir {work}")
    return 1 if reports else 0


if __name__ == "__main__":
    sys.exit(main())
