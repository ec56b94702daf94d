"""Build the native libraries in-tree (no JIT cache: the .so files travel with the repo).

  csrc/libzillum_cuda.so  CUDA kernels + C ABI (include/zillum_cuda.h), sm_100a only
  host/libzillum_host.so  C++ host: Scene / BVH / Camera / Integrators (+ include/zillum_host.h shim)
  host/zillum_render      headless CLI (scene.xml -> EXR / PFM)

nvcc cross-compiles without a GPU.  --fmad=false pins "no FMA contraction" for parity with
the CPU oracle (DESIGN.md "Numerics"); -lineinfo keeps ncu source pages usable; -rdc=true makes
ptxas keep the __noinline__ building blocks out of line (whole-program mode inlined them all
back: 240 KB kernels, instruction-cache bound); -maxrregcount=96 bounds those out-of-line functions (a kernel's register
count is the maximum over its callees: principledBRDFSample alone took 142) — kernels with launch bounds keep their own limits.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(ROOT)
CSRC = os.path.join(ROOT, "csrc")
HOST = os.path.join(ROOT, "host")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
# the image exports CXX=/opt/gcc/bin/g++ (no libgomp); the system compiler has OpenMP
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false", "-rdc=true", "-maxrregcount=96", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math", "--expt-relaxed-constexpr",
    "-ccbin", CXX,
]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-Wall",
             "-Wno-unused-parameter", "-Wno-sign-compare", "-Wno-misleading-indentation"]

HOST_SOURCES = ["BVH.cpp", "Sampler.cpp", "EnvironmentMap.cpp", "Camera.cpp", "MaterialLoader.cpp", "Xml.cpp",
                "ImageIO.cpp", "ImageDecode.cpp", "Model.cpp", "ProceduralMeshes.cpp", "Scene.cpp", "SceneBuiltin.cpp",
                "Integrator.cpp", "HostApi.cpp"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, cwd=None):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=cwd)


def build_cuda(force=False, extra=()):
    out = os.path.join(CSRC, "libzillum_cuda.so")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(REPO, "include", "zillum_cuda.h"))
    if force or _newer(out, deps):
        _run([NVCC, *NVCC_FLAGS, *extra, "-shared", "-o", out, os.path.join(CSRC, "zl_abi.cu"),
              os.path.join(CSRC, "zl_instrumented.cu"), "-ldl"])
    return out


def build_host(force=False):
    cuda = build_cuda()
    out = os.path.join(HOST, "libzillum_host.so")
    deps = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith((".cpp", ".h", ".inc"))]
    deps += [os.path.join(REPO, "include", "zillum_cuda.h"), os.path.join(REPO, "include", "zillum_host.h"), cuda]
    if force or _newer(out, deps):
        _run([CXX, *CXX_FLAGS, "-shared", "-o", out, *[os.path.join(HOST, s) for s in HOST_SOURCES],
              "-L" + CSRC, "-lzillum_cuda", "-Wl,-rpath,$ORIGIN/../csrc"])
    cli = os.path.join(HOST, "zillum_render")
    cli_src = os.path.join(HOST, "zillum_render.cpp")
    if os.path.exists(cli_src) and (force or _newer(cli, [cli_src, out])):
        _run([CXX, *CXX_FLAGS, "-o", cli, cli_src, "-L" + HOST, "-lzillum_host", "-L" + CSRC, "-lzillum_cuda",
              "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/../csrc"])
    return out


def build_hostprep(force=False):
    """host/libzillum_hostprep.so: the host classes linked against NoDevice.cpp instead of the CUDA library (scene preparation only;
    every device entry point fails with ZL_ERR_NO_DEVICE).  Loaded instead of the two product libraries when ZILLUM_HOST_PREP_ONLY=1 —
    bench.py's `--impl reference` arm, which must not pull libzillum_cuda.so into the CPU measurement."""
    out = os.path.join(HOST, "libzillum_hostprep.so")
    srcs = [os.path.join(HOST, s) for s in HOST_SOURCES + ["NoDevice.cpp"]]
    deps = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith((".cpp", ".h", ".inc"))] + [os.path.join(REPO, "include", "zillum_cuda.h")]
    if force or _newer(out, deps):
        _run([CXX, *CXX_FLAGS, "-shared", "-o", out, *srcs])
    return out


def build_oracle(force=False):
    """Test infrastructure: the CPU oracle (never linked into the product)."""
    odir = os.path.join(REPO, "oracle")
    if force:
        _run(["make", "clean"], cwd=odir)
    _run(["make"], cwd=odir)
    return os.path.join(odir, "liboracle.so")


def build_ref(force=False, reference="/root/reference"):
    """Test infrastructure: oracle/_ref/libzillum_ref.so, the reference's own host C++ and GLSL text compiled for the host
    (oracle/Makefile `ref`).  Needs the reference tree, which exists only in the build container; elsewhere the prebuilt
    library that travelled with the repository is used as it is."""
    odir = os.path.join(REPO, "oracle")
    out = os.path.join(odir, "_ref", "libzillum_ref.so")
    if not os.path.isdir(os.path.join(reference, "src", "shader")):
        return out if os.path.exists(out) else None
    if force:
        _run(["make", "clean-ref"], cwd=odir)
    _run(["make", "ref", "REF=" + reference], cwd=odir)
    return out


def build_all(force=False):
    return build_cuda(force), build_host(force), build_hostprep(force), build_oracle(force), build_ref(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
