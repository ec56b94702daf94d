"""CPU tests: the C++ host preparation against the oracle's restatement, the oracle against
independent references, and the C ABI surface.  No GPU compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, get_scene

SCENES = [("cornell", 64, 48), ("default", 64, 36), ("rungholt_small", 64, 36)]


@pytest.mark.parametrize("name,w,h", SCENES)
def test_bvh_and_hit_table_match_oracle(name, w, h, oracle):
    s, _ = get_scene(name, w, h)
    ob, ot = oracle.build_bvh(s.array("vertices"), s.array("indices"))
    assert np.array_equal(ob, s.array("bounds"))
    assert np.array_equal(ot, s.array("hitTable"))


def test_bvh_parallel_build_is_deterministic(oracle, zl):
    # > 65536 triangles takes the task-parallel path of BVH::quickBuild; the tree must not depend on it
    s = zl.Scene.builtin("rungholt_small?nx=96&ny=64", 32, 32)
    s.flatten()
    assert s.info["numTriangles"] > 65536
    ob, ot = oracle.build_bvh(s.array("vertices"), s.array("indices"))
    assert np.array_equal(ob, s.array("bounds"))
    assert np.array_equal(ot, s.array("hitTable"))


@pytest.mark.parametrize("name,w,h", SCENES)
def test_mtbvh_structure(name, w, h):
    """Threaded-table invariants of BVH.cpp:298-346: every face is a permutation of the nodes,
    miss links point forward, a leaf's miss link is the next entry, the root spans the table."""
    s, _ = get_scene(name, w, h)
    n = s.info["bvhSize"]
    T = s.info["numTriangles"]
    assert n == 2 * T - 1
    table = s.array("hitTable").reshape(6, n, 3)
    for f in range(6):
        node, prim, miss = table[f, :, 0], table[f, :, 1], table[f, :, 2]
        assert np.array_equal(np.sort(node), np.arange(n))
        assert miss[0] == n
        assert np.all(miss > np.arange(n)) and np.all(miss <= n)
        leaf = prim >= 0
        assert leaf.sum() == T and np.array_equal(np.sort(prim[leaf]), np.arange(T))
        assert np.all(miss[leaf] == np.arange(n)[leaf] + 1)
    # all six faces describe the same tree: node -> prim mapping is identical
    ref = dict(zip(table[0, :, 0], table[0, :, 1]))
    for f in range(1, 6):
        assert all(ref[k] == v for k, v in zip(table[f, :, 0], table[f, :, 1]))


@pytest.mark.parametrize("name,w,h", SCENES)
def test_bounds_enclose_triangles(name, w, h):
    s, _ = get_scene(name, w, h)
    n = s.info["bvhSize"]
    b = s.array("bounds").reshape(n, 6)
    v = s.array("vertices").reshape(-1, 3)
    idx = s.array("indices").reshape(-1, 3)
    table = s.array("hitTable").reshape(6, n, 3)[0]
    leaf = table[:, 1] >= 0
    tri = v[idx[table[leaf, 1]]]
    bb = b[table[leaf, 0]]
    assert np.array_equal(tri.min(axis=1), bb[:, :3]) and np.array_equal(tri.max(axis=1), bb[:, 3:])
    assert np.array_equal(b[0, :3], v[idx.reshape(-1)].min(axis=0)) and np.array_equal(b[0, 3:], v[idx.reshape(-1)].max(axis=0))


@pytest.mark.parametrize("name,w,h", SCENES)
def test_light_table_matches_oracle(name, w, h, oracle):
    s, _ = get_scene(name, w, h)
    first, num, power = s.light_meshes()
    lp, pdf, total = oracle.light_table(s.array("vertices"), s.array("indices"), first, num, power)
    alias, prob = oracle.alias_table(pdf)
    assert np.array_equal(lp, s.array("lightPower"))
    assert np.array_equal(alias, s.array("lightAlias")) and np.array_equal(prob, s.array("lightProb"))
    assert total == s.desc.contents.lightSum
    assert s.info["nLightTriangles"] == num.sum()


def test_alias_table_is_a_valid_distribution(zl, oracle):
    rng = np.random.default_rng(3)
    pdf = rng.random(257).astype(np.float32) ** 3
    alias, prob = oracle.alias_table(pdf)
    n = pdf.size
    recon = prob.astype(np.float64).copy()
    np.add.at(recon, alias, 1.0 - prob.astype(np.float64))
    assert np.allclose(recon / n, pdf / pdf.sum(dtype=np.float64), atol=2e-6)
    from zillumgl_b200 import _native as N
    a2, p2 = np.empty(n, np.int32), np.empty(n, np.float32)
    N.host.zh_alias_table(pdf.ctypes.data_as(C.POINTER(C.c_float)), n, a2.ctypes.data_as(C.POINTER(C.c_int32)),
                          p2.ctypes.data_as(C.POINTER(C.c_float)))
    assert np.array_equal(a2, alias) and np.array_equal(p2, prob)


def test_env_tables_match_oracle(oracle):
    s, _ = get_scene("rungholt_small", 64, 36)
    d = s.desc.contents
    assert (d.envW, d.envH) == (2048, 1024)
    alias, prob, total = oracle.env_tables(s.array("envMap"), d.envW, d.envH)
    assert np.array_equal(alias, s.array("envAlias")) and np.array_equal(prob, s.array("envAliasProb"))
    assert d.envSum == float(int(total))          # int-truncated on purpose (EnvironmentMap.h:22)
    # marginal column reproduces the row weights
    w, h = d.envW, d.envH
    rgb = s.array("envMap").reshape(h, w, 3).astype(np.float64)
    lum = rgb @ np.array([0.2126, 0.7152, 0.0722])
    rows = (lum * np.sin((np.arange(h) + 0.5) / h * np.pi)[:, None]).sum(axis=1)
    pm, am = prob.reshape(h, w + 1)[:, w].astype(np.float64), alias.reshape(h, w + 1)[:, w]
    recon = pm.copy()
    np.add.at(recon, am, 1.0 - pm)
    assert np.allclose(recon / h, rows / rows.sum(), atol=1e-5)


def test_sobol_matrices_and_sampler(zl, oracle):
    """The generator matrices are the Joe-Kuo table (golden copy generated by
    tools/gen_sobol_matrices.py, checked there against the reference header); sobolSample is
    checked against scipy's independent implementation (Gray-code ordered)."""
    from scipy.stats import qmc
    from zillumgl_b200 import _native as N
    golden = np.load(os.path.join(ROOT, "tests", "golden", "sobol_matrices_256x32.npy"))
    s, _ = get_scene("cornell", 64, 48)
    assert np.array_equal(s.array("sobolMatrices").reshape(256, 32), golden)
    dims, npts = 64, 1024
    pts = qmc.Sobol(d=dims, scramble=False, bits=32).random(npts)
    for i in (0, 1, 2, 3, 5, 64, 255, 1023):
        g = i ^ (i >> 1)
        for dim in (0, 1, 2, 7, 31, 63):
            v = oracle.sobol_sample(golden, g, dim)
            assert v == N.host.zh_sobol_sample(g, dim)
            assert v / 2.0 ** 32 == pts[i, dim]


def test_wang_hash_known_answers(oracle):
    def wang(seed):
        seed = ((seed ^ 61) ^ (seed >> 16)) & 0xffffffff
        seed = (seed * 9) & 0xffffffff
        seed ^= seed >> 4
        seed = (seed * 0x27d4eb2d) & 0xffffffff
        seed ^= seed >> 15
        return seed
    for x in (0, 1, 2, 61, 0xdeadbeef, 0xffffffff, 123456789):
        assert oracle.lib.zo_hash(x) == wang(x)


def test_camera_uniforms_match_oracle(zl, oracle):
    from zillumgl_b200 import ZlCamera
    s = zl.Scene.builtin("cornell", 64, 48)
    for pos, ang, fov, lens, focal in [((0, -8, 3), (0, 0, 0), 45, 0, 1), ((1, 2, 3), (33, -12, 0.3), 60, 0.05, 4.5),
                                       ((-3, 0.5, 9), (181, 40, -1.2), 25, 0.0, 2.0)]:
        s.set_camera(pos, ang, fov, lens, focal)
        cam = s.camera()
        ref = ZlCamera()
        p, a = np.asarray(pos, np.float32), np.asarray(ang, np.float32)
        oracle.lib.zo_camera_update(p.ctypes.data_as(C.POINTER(C.c_float)), a.ctypes.data_as(C.POINTER(C.c_float)),
                                    fov, cam.asp, lens, focal, C.cast(C.byref(ref), C.c_void_p))
        assert bytes(cam) == bytes(ref)
        F, R, U = np.array(cam.F), np.array(cam.R), np.array(cam.U)
        assert abs(F @ R) < 1e-6 and abs(F @ U) < 1e-6 and abs(R @ U) < 1e-6
        M = np.array(cam.matInv).reshape(3, 3).T @ np.stack([R, U, F], axis=1)
        assert np.allclose(M, np.eye(3), atol=1e-5)


def test_half_rounding_matches_numpy(oracle):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(size=2000) * 10.0 ** rng.integers(-9, 6, 2000), [0.0, 65504.0, 65520.0, 6e-8, 3e-8, 1e-7]]).astype(np.float32)
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).astype(np.float32)
    got = np.array([oracle.lib.zo_round_to_half(float(v)) for v in x], np.float32)
    assert np.array_equal(got, want)


def test_c_abi_exports_every_declared_symbol(zl):
    from zillumgl_b200 import _native as N
    for header, lib in (("zillum_cuda.h", N.cuda), ("zillum_host.h", N.host)):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names = set(re.findall(r"\b(z[lh]_[a-z0-9_]+)\s*\(", text))
        assert len(names) > 20
        for n in sorted(names):
            assert hasattr(lib, n), f"{header}: {n} is declared but not exported"
    assert N.cuda.zl_abi_version() == 3


def test_headers_are_plain_c_and_the_library_links_from_c(zl, tmp_path):
    """The drop-in boundary is a C ABI: include/*.h compile as C99 (-pedantic) and as C++11, and a C program linked against
    libzillum_cuda.so resolves and calls an entry point (no compute: zl_abi_version, zl_last_error_string)."""
    import shutil
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc, lib = os.path.join(root, "include"), os.path.join(root, "zillumgl_b200", "csrc")
    (tmp_path / "c.c").write_text('#include "zillum_cuda.h"\n#include "zillum_host.h"\n#include <stdio.h>\n'
                                  'int main(void) { printf("%d %s|\\n", zl_abi_version(), zl_last_error_string()); return 0; }\n')
    (tmp_path / "c.cpp").write_text('#include "zillum_cuda.h"\n#include "zillum_host.h"\nint main() { return zl_abi_version() > 0 ? 0 : 1; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(tmp_path / "c.c")], check=True)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(tmp_path / "c.cpp")], check=True)
    exe = str(tmp_path / "c_abi")
    subprocess.run(["gcc", "-std=c99", "-I", inc, str(tmp_path / "c.c"), "-o", exe, "-L", lib, "-lzillum_cuda", f"-Wl,-rpath,{lib}"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.split()[0] == str(zl._native.cuda.zl_abi_version())


def test_abi_struct_layout(zl):
    from zillumgl_b200 import ZlCamera, ZlRenderParams, ZlSceneDesc
    assert C.sizeof(ZlCamera) == 4 * 25
    assert C.sizeof(ZlRenderParams) == 4 * 25 + 4 * 14
    assert C.sizeof(ZlSceneDesc) == 18 * 8 + 14 * 4 + 8 + 8      # ABI v3: + sizeIndices


def test_scene_xml_dialect(zl, tmp_path):
    """res/scene.xml semantics: type=light instances become emitters, material type=default keeps
    the model's own, unknown tags (albedo/tint) are ignored like MaterialLoader.cpp does."""
    xml = zl._native.host.zh_builtin_scene_xml(b"default", 1280, 720).decode()
    assert "<albedo" in xml and 'type="light"' in xml
    p = tmp_path / "scene.xml"
    p.write_text(xml)
    s = zl.Scene.from_file(p)
    s.flatten()
    info = s.info
    assert (info["filmWidth"], info["filmHeight"], info["sampler"]) == (1280, 720, 1)
    assert info["nLightTriangles"] == 2 and info["objPrimCount"] == info["numTriangles"] - 2
    mats = s.array("materials").reshape(-1, 16)
    types = mats[:, 13].view(np.int32)
    assert list(types) == [0, 2, 3]                      # default lambertian, metalWorkflow, dielectric
    assert np.allclose(mats[1, :3], 1.0) and mats[1, 5] == 1.0 and np.isclose(mats[1, 3], 0.1)   # <albedo> ignored -> white
    assert mats[2, 12] == 1.5 and mats[2, 3] == 0.0
    # light faces down: its two triangles have normals (0,0,-1)
    v = s.array("vertices").reshape(-1, 3)
    idx = s.array("indices").reshape(-1, 3)
    lt = v[idx[-2:]]
    assert np.allclose(lt[..., 2], 10.0, atol=1e-5)
    nrm = s.array("normals").reshape(-1, 3)[idx[-1]]
    assert np.allclose(nrm, [0, 0, -1], atol=1e-5)


def test_scene_xml_syntax_variants(zl, tmp_path):
    """The layout of the reference's res/scene.xml (tabs, `<tag ... />` with a blank before the slash, whole instances and the
    envMap inside comments, tags nobody reads such as toneMapping / numSamples) plus what an XML writer may emit: single-quoted
    attributes, blanks around `=`, an encoding declaration, a leading comment, an entity in an attribute."""
    xml = """<?xml version='1.0' encoding="UTF-8"?>
<!-- leading comment -->
<scene name="syntax &amp; variants">
\t<integrator type='path'>
\t\t<maxBounce value="3" />
\t\t<size width="64"   height = "36"/>
\t\t<toneMapping type="filmic" />
\t</integrator>
\t<sampler type="sobol"><numSamples value="256" /></sampler>
\t<camera type="thinLens">
\t\t<position value="0 -8 3" /><angle value="0 0 0" />
\t\t<fov value="45" /><lensRadius value="0" /><focalDistance value="1" />
\t</camera>
\t<modelInstances>
\t\t<!-- a commented instance
\t\t<modelInstance path="builtin:cube" name="ghost" type="object"><transform translate="0 0 0" scale="1 1 1" rotate="0 0 0" /><material type="default" /></modelInstance>
\t\t-->
\t\t<modelInstance path="builtin:square" name="floor" type="object">
\t\t\t<transform translate="0 0 0" scale="100 100 1" rotate="0 0 0" />
\t\t\t<material type="default" />
\t\t</modelInstance>
\t\t<modelInstance path="builtin:square" name="lamp" type="light">
\t\t\t<transform translate="0 0 10" scale="2 2 1" rotate="180 0 0" />
\t\t\t<radiance value="20 20 20" />
\t\t</modelInstance>
\t</modelInstances>
\t<!-- <envMap path="none.png" /> -->
</scene>
"""
    (tmp_path / "s.xml").write_text(xml)
    s = zl.Scene.from_file(tmp_path / "s.xml")
    s.flatten()
    i = s.info
    assert (i["numTriangles"], i["objPrimCount"], i["nLightTriangles"], i["filmWidth"], i["filmHeight"], i["sampler"]) == (4, 2, 2, 64, 36, 1)
    assert (i["envW"], i["envH"]) == (1, 1)                                   # the commented envMap is not loaded: 1x1 black
    assert np.allclose(s.array("vertices").reshape(-1, 3)[:4, :2].max(), 50.0)    # the 100 x 100 floor; the commented cube is absent


def test_obj_reader_dialect(zl, tmp_path):
    """host/Model.cpp reads the OBJ text in one pass over its bytes (Resource.cpp:37-93 hands this to Assimp): CRLF line ends, blank and
    comment lines, tabs, '+' signs, polygons triangulated as fans, negative (relative) indices, all four corner forms (v, v/vt, v//vn,
    v/vt/vn) incl. empty trailing fields, one mesh per material, vertices joined on the resolved (v, vt, vn) triple, a face with an
    index out of range skipped, smooth normals generated for a mesh that has a corner without one, v flipped (aiProcess_FlipUVs)."""
    obj = ("mtllib t.mtl\r\n# comment\r\n\r\nv 0 0 0\r\nv 1 0 0\r\nv 1 1 0\r\nv 0 1 0\r\nv  0.5\t0.5 1e0\r\nv +2 -0.0 .5\r\n"
           "vt 0 0\r\nvt 1 0\r\nvt 1 0.75\r\nvt 0 1\r\nvn 0 0 1\r\n"
           "usemtl red\r\nf 1/1/1 2/2/1 3/3/1 4/4/1\r\n"
           "usemtl blue\r\nf 1 2 5\r\nf -5//1 -4//1 -2//1\r\nf 2/2 3/3 5/1\r\n"
           "o obj2\r\ng grp\r\ns off\r\n"
           "usemtl red\r\nf 1/1/1 3/3/1 6/2/1\r\nf 1/1/ 2/2/ 99/1/1\r\n"
           "usemtl nomat\r\nf 1//1 2//1 6//1\r\n")
    (tmp_path / "t.obj").write_bytes(obj.encode())
    (tmp_path / "t.mtl").write_text("newmtl red\nKd 1 0 0\nnewmtl blue\nKd 0 0.5 1\n")
    probe = zl.Scene.builtin("cornell", 32, 24)
    xml = probe.builtin_xml("cornell", 32, 24).replace('path="builtin:cornell"', f'path="{tmp_path / "t.obj"}"')
    (tmp_path / "scene.xml").write_text(xml)
    s = zl.Scene.from_file(tmp_path / "scene.xml")
    model = [m for m in s.models() if not m["isLight"]][0]
    assert np.array_equal(np.array(model["materials"])[:, :3], np.array([[1, 0, 0], [0, 0.5, 1]], np.float32))
    red, blue, nomat = model["meshes"]
    assert (int(red["matIndex"]), int(blue["matIndex"]), int(nomat["matIndex"])) == (0, 1, 0)      # unknown material -> the first one
    # red: the quad as a fan (0 1 2, 0 2 3), then a triangle that re-uses two joined corners; the skipped face left two corners
    # (1/1/ and 2/2/, without normals) behind, so this mesh gets generated normals
    assert list(red["idx"]) == [0, 1, 2, 0, 2, 3, 0, 2, 4]
    assert np.array_equal(red["pos"][:5], np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [2, -0.0, 0.5]], np.float32))
    assert np.array_equal(red["tex"][:5], np.array([[0, 1], [1, 1], [1, 0.25], [0, 0], [1, 1]], np.float32))
    assert red["pos"].shape[0] == 7 and np.allclose(np.linalg.norm(red["nrm"][:5], axis=1), 1.0, atol=1e-6)
    # blue: v-only corners, relative indices (-5 = vertex 2 of the six read so far), v/vt corners; no normals in the file for some
    assert list(blue["idx"]) == [0, 1, 2, 3, 4, 5, 6, 7, 8]
    assert np.array_equal(blue["pos"][3:6], np.array([[1, 0, 0], [1, 1, 0], [0.5, 0.5, 1]], np.float32))
    assert np.array_equal(blue["tex"][6:9], np.array([[1, 1], [1, 0.25], [0, 1]], np.float32))
    # nomat: every corner has the file's normal -> kept as read
    assert list(nomat["idx"]) == [0, 1, 2] and np.array_equal(nomat["nrm"], np.tile(np.array([[0, 0, 1]], np.float32), (3, 1)))


def test_sponza_and_rungholt_triangle_budgets(zl):
    s = zl.Scene.builtin("sponza", 32, 32)
    s.flatten()
    assert s.info["numTriangles"] == 262144 and s.info["numTextures"] == 1
    s2 = zl.Scene.builtin("rungholt?nx=32&ny=16", 32, 32)
    s2.flatten()
    assert s2.info["numTriangles"] == 32 * 16 * 12


def test_image_writers(zl, tmp_path):
    rng = np.random.default_rng(1)
    img = rng.random((5, 7, 4)).astype(np.float32)
    assert zl.write_pfm(tmp_path / "a.pfm", img) and zl.write_exr(tmp_path / "a.exr", img)
    raw = (tmp_path / "a.pfm").read_bytes()
    head, data = raw.split(b"-1.0\n", 1)
    assert head == b"PF\n7 5\n"
    assert np.array_equal(np.frombuffer(data, np.float32).reshape(5, 7, 3), img[..., :3])
    exr = (tmp_path / "a.exr").read_bytes()
    assert exr[:4] == bytes([0x76, 0x2f, 0x31, 0x01]) and len(exr) > 5 * 7 * 12


@pytest.mark.parametrize("name,w,h", SCENES)
def test_size_indices_describe_the_threaded_tree(name, w, h):
    """ZlSceneDesc::sizeIndices (the builder's pre-order tree, input of the device-side MTBVH threading) against the
    host hit table: subtree sizes are the miss-link distances, leaves carry primIndex | 0x80000000 (BVH.cpp:217-346)."""
    s, _ = get_scene(name, w, h)
    n = s.info["bvhSize"]
    size = s.array("sizeIndices")
    assert size.shape == (n,)
    table = s.array("hitTable").reshape(6, n, 3)
    leaf = size < 0
    assert leaf.sum() == s.info["numTriangles"] and size[0] == (n if n > 1 else size[0])
    for f in (0, 3):
        node, prim, miss = table[f, :, 0], table[f, :, 1], table[f, :, 2]
        span = miss - np.arange(n)
        assert np.array_equal(span[~leaf[node]], size[node][~leaf[node]])
        assert np.array_equal(prim[leaf[node]], size[node][leaf[node]] & 0x7fffffff)
    # pre-order: the left child of an interior node is the next node, the right one follows the left subtree
    interior = np.nonzero(~leaf)[0]
    lsize = np.where(leaf[interior + 1], 1, size[interior + 1])
    right = interior + 1 + lsize
    rsize = np.where(leaf[right], 1, size[right])
    assert np.array_equal(size[interior], 1 + lsize + rsize)


def test_device_mtbvh_option_skips_the_host_table(zl, oracle):
    s = zl.Scene.builtin("cornell", 32, 24)
    s.set_device_mtbvh(True)
    s.flatten()
    assert s.array("hitTable").size == 0 and s.array("sizeIndices").size == s.info["bvhSize"]
    # the oracle threads its own table for such a scene and traces it like the host-threaded twin
    import oracle_lib
    o = oracle_lib.OracleScene(s.desc)
    s2, o2 = get_scene("cornell", 32, 24)
    from conftest import random_rays
    rays = random_rays(s2, 2000, 5)
    a, b = o.trace_rays(rays), o2.trace_rays(rays)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_product_path_fails_loudly_without_a_device(zl):
    """There is no CPU fallback: on a machine without a GPU every compute entry point of the C ABI returns an error with a
    message (and the argument checks hold everywhere).  Skipped on a GPU box, where the -m gpu tests exercise the calls."""
    import zillumgl_b200._native as N
    if zl.device_count() > 0:
        pytest.skip("a CUDA device is present")
    s, _ = get_scene("cornell", 32, 24)
    out = C.c_void_p()
    rc = N.cuda.zl_scene_create(C.cast(s.desc, C.POINTER(N.ZlSceneDesc)), C.byref(out))
    assert rc != 0 and not out.value and b"device" in N.cuda.zl_last_error_string().lower()
    with pytest.raises(zl.ZillumError):
        s.upload()
    v = s.array("vertices"); idx = s.array("indices")
    with pytest.raises(zl.ZillumError):
        zl.build_bvh(v, idx)
    # argument checks do not need a device
    assert N.cuda.zl_film_flush(None, None) != 0
    assert N.cuda.zl_film_download_wait(None) != 0
    assert N.cuda.zl_build_bvh(None, 0, None, 0, None, None, None) != 0
    assert N.cuda.zl_launch_path_pass(None, None, None, 2, None) != 0
    assert N.cuda.zl_film_postprocess(None, 1.0, 1, None, None, None) != 0


def test_parallel_generator_and_flatten_equal_the_single_threaded_ones(zl):
    """The Rungholt-class generator fills row chunks in parallel and the scene flatten fills per-mesh slices in parallel (round 2):
    with one OpenMP thread (a separate process) the arrays must be the same, byte for byte."""
    import hashlib
    import subprocess
    import sys
    from conftest import ROOT
    keys = ("vertices", "normals", "texcoords", "indices", "matTexIndices", "materials")

    def digest(scene):
        return {k: hashlib.md5(scene.array(k).tobytes()).hexdigest() for k in keys}
    s = zl.Scene.builtin("rungholt_small", 64, 36)
    s.flatten()
    here = digest(s)
    code = ("import sys, json, hashlib; sys.path.insert(0, %r); import zillumgl_b200 as zl; s = zl.Scene.builtin('rungholt_small', 64, 36); s.flatten(); "
            "print(json.dumps({k: hashlib.md5(s.array(k).tobytes()).hexdigest() for k in %r}))" % (ROOT, keys))
    env = dict(os.environ, OMP_NUM_THREADS="1", ZILLUM_HOST_PREP_ONLY="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    import json
    assert json.loads(out.stdout.strip().splitlines()[-1]) == here
