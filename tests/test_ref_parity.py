"""Parity against THE REFERENCE ITSELF (oracle/_ref/libzillum_ref.so): the reference's own C++ host code
compiled unmodified against stand-in third-party headers, and the reference's own GLSL text compiled as
C++ (recipe: oracle/Makefile `ref`).  Everything here is bit for bit.

CPU tests (not gpu): the oracle (oracle/) and the product's host library against the reference —
this is what pins the oracle: BVH + MTBVH hit table, alias tables, Sobol, camera,
environment tables, every per-function KAT, traversal on the §8(d) ray mix, all three integrators,
post-processing, and whole scenes through the reference's own Scene::load / createGLContext and its
integrator host glue (NaivePath.cpp, LightPath.cpp, TriplePath.cpp).

GPU tests: the CUDA path against the reference directly (KATs, 2^22-ray traversal sets, films)."""
import ctypes as C
import re

import numpy as np
import pytest

import ref_lib
from conftest import get_scene, random_rays, random_soup

if not ref_lib.available():
    pytest.skip("oracle/_ref is not built and the reference tree is not present", allow_module_level=True)

SCENES = [("cornell", 64, 48), ("default", 64, 36), ("rungholt_small", 64, 36), ("sponza_light", 48, 27)]
_REF = {}


def ref_scene(name, w, h):
    if (name, w, h) not in _REF:
        s, _ = get_scene(name, w, h)
        _REF[(name, w, h)] = ref_lib.RefScene(s.desc)
    return _REF[(name, w, h)]


def assert_same_bits(a, b, what=""):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    if a.dtype == np.float32:
        bad = (a.view(np.uint32) != b.view(np.uint32)) & ~(np.isnan(a) & np.isnan(b))      # a generated NaN's payload is hardware-defined
    else:
        bad = a != b
    assert not bad.any(), (what, int(bad.sum()), a[bad][:4], b[bad][:4])


def params(zl, s, w, h, **kw):
    p = zl.ZlRenderParams()
    p.camera = s.camera(); p.camera.asp = w / h
    p.filmW, p.filmH = w, h
    p.maxDepth, p.sampleLight, p.lightPortion, p.sampler = 4, 1, 0.5, 1
    p.spp, p.freeCounter = 3, 4
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def bits(a):
    return np.asarray(a).astype(np.int32).view(np.float32)


def unit(rng, k):
    v = rng.normal(size=(k, 3)).astype(np.float32)
    return v / np.linalg.norm(v, axis=1, keepdims=True)


# ---------------------------------------------------------------------------------------------------------------------
# host preparation: one reference function at a time
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,w,h", SCENES + [("sponza", 48, 27)])
def test_bvh_and_mtbvh_equal_reference(name, w, h, zl, oracle):
    """BVH::build = quickBuild + buildHitTable (BVH.cpp:116-346): bounds and all six threaded orderings."""
    s, _ = get_scene(name, w, h)
    v, i = s.array("vertices"), s.array("indices")
    rb, rt = ref_lib.build_bvh(v, i)
    ob, ot = oracle.build_bvh(v, i)
    assert_same_bits(ob, rb, "oracle bounds"); assert_same_bits(ot, rt, "oracle hit table")
    assert_same_bits(s.array("bounds"), rb, "host bounds"); assert_same_bits(s.array("hitTable"), rt, "host hit table")


def test_bvh_degenerate_inputs_equal_reference(oracle, zl):
    """coincident centroids (the r == size fallback of partition, BVH.cpp:111-112), 1 and 2 triangles, zero-area triangles"""
    rng = np.random.default_rng(3)
    cases = []
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    cases.append(np.tile(tri, (1, 1, 1)))
    cases.append(np.stack([tri, tri + 2]))
    cases.append(np.tile(tri, (9, 1, 1)))                                              # identical triangles
    cases.append(np.tile(tri, (5, 1, 1)) + np.arange(5, dtype=np.float32).reshape(5, 1, 1) * np.array([0, 0, 1], np.float32))
    cases.append(rng.random((257, 3, 3)).astype(np.float32))
    z = rng.random((40, 3, 3)).astype(np.float32); z[::3, 2] = z[::3, 1]                 # zero-area
    cases.append(z)
    for t in cases:
        v = t.reshape(-1, 3); idx = np.arange(v.shape[0], dtype=np.uint32)
        rb, rt = ref_lib.build_bvh(v, idx)
        ob, ot = oracle.build_bvh(v, idx)
        assert_same_bits(ob, rb); assert_same_bits(ot, rt)
        from zillumgl_b200 import _native as N
        n = 2 * (idx.size // 3) - 1
        hb, ht, sec = np.empty(6 * n, np.float32), np.empty(18 * n, np.int32), (C.c_double * 2)()
        vv = np.ascontiguousarray(v, np.float32)
        N.host.zh_build_bvh(vv.ctypes.data_as(C.POINTER(C.c_float)), vv.shape[0], idx.ctypes.data_as(C.POINTER(C.c_uint32)), idx.size // 3,
                            hb.ctypes.data_as(C.POINTER(C.c_float)), ht.ctypes.data_as(C.POINTER(C.c_int32)), sec)
        assert_same_bits(hb, rb); assert_same_bits(ht, rt)


def _host_build_bvh(v, idx):
    from zillumgl_b200 import _native as N
    n = 2 * (idx.size // 3) - 1
    hb, ht, sec = np.empty(6 * n, np.float32), np.empty(18 * n, np.int32), (C.c_double * 2)()
    vv = np.ascontiguousarray(v, np.float32)
    N.host.zh_build_bvh(vv.ctypes.data_as(C.POINTER(C.c_float)), vv.shape[0], idx.ctypes.data_as(C.POINTER(C.c_uint32)), idx.size // 3,
                        hb.ctypes.data_as(C.POINTER(C.c_float)), ht.ctypes.data_as(C.POINTER(C.c_int32)), sec)
    return hb, ht


@pytest.mark.parametrize("seed", range(12))
def test_bvh_random_soups_equal_reference(seed, oracle, zl):
    """Randomised inputs for the binned-SAH split and its tie rules (BVH.cpp:217-296): clustered soups, vertices snapped to a coarse
    grid (many equal centroids and bucket boundaries hit exactly), duplicated triangles, slivers along one axis, and indexed meshes
    with shared vertices — oracle and host library against the reference's own BVH.cpp, bounds and all six orderings."""
    v, idx = random_soup(seed)
    rb, rt = ref_lib.build_bvh(v, idx)
    ob, ot = oracle.build_bvh(v, idx)
    assert_same_bits(ob, rb, "oracle bounds"); assert_same_bits(ot, rt, "oracle hit table")
    hb, ht = _host_build_bvh(v, idx)
    assert_same_bits(hb, rb, "host bounds"); assert_same_bits(ht, rt, "host hit table")


def test_alias_table_equals_reference(oracle, zl):
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 7, 100, 4099, 65536):
        for pdf in (rng.random(n).astype(np.float32) ** 3, np.ones(n, np.float32), (rng.random(n) < 0.3).astype(np.float32) + 1e-3):
            ra, rp = ref_lib.alias_table(pdf)
            oa, op = oracle.alias_table(pdf)
            assert_same_bits(oa, ra); assert_same_bits(op, rp)
            ha, hp = np.empty(n, np.int32), np.empty(n, np.float32)
            from zillumgl_b200 import _native as N
            N.host.zh_alias_table(pdf.ctypes.data_as(C.POINTER(C.c_float)), n, ha.ctypes.data_as(C.POINTER(C.c_int32)),
                                  hp.ctypes.data_as(C.POINTER(C.c_float)))
            assert_same_bits(ha, ra); assert_same_bits(hp, rp)


def test_sobol_equals_reference(oracle, zl):
    from zillumgl_b200 import _native as N
    mats = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "sobol_matrices_256x32.npy")).astype(np.uint32).reshape(-1)
    assert np.array_equal(mats, ref_lib.sobol_matrices())                  # the reference's own table (SobolMatrices256x32.h)
    rng = np.random.default_rng(2)
    for i, d in zip(rng.integers(0, 131072, 3000), rng.integers(0, 256, 3000)):
        r = ref_lib.sobol_sample(i, d)
        assert r == oracle.sobol_sample(mats, int(i), int(d)) == N.host.zh_sobol_sample(int(i), int(d))


def test_camera_equals_reference(oracle, zl):
    """Camera::update (Camera.cpp:149-162) + the camera uniforms of NaivePath.cpp:49-58, 300 random poses incl. roll"""
    rng = np.random.default_rng(4)
    fp = C.POINTER(C.c_float)
    host = zl.Scene.builtin("cornell", 64, 48)
    for _ in range(300):
        pos = (rng.normal(size=3) * 5).astype(np.float32)
        ang = (rng.random(3) * np.array([360, 170, 60]) - np.array([180, 85, 30])).astype(np.float32)
        fov, asp, lens, foc = float(rng.uniform(10, 89)), float(rng.uniform(0.5, 2.5)), float(rng.uniform(0, 0.2)), float(rng.uniform(0.5, 10))
        rc = ref_lib.camera_update(zl.ZlCamera, pos, ang, fov, asp, lens, foc)
        oc = zl.ZlCamera()
        oracle.lib.zo_camera_update(pos.ctypes.data_as(fp), ang.ctypes.data_as(fp), fov, asp, lens, foc, C.cast(C.byref(oc), C.c_void_p))
        assert bytes(rc) == bytes(oc)
        host.set_camera(pos, ang, fov, lens, foc)                      # the product's own Camera class at every pose
        hc = host.camera(); hc.asp = asp
        assert bytes(hc) == bytes(rc)
    s = zl.Scene.builtin("cornell", 64, 48)
    s.set_camera(pos, ang, fov, lens, foc)
    hc = s.camera(); hc.asp = asp
    assert bytes(hc) == bytes(ref_lib.camera_update(zl.ZlCamera, pos, ang, fov, asp, lens, foc))


def test_environment_tables_equal_reference(oracle, zl):
    """EnvironmentMap ctor (EnvironmentMap.cpp:8-59, 66-114): both alias levels, the int-truncated sum, the RGB16F texels"""
    from zillumgl_b200 import _native as N
    rng = np.random.default_rng(5)
    for w, h in ((64, 32), (33, 17), (1, 1), (128, 64)):
        img = (rng.random((h, w, 3)) ** 4 * 50).astype(np.float32)
        if w > 8:
            img[h // 3, w // 5] = [30000, 20000, 90000]                      # beyond the binary16 range -> +inf texel
            img[0, 0] = [1e-7, 3e-6, 6.1e-5]                                    # binary16 subnormals
        ra, rp, rs, rtex = ref_lib.env_tables(img, w, h)
        oa, op, os_ = oracle.env_tables(img, w, h)
        assert_same_bits(oa, ra); assert_same_bits(op, rp)
        assert float(int(os_)) == rs                                          # EnvironmentMap.h:22: int sumPdf()
        half = np.array([oracle.lib.zo_round_to_half(float(v)) for v in img.reshape(-1)], np.float32)
        assert_same_bits(half, rtex)
        ha, hp = np.zeros((w + 1) * h, np.int32), np.zeros((w + 1) * h, np.float32)
        hs = N.host.zh_env_tables(img.ctypes.data_as(C.POINTER(C.c_float)), w, h, ha.ctypes.data_as(C.POINTER(C.c_int32)), hp.ctypes.data_as(C.POINTER(C.c_float)))
        assert_same_bits(ha, ra); assert_same_bits(hp, rp)
        assert float(int(hs)) == rs


# ---------------------------------------------------------------------------------------------------------------------
# the shader library, function by function, and the integrators: oracle == reference GLSL
# ---------------------------------------------------------------------------------------------------------------------
def kat_inputs(zl, s, p, rng, n, mats):
    """(op, inputs, nout) for every row of the zl_debug_eval table"""
    out = []
    out.append(("HASH", rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32).reshape(-1, 1), 1))
    out.append(("SOBOL", np.stack([bits(rng.integers(0, 131072, n)), bits(rng.integers(0, 256, n))], axis=1), 1))
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[:64] = np.repeat(np.array([[1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]], np.float32), 16, axis=0) * rng.choice([-1, 1], (64, 3))
    out.append(("CUBEMAP_FACE", d, 1))
    rays = random_rays(s, n, seed=9)
    nb = s.info["bvhSize"]
    out.append(("BOXHIT", np.concatenate([bits(rng.integers(0, nb, n)).reshape(-1, 1), rays], axis=1), 2))
    tri = rng.integers(0, s.info["numTriangles"], n)
    v = s.array("vertices").reshape(-1, 3)[s.array("indices").reshape(-1, 3)[tri]]
    target = np.einsum("nk,nkc->nc", rng.dirichlet([1, 1, 1], n).astype(np.float32), v)
    aim = rays.copy()
    dd = target - aim[:, :3]
    aim[:, 3:] = dd / (np.linalg.norm(dd, axis=1, keepdims=True) + 1e-30)
    out.append(("TRIANGLE", np.concatenate([bits(tri).reshape(-1, 1), aim], axis=1), 2))
    out.append(("SURFACE", np.concatenate([bits(tri).reshape(-1, 1), target.astype(np.float32)], axis=1), 8))
    out.append(("CAMERA_RAY", rng.random((n, 6), dtype=np.float32), 6))
    ref = (rng.random((n, 3), dtype=np.float32) * 2 - 1) + np.array([0, 0, 1], np.float32)
    out.append(("CAMERA_II", np.concatenate([ref, rng.random((n, 2), dtype=np.float32)], axis=1), 10))
    out.append(("CAMERA_PDF", np.concatenate([np.tile(np.array(p.camera.pos, np.float32), (n, 1)), unit(rng, n)], axis=1), 2))
    for mat in mats:
        for mode in (0, 1):
            nrm, wo, wi = unit(rng, n), unit(rng, n), unit(rng, n)
            ev = np.zeros((n, 14), np.float32)
            ev[:, 0] = bits([mat])[0]; ev[:, 1] = bits([-1])[0]
            ev[:, 4:7], ev[:, 7:10], ev[:, 10:13], ev[:, 13] = wo, wi, nrm, bits([mode])[0]
            out.append(("BSDF_EVAL", ev, 4))
            sm = np.zeros((n, 15), np.float32)
            sm[:, 0] = bits([mat])[0]; sm[:, 1] = bits([-1])[0]
            sm[:, 4:7], sm[:, 7:10], sm[:, 10] = wo, nrm, bits([mode])[0]
            sm[:, 11:14] = rng.random((n, 3), dtype=np.float32)
            sm[:, 14] = bits(rng.integers(0, 2 ** 31, n))
            out.append(("BSDF_SAMPLE", sm, 9))
    if s.desc.contents.numTextures > 0:
        ev = np.zeros((n, 14), np.float32)
        ev[:, 0] = bits([0])[0]; ev[:, 1] = bits([0])[0]
        ev[:, 2:4] = rng.random((n, 2), dtype=np.float32) * 20 - 5
        ev[:, 4:7] = ev[:, 7:10] = ev[:, 10:13] = np.array([0, 0, 1], np.float32)
        out.append(("BSDF_EVAL", ev, 4))
    out.append(("ENV_LE", unit(rng, n), 4))
    out.append(("ENV_SAMPLE", rng.random((n, 4), dtype=np.float32), 4))
    nl = s.info["nLightTriangles"]
    if nl:
        lid = rng.integers(0, nl, n)
        out.append(("LIGHT_SAMPLE_LE", np.concatenate([bits(lid).reshape(-1, 1), rng.random((n, 4), dtype=np.float32)], axis=1), 11))
        b = s.array("bounds").reshape(-1, 6)[0]
        x = (b[:3] + rng.random((n, 3), dtype=np.float32) * (b[3:] - b[:3])).astype(np.float32)
        y = (b[:3] + rng.random((n, 3), dtype=np.float32) * (b[3:] - b[:3])).astype(np.float32)
        out.append(("LIGHT_LE", np.concatenate([bits(lid).reshape(-1, 1), y, unit(rng, n), x], axis=1), 4))
        out.append(("SAMPLE_LIGHT_ENV", np.concatenate([x, rng.random((n, 5), dtype=np.float32)], axis=1), 7))
    return out


KAT_SCENES = [("cornell", 64, 48, [0, 1, 3, 4]), ("default", 64, 36, [1, 2]), ("sponza_light", 64, 36, [0, 5, 8, 9]), ("rungholt_small", 64, 36, [0])]


@pytest.mark.parametrize("name,w,h,mats", KAT_SCENES)
def test_oracle_kats_equal_reference_glsl(name, w, h, mats, zl):
    """every function of math / random / intersection / camera / microfacet / material / material_loader / light.glsl that the
    KAT table exposes: the oracle's restatement against the reference's own text"""
    s, o = get_scene(name, w, h)
    r = ref_scene(name, w, h)
    p = params(zl, s, w, h, envRotation=0.7)
    p.camera.lensRadius, p.camera.focalDist = 0.05, 3.0
    for op, inp, nout in kat_inputs(zl, s, p, np.random.default_rng(11), 4096, mats):
        assert_same_bits(o.debug_eval(p, zl.KAT[op], inp, nout), r.debug_eval(p, zl.KAT[op], inp, nout), (name, op))


def random_materials(rng, count):
    """(count, 16) float32 material records (Material.h:32-53): every type, parameters uniform in their ranges with the range ends
    and near-singular values (roughness 0 / 1e-3, ior close to 1, black base colour) mixed in"""
    m = np.zeros((count, 16), np.float32)
    u = rng.random((count, 13)).astype(np.float32)
    ends = rng.random((count, 13))
    u = np.where(ends < 0.12, np.float32(0.0), np.where(ends < 0.24, np.float32(1.0), np.where(ends < 0.3, np.float32(1e-3), u))).astype(np.float32)
    m[:, :12] = u[:, :12]
    m[:, 12] = np.where(ends[:, 12] < 0.15, np.float32(1.0001), np.float32(1.01) + u[:, 12] * np.float32(1.5))      # ior
    m[:, 13] = (np.arange(count) % 5).astype(np.int32).view(np.float32)                                                # type bits
    return m


def scene_with_materials(zl, name, w, h, mats):
    """a private copy of a builtin scene whose flattened material records are overwritten in place (the arrays of the session-cached
    scenes are left alone); the oracle, the reference and the upload all read them through the same ZlSceneDesc"""
    s = zl.Scene.builtin(name, w, h)
    s.flatten()
    d = s.desc.contents
    assert mats.shape[0] == d.numMaterials
    C.memmove(C.cast(d.materials, C.c_void_p), np.ascontiguousarray(mats, np.float32).ctypes.data, mats.size * 4)
    assert np.array_equal(s.array("materials").view(np.uint32), mats.reshape(-1).view(np.uint32))
    return s


def bsdf_kat_inputs(rng, n, mats):
    """BSDF_EVAL / BSDF_SAMPLE rows for the given material indices, both transport modes; a tenth of the directions lie in or close to
    the tangent plane or along the normal (the guards of material.glsl:85-555)"""
    out = []
    for mat in mats:
        for mode in (0, 1):
            nrm, wo, wi = unit(rng, n), unit(rng, n), unit(rng, n)
            k = n // 10
            wi[:k] = nrm[:k]; wo[k:2 * k] = -wi[k:2 * k]
            t = np.cross(nrm[2 * k:3 * k], wo[2 * k:3 * k]); wi[2 * k:3 * k] = t / (np.linalg.norm(t, axis=1, keepdims=True) + np.float32(1e-20))
            ev = np.zeros((n, 14), np.float32)
            ev[:, 0] = bits([mat])[0]; ev[:, 1] = bits([-1])[0]
            ev[:, 4:7], ev[:, 7:10], ev[:, 10:13], ev[:, 13] = wo, wi, nrm, bits([mode])[0]
            out.append(("BSDF_EVAL", ev, 4))
            sm = np.zeros((n, 15), np.float32)
            sm[:, 0] = bits([mat])[0]; sm[:, 1] = bits([-1])[0]
            sm[:, 4:7], sm[:, 7:10], sm[:, 10] = wo, nrm, bits([mode])[0]
            sm[:, 11:14] = rng.random((n, 3), dtype=np.float32)
            sm[:, 14] = bits(rng.integers(0, 2 ** 31, n))
            out.append(("BSDF_SAMPLE", sm, 9))
    return out


@pytest.mark.parametrize("seed", range(3))
def test_oracle_bsdfs_with_random_materials_equal_reference_glsl(seed, zl, oracle):
    """material.glsl / microfacet.glsl over the whole parameter space rather than the builtin scenes' few materials: every record of
    the scene replaced by a random one (all five types), evaluation, pdf and sampling in both transport modes"""
    import oracle_lib
    rng = np.random.default_rng(300 + seed)
    w, h = 48, 27
    probe, _ = get_scene("sponza_light", w, h)
    mats = random_materials(rng, probe.info["numMaterials"])
    s = scene_with_materials(zl, "sponza_light", w, h, mats)
    o, r = oracle_lib.OracleScene(s.desc), ref_lib.RefScene(s.desc)
    p = params(zl, s, w, h)
    for op, inp, nout in bsdf_kat_inputs(rng, 2048, range(mats.shape[0])):
        assert_same_bits(o.debug_eval(p, zl.KAT[op], inp, nout), r.debug_eval(p, zl.KAT[op], inp, nout), (seed, op, int(inp[0, 0].view(np.int32))))


@pytest.mark.parametrize("seed", range(2))
def test_oracle_integrators_with_random_materials_equal_reference_glsl(seed, zl, oracle):
    """the three integrators on a scene whose every object carries a random material (all five BSDF types in one image: refraction
    with its eta scale, delta lobes, Russian roulette on coloured throughput): films identical"""
    import oracle_lib
    rng = np.random.default_rng(400 + seed)
    w, h = 48, 27
    probe, _ = get_scene("sponza_light", w, h)
    s = scene_with_materials(zl, "sponza_light", w, h, random_materials(rng, probe.info["numMaterials"]))
    o, r = oracle_lib.OracleScene(s.desc), ref_lib.RefScene(s.desc)
    for kw in (dict(), dict(russianRoulette=1, maxDepth=7), dict(sampler=0, sampleLight=0)):
        fo, fr = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
        for k in range(2):
            q = params(zl, s, w, h, spp=k, freeCounter=k + 1, **kw)
            o.path_pass(q, fo); r.path_pass(q, fr)
        assert_same_bits(fo[..., :3], fr[..., :3], (seed, "path", kw))
        assert (fo[..., :3] > 0).mean() > 0.2
    oracle.lib.zo_set_threads(1); ref_lib.set_threads(1)
    try:
        for kind in ("light", "triple"):
            fo, fr = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
            for k in range(2):
                q = params(zl, s, w, h, spp=k, freeCounter=k + 1, russianRoulette=k, maxDepth=5)
                q.blocksOnePass, q.loopsPerPass = 1, 1
                q.scale = w * h / 1536.0
                if kind == "light":
                    o.light_pass(q, fo); r.light_pass(q, fr)
                else:
                    o.triple_pt_pass(q, fo); r.triple_pt_pass(q, fr)
                    o.triple_lpt_pass(q, fo); r.triple_lpt_pass(q, fr)
            assert_same_bits(fo[..., :3], fr[..., :3], (seed, kind))
    finally:
        import os
        oracle.lib.zo_set_threads(os.cpu_count()); ref_lib.set_threads(os.cpu_count())


@pytest.mark.parametrize("name,w,h", SCENES)
def test_oracle_traversal_equals_reference_glsl(name, w, h, zl):
    """bvhHit / bvhTest (intersection.glsl:367-427) on the §8(d) ray mix (5 % axis-parallel, 5 % near-zero component) and on the
    camera's pixel-centre rays: ids and distances"""
    s, o = get_scene(name, w, h)
    r = ref_scene(name, w, h)
    rays = random_rays(s, 1 << 17, seed=21)
    oi, ot = o.trace_rays(rays)
    ri, rt = r.trace_rays(rays)
    assert_same_bits(oi, ri); assert_same_bits(ot, rt)
    assert 0.02 < (ri >= 0).mean() < 0.999
    tm = np.where(ri >= 0, rt * np.float32(0.999), np.float32(1e8)).astype(np.float32)
    assert_same_bits(o.trace_rays(rays, anyhit=True, tmax=tm)[0], r.trace_rays(rays, anyhit=True, tmax=tm)[0])
    tm = (rt * np.float32(1.001)).astype(np.float32)
    assert_same_bits(o.trace_rays(rays, anyhit=True, tmax=tm)[0], r.trace_rays(rays, anyhit=True, tmax=tm)[0])


def soup_scene(zl, seed, tmp_path, w=40, h=24):
    """conftest.random_soup(seed) as a scene: written as an OBJ file, read back by host/Model.cpp (generated normals), lit by one
    square lamp above its bounding box"""
    v, idx = random_soup(seed)
    with open(tmp_path / "soup.obj", "w") as f:
        f.write("".join(f"v {float(a)!r} {float(b)!r} {float(c)!r}\n" for a, b, c in v))
        f.write("".join(f"f {a + 1} {b + 1} {c + 1}\n" for a, b, c in idx.reshape(-1, 3)))
    lo, hi = v.min(0), v.max(0)
    c, e = (lo + hi) / 2, float(max((hi - lo).max(), 1e-2))
    xml = (f'<?xml version="1.0"?>\n<scene name="soup{seed}">\n  <integrator type="path"><maxBounce value="4"/><size width="{w}" height="{h}"/></integrator>\n'
           f'  <sampler type="sobol"/>\n  <camera type="thinLens"><position value="{c[0]!r} {c[2] + 2 * e!r} {-c[1]!r}"/><angle value="-90 0 0"/><fov value="50"/>'
           f'<lensRadius value="0"/><focalDistance value="1"/></camera>\n  <modelInstances>\n'
           f'    <modelInstance path="{tmp_path / "soup.obj"}" name="soup" type="object"><transform translate="0 0 0" scale="1 1 1" rotate="0 0 0"/><material type="default"/></modelInstance>\n'
           f'    <modelInstance path="builtin:square" name="lamp" type="light"><transform translate="{c[0]!r} {-c[2]!r} {hi[1] + e!r}" scale="{e!r} {e!r} 1" rotate="180 0 0"/>'
           f'<radiance value="40 40 40"/></modelInstance>\n  </modelInstances>\n</scene>\n')
    (tmp_path / "scene.xml").write_text(xml)
    s = zl.Scene.from_file(tmp_path / "scene.xml")
    s.flatten()
    assert s.info["objPrimCount"] == idx.size // 3
    return s


@pytest.mark.parametrize("seed", range(8))
def test_oracle_on_random_soups_equals_reference_glsl(seed, zl, tmp_path):
    """bvhHit / bvhTest and the path tracer on the trees of degenerate geometry (duplicated triangles, slivers, grid-snapped
    vertices with many exact ties, height fields with shared edges): oracle == reference GLSL"""
    import oracle_lib
    w, h = 40, 24
    s = soup_scene(zl, seed, tmp_path, w, h)
    o, r = oracle_lib.OracleScene(s.desc), ref_lib.RefScene(s.desc)
    rays = random_rays(s, 1 << 15, seed=31 + seed)
    # every second ray is aimed at a point of a random triangle (corners, edges and interiors), so that sparse soups are hit too
    rng = np.random.default_rng(77 + seed)
    tri = s.array("vertices").reshape(-1, 3)[s.array("indices").reshape(-1, 3)[rng.integers(0, s.info["numTriangles"], rays.shape[0] // 2)]]
    bary = rng.dirichlet([0.3, 0.3, 0.3], tri.shape[0]).astype(np.float32)
    dd = np.einsum("nk,nkc->nc", bary, tri) - rays[::2, :3]
    rays[::2, 3:] = dd / (np.linalg.norm(dd, axis=1, keepdims=True) + np.float32(1e-30))
    oi, ot = o.trace_rays(rays)
    ri, rt = r.trace_rays(rays)
    assert_same_bits(oi, ri); assert_same_bits(ot, rt)
    assert (ri >= 0).mean() > 0.3
    tm = np.where(ri >= 0, rt * np.float32(0.999), np.float32(1e8)).astype(np.float32)
    assert_same_bits(o.trace_rays(rays, anyhit=True, tmax=tm)[0], r.trace_rays(rays, anyhit=True, tmax=tm)[0])
    fo, fr = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
    for k in range(2):
        q = params(zl, s, w, h, spp=k, freeCounter=k + 1)
        o.path_pass(q, fo); r.path_pass(q, fr)
    assert_same_bits(fo[..., :3], fr[..., :3])


@pytest.mark.parametrize("name,w,h", SCENES)
@pytest.mark.parametrize("kw", [dict(), dict(russianRoulette=1, maxDepth=6), dict(sampleLight=0), dict(sampler=0), dict(lightEnvUniformSample=1, lightPortion=0.3)])
def test_oracle_path_tracer_equals_reference_glsl(name, w, h, kw, zl):
    """path_integ_naive.glsl:35-174, three passes: the films are identical"""
    s, o = get_scene(name, w, h)
    r = ref_scene(name, w, h)
    fo, fr = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
    for k in range(3):
        p = params(zl, s, w, h, spp=k, freeCounter=k + 1, **kw)
        o.path_pass(p, fo); r.path_pass(p, fr)
    assert_same_bits(fo[..., :3], fr[..., :3])
    assert fo[..., :3].mean() > 1e-4


@pytest.mark.parametrize("name,w,h", SCENES)
def test_oracle_light_and_triple_tracers_equal_reference_glsl(name, w, h, zl, oracle):
    """light_path_integ.glsl, triple_path_pass_{pt,lpt}.glsl.  Splats are float atomics; with one thread both sides add them in
    invocation order, so even these films are identical."""
    s, o = get_scene(name, w, h)
    r = ref_scene(name, w, h)
    oracle.lib.zo_set_threads(1); ref_lib.set_threads(1)
    try:
        for kind, kw in (("light", dict()), ("light", dict(russianRoulette=1, maxDepth=6)), ("triple", dict()), ("triple", dict(russianRoulette=1))):
            fo, fr = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
            for k in range(2):
                p = params(zl, s, w, h, spp=k, freeCounter=k + 1, **kw)
                p.blocksOnePass, p.loopsPerPass = 1, (2 if kind == "triple" else 1)
                p.scale = w * h / (p.blocksOnePass * p.loopsPerPass * 1536.0)
                if kind == "light":
                    o.light_pass(p, fo); r.light_pass(p, fr)
                else:
                    o.triple_pt_pass(p, fo); r.triple_pt_pass(p, fr)
                    o.triple_lpt_pass(p, fo); r.triple_lpt_pass(p, fr)
            assert_same_bits(fo[..., :3], fr[..., :3], (name, kind, kw))
            assert fo[..., :3].sum() > 0
    finally:
        oracle.lib.zo_set_threads(0); ref_lib.set_threads(ref_lib.threads())
        import os
        oracle.lib.zo_set_threads(os.cpu_count()); ref_lib.set_threads(os.cpu_count())


def test_oracle_post_proc_equals_reference_glsl(oracle):
    rng = np.random.default_rng(7)
    film = (rng.random((40, 60, 4)) ** 3 * 8).astype(np.float32)
    film[3, 4, :3] = [-1.0, 0.0, 1e35]                                     # the clamp(color, 0, 1e30) branch (post_proc.glsl:45)
    for tm in (0, 1, 2):
        a, _ = oracle.post_proc(film, 0.37, tm)
        assert_same_bits(a[..., :3], ref_lib.post_proc(film, 0.37, tm)[..., :3])


# ---------------------------------------------------------------------------------------------------------------------
# whole scenes through the reference's own Scene and Integrator classes
# ---------------------------------------------------------------------------------------------------------------------
def load_in_reference(zl, name, w, h):
    """The product's Scene and the reference's Scene over the same model files, XML text, env image and seed image."""
    s = zl.Scene.builtin(name, w, h)
    s.flatten()
    ref_lib.full_reset()
    for i, im in enumerate(zl.Scene.images()):
        ref_lib.register_image(f"mem:tex{i}", im.shape[1], im.shape[0], rgb8=im)
    seen = set()
    for m in s.models():
        if m["path"] in seen:
            continue
        seen.add(m["path"])
        meshes = [dict(pos=x["pos"], nrm=x["nrm"], tex=x["tex"], idx=x["idx"], matIndex=x["matIndex"],
                       texture=(f"mem:tex{x['texIndex']}" if x["texIndex"] >= 0 else "")) for x in m["meshes"]]
        ref_lib.register_model(m["path"], meshes, m["materials"])
    d = s.desc.contents
    xml = s.builtin_xml(name, w, h)
    env = re.search(r'<envMap path="([^"]*)"', xml)
    if env and d.envW > 0:
        ref_lib.register_image(env.group(1), d.envW, d.envH, rgb_float=s.array("envMap"))
    else:
        ref_lib.register_image("", 1, 1, rgb_float=np.zeros(3, np.float32))    # the reference needs an env map (EnvironmentMap.cpp:12-15): 1x1 black
    return s, ref_lib.FullScene(xml, noise=s.array("noise"))


@pytest.mark.parametrize("name,w,h", SCENES)
def test_scene_flatten_equals_reference_scene(name, w, h, zl):
    """Scene::load (Scene.cpp:58-127: XML, transforms, material overrides, lights, camera) + Scene::createGLContext
    (Scene.cpp:133-270: world-space flatten via glm, BVH, light table, uploads): every array the kernels read"""
    s, f = load_in_reference(zl, name, w, h)
    fi, hi, d = f.info, s.info, s.desc.contents
    for k in ("numVertices", "numTriangles", "bvhSize", "objPrimCount", "nLightTriangles", "numMaterials", "filmWidth", "filmHeight", "sampler", "numTextures"):
        assert fi[k] == hi[k], k
    assert np.float32(fi["lightSum"]) == np.float32(d.lightSum) and np.float32(fi["envSum"]) == np.float32(d.envSum)
    for a in ("vertices", "normals", "texcoords", "indices", "bounds", "hitTable", "matTexIndices", "materials", "lightPower", "lightAlias",
              "lightProb", "texUVScale", "texels", "noise"):
        assert_same_bits(f.array(a), s.array(a), (name, a))
    if d.envW > 0:
        assert_same_bits(f.array("envAlias"), s.array("envAlias")); assert_same_bits(f.array("envAliasProb"), s.array("envAliasProb"))
    assert bytes(f.camera(zl.ZlCamera)) == bytes(s.camera())


@pytest.mark.parametrize("seed", range(6))
def test_random_scene_xml_equals_reference_scene(seed, zl, tmp_path):
    """scene.xml files with random content through both loaders: several instances of the same model with their own transforms
    (non-uniform scale -> inverse-transpose normal matrix, all three rotation angles, Model.cpp:62-72), every material type as an
    override (with tags the loader ignores), several lights of different power, a rotated thin-lens camera, either sampler."""
    rng = np.random.default_rng(500 + seed)
    w, h = 40, 24
    f3 = lambda lo, hi: " ".join(repr(float(np.float32(x))) for x in rng.uniform(lo, hi, 3))
    f1 = lambda lo, hi: repr(float(np.float32(rng.uniform(lo, hi))))
    mats = ['<material type="default"/>',
            f'<material type="lambertian"><baseColor value="{f3(0, 1)}"/></material>',
            f'<material type="principled"><baseColor value="{f3(0, 1)}"/><subsurface value="{f1(0, 1)}"/><metallic value="{f1(0, 1)}"/><roughness value="{f1(0, 1)}"/>'
            f'<specular value="{f1(0, 1)}"/><specularTint value="{f1(0, 1)}"/><sheen value="{f1(0, 1)}"/><sheenTint value="{f1(0, 1)}"/><clearcoat value="{f1(0, 1)}"/>'
            f'<clearcoatGloss value="{f1(0, 1)}"/><albedo value="1 0 0"/></material>',
            f'<material type="metalWorkflow"><baseColor value="{f3(0, 1)}"/><metallic value="{f1(0, 1)}"/><roughness value="{f1(0, 1)}"/></material>',
            f'<material type="dielectric"><baseColor value="{f3(0, 1)}"/><ior value="{f1(1.1, 2.2)}"/><roughness value="{f1(0, 0.5)}"/><tint value="0 1 0"/></material>',
            f'<material type="thinDielectric"><baseColor value="{f3(0, 1)}"/><ior value="{f1(1.1, 2.2)}"/></material>',
            '<material type="noSuchType"><baseColor value="0.1 0.2 0.3"/></material>']
    models = ["builtin:cube", "builtin:sphere", "builtin:teapotBody", "builtin:teapotCap", "builtin:square", "builtin:cornell"]
    inst = []
    for k in range(int(rng.integers(4, 9))):
        inst.append(f'<modelInstance path="{models[int(rng.integers(0, len(models)))]}" name="obj{k}" type="object"><transform translate="{f3(-6, 6)}" '
                    f'scale="{f3(0.2, 3)}" rotate="{f3(-180, 180)}"/>{mats[int(rng.integers(0, len(mats)))]}</modelInstance>')
    for k in range(int(rng.integers(1, 4))):
        inst.append(f'<modelInstance path="{models[int(rng.choice([0, 4, 5, 1]))]}" name="light{k}" type="light"><transform translate="{f3(-6, 6)}" scale="{f3(0.5, 2)}" '
                    f'rotate="{f3(-180, 180)}"/><radiance value="{f3(1, 50)}"/></modelInstance>')
    xml = (f'<?xml version="1.0"?>\n<scene name="random{seed}">\n  <integrator type="path"><maxBounce value="4"/><size width="{w}" height="{h}"/></integrator>\n'
           f'  <sampler type="{"sobol" if seed % 2 else "independent"}"/>\n  <camera type="thinLens"><position value="{f3(-9, 9)}"/><angle value="{f3(-180, 180)}"/>'
           f'<fov value="{f1(20, 90)}"/><lensRadius value="{f1(0, 0.2)}"/><focalDistance value="{f1(1, 8)}"/></camera>\n  <modelInstances>\n    '
           + "\n    ".join(inst) + "\n  </modelInstances>\n</scene>\n")
    path = tmp_path / "scene.xml"
    path.write_text(xml)
    s = zl.Scene.from_file(path)
    s.flatten()
    # the model files as the reference's importer would deliver them: every model once, with its OWN materials (an instance that
    # carries a <material> override reports the overridden ones)
    plain = tmp_path / "models.xml"
    plain.write_text(re.sub(r"<modelInstances>.*</modelInstances>", "<modelInstances>" + "".join(
        f'<modelInstance path="{m}" name="m" type="object"><transform translate="0 0 0" scale="1 1 1" rotate="0 0 0"/><material type="default"/></modelInstance>'
        for m in models) + "</modelInstances>", xml, flags=re.S))
    ref_lib.full_reset()
    for m in zl.Scene.from_file(plain).models():
        ref_lib.register_model(m["path"], [dict(pos=x["pos"], nrm=x["nrm"], tex=x["tex"], idx=x["idx"], matIndex=x["matIndex"], texture="") for x in m["meshes"]],
                               m["materials"])
    ref_lib.register_image("", 1, 1, rgb_float=np.zeros(3, np.float32))
    f = ref_lib.FullScene(xml, noise=s.array("noise"))
    fi, hi, d = f.info, s.info, s.desc.contents
    for k in ("numVertices", "numTriangles", "bvhSize", "objPrimCount", "nLightTriangles", "numMaterials", "filmWidth", "filmHeight", "sampler"):
        assert fi[k] == hi[k], k
    assert np.float32(fi["lightSum"]) == np.float32(d.lightSum)
    for a in ("vertices", "normals", "texcoords", "indices", "bounds", "hitTable", "matTexIndices", "materials", "lightPower", "lightAlias", "lightProb"):
        assert_same_bits(f.array(a), s.array(a), (seed, a))
    assert bytes(f.camera(zl.ZlCamera)) == bytes(s.camera())
    # and the shaders on that scene (several emitters of different power and orientation, a lens with an aperture): every KAT row
    # and a few passes, the oracle against the reference's GLSL
    import oracle_lib
    o, r = oracle_lib.OracleScene(s.desc), ref_lib.RefScene(s.desc)
    q = params(zl, s, w, h, envRotation=0.3, sampler=s.info["sampler"])
    for op, inp, nout in kat_inputs(zl, s, q, np.random.default_rng(600 + seed), 1024, range(min(hi["numMaterials"], 4))):
        assert_same_bits(o.debug_eval(q, zl.KAT[op], inp, nout), r.debug_eval(q, zl.KAT[op], inp, nout), (seed, op))
    fo, fr = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
    for k in range(2):
        q = params(zl, s, w, h, spp=k, freeCounter=k + 1, sampler=s.info["sampler"], russianRoulette=k)
        o.path_pass(q, fo); r.path_pass(q, fr)
    assert_same_bits(fo[..., :3], fr[..., :3], (seed, "path film"))
    oracle_lib.lib.zo_set_threads(1); ref_lib.set_threads(1)                 # one thread: splats add in invocation order on both sides
    try:
        for kind in ("light", "triple"):
            fo, fr = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
            for k in range(2):
                q = params(zl, s, w, h, spp=k, freeCounter=k + 1, russianRoulette=1 - k, maxDepth=5, sampler=s.info["sampler"])
                q.blocksOnePass, q.loopsPerPass = 1, 1
                q.scale = w * h / 1536.0
                if kind == "light":
                    o.light_pass(q, fo); r.light_pass(q, fr)
                else:
                    o.triple_pt_pass(q, fo); r.triple_pt_pass(q, fr)
                    o.triple_lpt_pass(q, fo); r.triple_lpt_pass(q, fr)
            assert_same_bits(fo[..., :3], fr[..., :3], (seed, kind))
    finally:
        import os
        oracle_lib.lib.zo_set_threads(os.cpu_count()); ref_lib.set_threads(os.cpu_count())


@pytest.mark.parametrize("ew,eh,hot", [(33, 17, False), (7, 5, False), (64, 32, True), (1, 1, False)])
def test_environment_maps_of_odd_sizes_equal_reference_glsl(ew, eh, hot, zl, tmp_path):
    """light.glsl:163-219 on environment maps that are not the builtin sky: odd sizes (bilinear + repeat across the seam and the
    poles, the (w + 1)-wide two-level alias table), binary16 subnormals, and with `hot` a texel beyond the binary16 range (+inf
    radiance, NaN weights downstream): envLe / envSampleLi / sampleLightAndEnv rows and env-lit path-tracer films, the oracle
    against the reference's GLSL; the map goes in as a PFM file through the product's loader."""
    import oracle_lib
    rng = np.random.default_rng(ew * 100 + eh)
    env = (rng.random((eh, ew, 3)) ** 4 * 30).astype(np.float32)
    if ew > 4:
        env[eh // 2, ew // 3] = [2e-7, 3e-6, 6.1e-5]
    if hot:
        env[eh // 3, ew // 5] = [30000, 20000, 90000]
    with open(tmp_path / "env.pfm", "wb") as f:
        f.write(b"PF\n%d %d\n-1.0\n" % (ew, eh)); f.write(np.ascontiguousarray(env[::-1], "<f4").tobytes())
    w, h = 32, 20
    xml = (f'<?xml version="1.0"?>\n<scene name="env">\n  <integrator type="path"><maxBounce value="3"/><size width="{w}" height="{h}"/></integrator>\n  <sampler type="sobol"/>\n'
           '  <camera type="thinLens"><position value="0 -6 2"/><angle value="0 -10 0"/><fov value="50"/><lensRadius value="0"/><focalDistance value="1"/></camera>\n  <modelInstances>\n'
           '    <modelInstance path="builtin:teapotBody" name="pot" type="object"><transform translate="0 0 0" scale="1 1 1" rotate="0 0 0"/><material type="principled">'
           '<baseColor value="0.8 0.5 0.3"/><metallic value="0.3"/><roughness value="0.4"/><clearcoat value="0.5"/></material></modelInstance>\n'
           '    <modelInstance path="builtin:square" name="floor" type="object"><transform translate="0 0 0" scale="8 8 1" rotate="0 0 0"/><material type="default"/></modelInstance>\n'
           '    <modelInstance path="builtin:square" name="lamp" type="light"><transform translate="0 0 6" scale="1 1 1" rotate="180 0 0"/><radiance value="10 10 10"/></modelInstance>\n'
           f'  </modelInstances>\n  <envMap path="{tmp_path / "env.pfm"}"/>\n</scene>\n')
    (tmp_path / "scene.xml").write_text(xml)
    s = zl.Scene.from_file(tmp_path / "scene.xml")
    s.flatten()
    d = s.desc.contents
    assert (d.envW, d.envH) == (ew, eh) and np.array_equal(s.array("envMap").reshape(eh, ew, 3), env)
    o, r = oracle_lib.OracleScene(s.desc), ref_lib.RefScene(s.desc)
    n = 4096
    b = s.array("bounds").reshape(-1, 6)[0]
    for rot in (0.0, 2.1):
        q = params(zl, s, w, h, envRotation=rot)
        x = (b[:3] + rng.random((n, 3), dtype=np.float32) * (b[3:] - b[:3])).astype(np.float32)
        dirs = unit(rng, n); dirs[:64] = np.array([0, 0, 1], np.float32); dirs[64:128] = np.array([0, 0, -1], np.float32); dirs[128:192, 1] = 0     # poles, the seam
        dirs[128:192] /= np.linalg.norm(dirs[128:192], axis=1, keepdims=True)
        for op, inp, nout in (("ENV_LE", dirs, 4), ("ENV_SAMPLE", rng.random((n, 4), dtype=np.float32), 4),
                              ("SAMPLE_LIGHT_ENV", np.concatenate([x, rng.random((n, 5), dtype=np.float32)], axis=1), 7)):
            assert_same_bits(o.debug_eval(q, zl.KAT[op], inp, nout), r.debug_eval(q, zl.KAT[op], inp, nout), (ew, eh, rot, op))
    for kw in (dict(), dict(lightEnvUniformSample=1, lightPortion=0.3), dict(sampleLight=0, envRotation=1.0)):
        fo, fr = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
        for k in range(2):
            q = params(zl, s, w, h, spp=k, freeCounter=k + 1, **kw)
            o.path_pass(q, fo); r.path_pass(q, fr)
        assert_same_bits(fo[..., :3], fr[..., :3], (ew, eh, kw))


def test_texture_layers_of_different_sizes_equal_reference(zl, tmp_path):
    """Albedo textures of different sizes in one array (Texture.cpp:134-171: layers padded to the largest image, per-layer uv scale;
    material.glsl's textured base colour: GL_SRGB decode, LINEAR filter, REPEAT wrap): an OBJ + MTL + PNG model through the
    product's readers and the same images / meshes through the reference's Scene, then the textured BSDF rows at uv far outside
    [0, 1] and exactly on texel boundaries."""
    PIL = pytest.importorskip("PIL.Image")
    import oracle_lib
    rng = np.random.default_rng(8)
    sizes = {"a": (64, 32), "b": (16, 48)}
    for k, (tw, th) in sizes.items():
        PIL.fromarray(rng.integers(0, 256, (th, tw, 3), dtype=np.uint8)).save(tmp_path / f"tex_{k}.png")
    (tmp_path / "quads.mtl").write_text("newmtl ma\nKd 0.8 0.7 0.6\nmap_Kd tex_a.png\nnewmtl mb\nKd 0.2 0.9 0.4\nmap_Kd -s 1 1 1 tex_b.png\nnewmtl mc\nKd 0.5 0.5 0.5\n")
    (tmp_path / "quads.obj").write_text("mtllib quads.mtl\nv -1 0 -1\nv 1 0 -1\nv 1 0 1\nv -1 0 1\nv 2 0 -1\nv 4 0 -1\nv 4 0 1\nv 2 0 1\nv 5 0 -1\nv 7 0 -1\nv 7 0 1\n"
                                        "vt 0 0\nvt 3 0\nvt 3 2\nvt 0 2\nvn 0 1 0\n"
                                        "usemtl ma\nf 1/1/1 2/2/1 3/3/1 4/4/1\nusemtl mb\nf 5/1/1 6/2/1 7/3/1 8/4/1\nusemtl mc\nf 9/1/1 10/2/1 11/3/1\n")
    w, h = 32, 20
    xml = (f'<?xml version="1.0"?>\n<scene name="layers">\n  <integrator type="path"><maxBounce value="3"/><size width="{w}" height="{h}"/></integrator>\n  <sampler type="sobol"/>\n'
           '  <camera type="thinLens"><position value="3 -1 6"/><angle value="0 -80 0"/><fov value="60"/><lensRadius value="0"/><focalDistance value="1"/></camera>\n  <modelInstances>\n'
           '    <modelInstance path="builtin:cornell" name="first" type="object"><transform translate="0 9 0" scale="1 1 1" rotate="0 0 30"/><material type="default"/></modelInstance>\n'
           f'    <modelInstance path="{tmp_path / "quads.obj"}" name="quads" type="object"><transform translate="0 0 0" scale="1 1 1" rotate="0 0 0"/><material type="default"/></modelInstance>\n'
           '    <modelInstance path="builtin:square" name="lamp" type="light"><transform translate="3 0 5" scale="3 3 1" rotate="180 0 0"/><radiance value="30 30 30"/></modelInstance>\n'
           '  </modelInstances>\n</scene>\n')
    (tmp_path / "scene.xml").write_text(xml)
    s = zl.Scene.from_file(tmp_path / "scene.xml")
    s.flatten()
    d = s.desc.contents
    images = zl.Scene.images()
    assert d.numTextures == len(images) >= 2 and d.texMaxW >= 64 and d.texMaxH >= 48
    layers = sorted(set(int(x) >> 16 for x in s.array("matTexIndices")))
    assert len(layers) == 3 and layers[0] == -1                      # two textured meshes (their own layers) + untextured ones (-1)
    assert int(s.array("matTexIndices").max() & 0xffff) >= 3       # the textured meshes come after another model: their material ids are offset
    ref_lib.full_reset()
    for i, im in enumerate(images):
        ref_lib.register_image(f"mem:tex{i}", im.shape[1], im.shape[0], rgb8=im)
    for m in s.models():
        ref_lib.register_model(m["path"], [dict(pos=x["pos"], nrm=x["nrm"], tex=x["tex"], idx=x["idx"], matIndex=x["matIndex"],
                                                texture=(f"mem:tex{x['texIndex']}" if x["texIndex"] >= 0 else "")) for x in m["meshes"]], m["materials"])
    ref_lib.register_image("", 1, 1, rgb_float=np.zeros(3, np.float32))
    f = ref_lib.FullScene(xml, noise=s.array("noise"))
    for a in ("vertices", "normals", "texcoords", "indices", "matTexIndices", "materials", "texUVScale", "texels"):
        assert_same_bits(f.array(a), s.array(a), a)
    o, r = oracle_lib.OracleScene(s.desc), ref_lib.RefScene(s.desc)
    q = params(zl, s, w, h)
    n = 4096
    for layer in layers[1:]:
        tw, th = images[layer].shape[1], images[layer].shape[0]
        ev = np.zeros((n, 14), np.float32)
        ev[:, 0] = bits([0])[0]; ev[:, 1] = bits([layer])[0]
        ev[:, 2:4] = rng.random((n, 2), dtype=np.float32) * 9 - 4
        ev[: n // 4, 2] = (rng.integers(-3 * tw, 3 * tw, n // 4) / np.float32(tw)).astype(np.float32)            # texel boundaries
        ev[: n // 4, 3] = (rng.integers(-3 * th, 3 * th, n // 4) / np.float32(th)).astype(np.float32)
        ev[:, 4:7] = ev[:, 7:10] = ev[:, 10:13] = np.array([0, 0, 1], np.float32)
        a, b = o.debug_eval(q, zl.KAT["BSDF_EVAL"], ev, 4), r.debug_eval(q, zl.KAT["BSDF_EVAL"], ev, 4)
        assert_same_bits(a, b, ("textured eval", layer))
        assert np.unique(a[:, 0]).size > 100                         # the texture really modulates the result
    fo, fr = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
    for k in range(2):
        q = params(zl, s, w, h, spp=k, freeCounter=k + 1)
        o.path_pass(q, fo); r.path_pass(q, fr)
    assert_same_bits(fo[..., :3], fr[..., :3], "film")


@pytest.mark.parametrize("name,w,h", [("cornell", 64, 48), ("sponza_light", 48, 27)])
def test_reference_integrators_end_to_end(name, w, h, zl, oracle):
    """NaivePathIntegrator / LightPathIntegrator / TriplePathIntegrator of the reference (init, reset, updateUniforms,
    renderOnePass, resultScale) dispatching the reference's own shaders: the frame equals the oracle's passes driven with the
    product's parameter sequence, and resultScale() has the reference's off-by-one."""
    s, f = load_in_reference(zl, name, w, h)
    o = oracle.OracleScene(s.desc)
    oracle.lib.zo_set_threads(1); ref_lib.set_threads(1)
    try:
        for kind in ("path", "light", "triple"):
            ri = ref_lib.FullIntegrator(f, kind, w, h)
            if kind == "light":
                ri.set("threadBlocksOnePass", 2)
            if kind == "triple":
                ri.set("LPTBlocksOnePass", 1)
            ref = np.zeros((h, w, 4), np.float32)
            for k in range(3):
                ri.renderOnePass()
                p = params(zl, s, w, h, spp=k, freeCounter=k + 1)
                if kind == "path":
                    o.path_pass(p, ref)
                elif kind == "light":
                    p.blocksOnePass = 2
                    o.light_pass(p, ref)
                else:
                    p.blocksOnePass, p.loopsPerPass, p.scale = 1, 1, w * h / 1536.0
                    o.triple_pt_pass(p, ref); o.triple_lpt_pass(p, ref)
            assert_same_bits(ri.getFrame()[..., :3], ref[..., :3], (name, kind))
            per = {"path": None, "light": 2 * 1536 / (w * h), "triple": None}[kind]
            expect = 1.0 / 4 if per is None else 1.0 / np.float32(np.float32(per) * 4)
            assert ri.resultScale() == pytest.approx(expect, rel=1e-6)
    finally:
        import os
        oracle.lib.zo_set_threads(os.cpu_count()); ref_lib.set_threads(os.cpu_count())


@pytest.mark.parametrize("kind", ["path", "light", "triple"])
def test_product_integrator_glue_follows_the_reference_through_resets(kind, zl, oracle):
    """The product's Integrator classes (host/Integrator.cpp, driven without a device: they only keep the books) beside the
    reference's own NaivePath / LightPath / TriplePath glue dispatching the reference's shaders, through a scripted session:
    passes, a parameter change (the GUI's setShouldReset), a finite-sample render that runs out while renderOnePass keeps being
    called every frame (Application.cpp:644-663), and a restart.  At every call resultScale() agrees, and the reference's frame
    equals the oracle's passes driven with the PRODUCT's uniforms (uSpp, uFreeCounter, depth, block counts, LPT scale)."""
    name, w, h = "sponza_light", 48, 27
    s, f = load_in_reference(zl, name, w, h)
    o = oracle.OracleScene(s.desc)
    cls = {"path": zl.NaivePathIntegrator, "light": zl.LightPathIntegrator, "triple": zl.TriplePathIntegrator}[kind]
    oracle.lib.zo_set_threads(1); ref_lib.set_threads(1)
    try:
        ri, pi = ref_lib.FullIntegrator(f, kind, w, h), cls(s, w, h, host_only=True)
        film = np.zeros((h, w, 4), np.float32)

        def both(name_, value):                 # a GUI edit: the value changes and the integrator restarts
            ri.set(name_, value)
            setattr(pi.mParam, name_, value)
            pi.reset()
            film[:] = 0

        def step():
            before, p0, p1 = pi.curSample, pi.params(0), pi.params(1)
            pi.renderOnePass(); ri.renderOnePass()
            rendered = pi.curSample != before
            if rendered:
                if kind == "path":
                    o.path_pass(p0, film)
                elif kind == "light":
                    o.light_pass(p0, film)
                else:
                    o.triple_pt_pass(p0, film); o.triple_lpt_pass(p1, film)
            assert np.float32(pi.resultScale()) == np.float32(ri.resultScale()), (kind, pi.curSample)
            return rendered

        if kind == "light":
            both("threadBlocksOnePass", 2)
        if kind == "triple":
            both("LPTBlocksOnePass", 1)
        assert [step() for _ in range(3)] == [True] * 3
        assert_same_bits(ri.getFrame()[..., :3], film[..., :3], (kind, "first passes"))
        both("maxDepth", 2)
        assert [step() for _ in range(2)] == [True] * 2
        assert_same_bits(ri.getFrame()[..., :3], film[..., :3], (kind, "after a parameter change"))
        both("maxSample", 6 if kind == "light" else 1)
        both("finiteSample", 1)
        ran = [step() for _ in range(7)]
        assert ran[0] and not ran[-1] and ran == sorted(ran, reverse=True)        # it renders, runs out, and stays finished
        assert_same_bits(ri.getFrame()[..., :3], film[..., :3], (kind, "finite render", ran))
        both("finiteSample", 0)
        assert [step() for _ in range(2)] == [True] * 2
        assert_same_bits(ri.getFrame()[..., :3], film[..., :3], (kind, "restart after the idle frames"))
    finally:
        import os
        oracle.lib.zo_set_threads(os.cpu_count()); ref_lib.set_threads(os.cpu_count())


# ---------------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path against the reference directly
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h,mats", KAT_SCENES)
def test_cuda_kats_equal_reference_glsl(name, w, h, mats, zl):
    s, _ = get_scene(name, w, h)
    if not s.device:
        s.upload()
    r = ref_scene(name, w, h)
    p = params(zl, s, w, h, envRotation=0.7)
    p.camera.lensRadius, p.camera.focalDist = 0.05, 3.0
    for op, inp, nout in kat_inputs(zl, s, p, np.random.default_rng(12), 4096, mats):
        assert_same_bits(zl.debug_eval(s, p, zl.KAT[op], inp, nout), r.debug_eval(p, zl.KAT[op], inp, nout), (name, op))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_cuda_bsdfs_with_random_materials_equal_reference_glsl(seed, zl):
    """the CUDA BSDF library on the random material records of test_oracle_bsdfs_with_random_materials_equal_reference_glsl"""
    rng = np.random.default_rng(300 + seed)
    w, h = 48, 27
    probe, _ = get_scene("sponza_light", w, h)
    mats = random_materials(rng, probe.info["numMaterials"])
    s = scene_with_materials(zl, "sponza_light", w, h, mats)
    s.upload()
    r = ref_lib.RefScene(s.desc)
    p = params(zl, s, w, h)
    for op, inp, nout in bsdf_kat_inputs(rng, 2048, range(mats.shape[0])):
        assert_same_bits(zl.debug_eval(s, p, zl.KAT[op], inp, nout), r.debug_eval(p, zl.KAT[op], inp, nout), (seed, op, int(inp[0, 0].view(np.int32))))


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h", SCENES + [("sponza", 1920, 1080), ("rungholt", 1280, 720)])
def test_cuda_traversal_2p22_rays_equals_reference_glsl(name, w, h, zl):
    """SURVEY §8(d) fixed ray set at its stated size: 2^22 seeded random rays (5 % axis-parallel, 5 % with a component below
    1e-6), first-hit triangle ids and distances, on all scenes including the full Sponza-class (262 k) and Rungholt-class (6.3 M)."""
    if name in ("sponza", "rungholt"):
        s = zl.Scene.builtin(name, w, h); s.flatten()
        r = ref_lib.RefScene(s.desc)
    else:
        s, _ = get_scene(name, w, h)
        r = ref_scene(name, w, h)
    if not s.device:
        s.upload()
    rays = random_rays(s, 1 << 22, seed=77)
    gi, gt = zl.trace_rays(s, rays)
    ri, rt = r.trace_rays(rays)
    assert_same_bits(gi, ri); assert_same_bits(gt, rt)
    assert 0.02 < (ri >= 0).mean() < 0.999
    tm = np.where(ri >= 0, rt * np.float32(0.999), np.float32(1e8)).astype(np.float32)
    assert_same_bits(zl.trace_rays(s, rays[::8], anyhit=True, tmax=tm[::8])[0], r.trace_rays(rays[::8], anyhit=True, tmax=tm[::8])[0])


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h", SCENES)
def test_cuda_path_tracer_equals_reference_glsl(name, w, h, zl):
    s, _ = get_scene(name, w, h)
    if not s.device:
        s.upload()
    r = ref_scene(name, w, h)
    for variant in (0, 2):
        integ = zl.NaivePathIntegrator(s, w, h)
        integ.mParam.kernelVariant = variant
        fr = np.zeros((h, w, 4), np.float32)
        for _ in range(4):
            r.path_pass(integ.params(), fr)
            integ.renderOnePass()
        assert_same_bits(np.ascontiguousarray(integ.getFrame(1.0)[..., :3]), np.ascontiguousarray(fr[..., :3]), (name, variant))
