"""GPU known-answer tests per device function (zl_debug_eval) against the oracle: BIT-EXACT
for every function.  Integer / pure-arithmetic functions use + - * / sqrt only; the functions
that go through sin/cos/pow/log/atan2 take them from include/zl_libm.h, the same code on the
device and on the host, so there is no tolerance anywhere in this file (NaN matches NaN: the
payload of a generated NaN is hardware-defined)."""
import numpy as np
import pytest

from conftest import get_scene, random_rays

pytestmark = pytest.mark.gpu


def _setup(zl, name="cornell", w=64, h=48, **kw):
    s, o = get_scene(name, w, h)
    if not s.device:
        s.upload()
    p = zl.ZlRenderParams()
    p.camera = s.camera(); p.camera.asp = w / h
    p.filmW, p.filmH = w, h
    p.maxDepth, p.sampleLight, p.lightPortion, p.sampler = 4, 1, 0.5, 1
    p.spp, p.freeCounter = 3, 4
    for k, v in kw.items():
        setattr(p, k, v)
    return s, o, p


def _bits(a):
    return np.asarray(a).astype(np.int32).view(np.float32)


def _both(zl, s, o, p, op, inp, nout):
    return zl.debug_eval(s, p, zl.KAT[op], inp, nout), o.debug_eval(p, zl.KAT[op], inp, nout)


def _assert_same_bits(g, r, what=""):
    bad = (g.view(np.uint32) != r.view(np.uint32)) & ~(np.isnan(g) & np.isnan(r))
    assert not bad.any(), (what, int(bad.any(axis=-1).sum()), g[bad.any(axis=-1)][:3], r[bad.any(axis=-1)][:3])


def test_hash_sobol_face_bit_exact(zl):
    s, o, p = _setup(zl)
    rng = np.random.default_rng(0)
    seeds = rng.integers(0, 2 ** 32, 4096, dtype=np.uint64).astype(np.uint32).view(np.float32).reshape(-1, 1)
    g, r = _both(zl, s, o, p, "HASH", seeds, 1)
    assert np.array_equal(g.view(np.uint32), r.view(np.uint32))
    idx = np.stack([_bits(rng.integers(0, 131072, 4096)), _bits(rng.integers(0, 256, 4096))], axis=1)
    g, r = _both(zl, s, o, p, "SOBOL", idx, 1)
    assert np.array_equal(g.view(np.uint32), r.view(np.uint32))
    d = rng.normal(size=(4096, 3)).astype(np.float32)
    d[:64] = np.repeat(np.array([[1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]], np.float32), 16, axis=0) * rng.choice([-1, 1], (64, 3))
    g, r = _both(zl, s, o, p, "CUBEMAP_FACE", d, 1)
    assert np.array_equal(g.view(np.uint32), r.view(np.uint32))


@pytest.mark.parametrize("name,w,h", [("cornell", 64, 48), ("default", 64, 36)])
def test_boxhit_and_triangle_bit_exact(name, w, h, zl):
    s, o, p = _setup(zl, name, w, h)
    rng = np.random.default_rng(5)
    n = 1 << 14
    rays = random_rays(s, n, seed=9)
    # aim half of the rays at a point inside a random node's box so that hits and misses both occur
    nb = s.info["bvhSize"]
    bounds = s.array("bounds").reshape(nb, 6)
    node = rng.integers(0, nb, n)
    inside = bounds[node, :3] + rng.random((n, 3), dtype=np.float32) * (bounds[node, 3:] - bounds[node, :3])
    aimed = rng.random(n) < 0.5
    aimed[: n // 10] = False                      # keep the axis-parallel / near-zero special rays as they are
    d = inside - rays[:, :3]
    d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-30
    rays[aimed, 3:] = d[aimed].astype(np.float32)
    # threaded index of `node` in the hit table of each ray's face (cubemapFace(-dir), math.glsl:112-131)
    nd = -rays[:, 3:]
    ad = np.abs(nd)
    dim = np.where(ad[:, 0] > ad[:, 1], np.where(ad[:, 0] > ad[:, 2], 0, 2), np.where(ad[:, 1] > ad[:, 2], 1, 2))
    face = 2 * dim + (nd[np.arange(n), dim] <= 0)
    table = s.array("hitTable").reshape(6, nb, 3)
    inv = np.empty((6, nb), np.int64)
    for f in range(6):
        inv[f, table[f, :, 0]] = np.arange(nb)
    k = _bits(inv[face, node]).reshape(-1, 1)
    g, r = _both(zl, s, o, p, "BOXHIT", np.concatenate([k, rays], axis=1), 2)
    assert np.array_equal(g.view(np.uint32), r.view(np.uint32))
    assert 0.2 < r[:, 0].mean() < 0.98
    tri = _bits(rng.integers(0, s.info["numTriangles"], n)).reshape(-1, 1)
    v = s.array("vertices").reshape(-1, 3)[s.array("indices").reshape(-1, 3)[tri.view(np.int32)[:, 0]]]
    bary = rng.dirichlet([1, 1, 1], n).astype(np.float32)
    target = np.einsum("nk,nkc->nc", bary, v)
    aim = rays.copy()
    d = target - aim[:, :3]
    aim[:, 3:] = d / np.linalg.norm(d, axis=1, keepdims=True)
    g, r = _both(zl, s, o, p, "TRIANGLE", np.concatenate([tri, aim], axis=1), 2)
    assert np.array_equal(g.view(np.uint32), r.view(np.uint32))
    assert r[:, 0].mean() > 0.5


def test_surface_info(zl):
    # normalize / sqrt / division only: bit-exact
    s, o, p = _setup(zl, "default", 64, 36)
    rng = np.random.default_rng(6)
    n = 4096
    tri = rng.integers(0, s.info["numTriangles"], n)
    v = s.array("vertices").reshape(-1, 3)[s.array("indices").reshape(-1, 3)[tri]]
    pt = np.einsum("nk,nkc->nc", rng.dirichlet([1, 1, 1], n).astype(np.float32), v).astype(np.float32)
    g, r = _both(zl, s, o, p, "SURFACE", np.concatenate([_bits(tri).reshape(-1, 1), pt], axis=1), 8)
    # degenerate (zero-area) triangles of the procedural teapot give 1/0 * 0 = NaN on both sides; the NaN
    # payload/sign is hardware-defined (x86 SSE 0xFFC00000, sm_100 0x7FFFFFFF), so NaN matches NaN
    bad = (g.view(np.uint32) != r.view(np.uint32)) & ~(np.isnan(g) & np.isnan(r))
    assert not bad.any(), (bad.sum(axis=0), np.abs(g - r).max(axis=0), g[bad.any(axis=1)][:3], r[bad.any(axis=1)][:3])


def test_camera_functions(zl):
    s, o, p = _setup(zl, "cornell", 64, 48)
    p.camera.lensRadius, p.camera.focalDist = 0.05, 3.0
    rng = np.random.default_rng(7)
    n = 4096
    inp = rng.random((n, 6), dtype=np.float32)
    g, r = _both(zl, s, o, p, "CAMERA_RAY", inp, 6)
    _assert_same_bits(g, r, "CAMERA_RAY")                   # sin/cos in toConcentricDisk: zl_libm.h on both sides
    ref = (rng.random((n, 3), dtype=np.float32) * 2 - 1) * np.array([1, 1, 1], np.float32) + np.array([0, 0, 1], np.float32)
    g, r = _both(zl, s, o, p, "CAMERA_II", np.concatenate([ref, rng.random((n, 2), dtype=np.float32)], axis=1), 10)
    _assert_same_bits(g, r, "CAMERA_II")
    assert (r[:, 9] > 0).mean() > 0.9
    p.camera.lensRadius = 0.0                                # pinhole: no transcendental on the path -> bit-exact
    g, r = _both(zl, s, o, p, "CAMERA_II", np.concatenate([ref, rng.random((n, 2), dtype=np.float32)], axis=1), 10)
    assert np.array_equal(g.view(np.uint32), r.view(np.uint32))
    rays = np.concatenate([np.tile(np.array(p.camera.pos, np.float32), (n, 1)), o.debug_eval(p, zl.KAT["CAMERA_RAY"], inp, 6)[:, 3:]], axis=1)
    g, r = _both(zl, s, o, p, "CAMERA_PDF", rays, 2)
    assert np.array_equal(g.view(np.uint32), r.view(np.uint32))


@pytest.mark.parametrize("scene,mats", [("cornell", [0, 1, 3, 4]), ("default", [1, 2]), ("sponza_light", [0, 5, 8, 9])])
def test_bsdf_eval_and_sample(scene, mats, zl):
    """Every material type (Lambertian, Principled, MetalWorkflow, rough + delta Dielectric):
    eval/pdf and sample, both transport modes, bit for bit (material.glsl:68-555,
    microfacet.glsl:4-120, material_loader.glsl:99-169)."""
    w, h = (64, 48) if scene == "cornell" else (64, 36)
    s, o, p = _setup(zl, scene, w, h)
    rng = np.random.default_rng(8)
    n = 4096
    for mat in mats:
        def unit(k):
            v = rng.normal(size=(k, 3)).astype(np.float32)
            return v / np.linalg.norm(v, axis=1, keepdims=True)
        nrm, wo, wi = unit(n), unit(n), unit(n)
        for mode in (0, 1):
            ev = np.zeros((n, 14), np.float32)
            ev[:, 0] = _bits([mat])[0]; ev[:, 1] = _bits([-1])[0]
            ev[:, 4:7], ev[:, 7:10], ev[:, 10:13], ev[:, 13] = wo, wi, nrm, _bits([mode])[0]
            g, r = _both(zl, s, o, p, "BSDF_EVAL", ev, 4)
            _assert_same_bits(g, r, (scene, mat, mode, "eval"))
            sm = np.zeros((n, 15), np.float32)
            sm[:, 0] = _bits([mat])[0]; sm[:, 1] = _bits([-1])[0]
            sm[:, 4:7], sm[:, 7:10], sm[:, 10] = wo, nrm, _bits([mode])[0]
            sm[:, 11:14] = rng.random((n, 3), dtype=np.float32)
            sm[:, 14] = _bits(rng.integers(0, 2 ** 31, n))
            g, r = _both(zl, s, o, p, "BSDF_SAMPLE", sm, 9)
            _assert_same_bits(g, r, (scene, mat, mode, "sample"))
            assert (r[:, 3] > 0).mean() > 0.1


def test_textured_material_lookup(zl):
    # sponza floor: albedo from the sRGB texture array with bilinear filtering -> exact arithmetic
    s, o, p = _setup(zl, "sponza_light", 64, 36)
    rng = np.random.default_rng(9)
    n = 2048
    ev = np.zeros((n, 14), np.float32)
    ev[:, 0] = _bits([0])[0]; ev[:, 1] = _bits([0])[0]
    ev[:, 2:4] = rng.random((n, 2), dtype=np.float32) * 20 - 5
    ev[:, 4:7] = ev[:, 7:10] = ev[:, 10:13] = np.array([0, 0, 1], np.float32)
    g, r = _both(zl, s, o, p, "BSDF_EVAL", ev, 4)
    _assert_same_bits(g, r, "textured albedo")
    assert g[:, :3].std() > 0.01                               # the texture really varies


def test_environment_map_functions(zl):
    s, o, p = _setup(zl, "rungholt_small", 64, 36, envRotation=0.7)
    rng = np.random.default_rng(10)
    n = 8192
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    g, r = _both(zl, s, o, p, "ENV_LE", d, 4)
    _assert_same_bits(g, r, "ENV_LE")                          # atan2 / sin / cos of sphereToPlane, rotateZ
    g, r = _both(zl, s, o, p, "ENV_SAMPLE", rng.random((n, 4), dtype=np.float32), 4)
    _assert_same_bits(g, r, "ENV_SAMPLE")
    assert (r[:, 3] > 0).all()


def test_light_functions(zl):
    s, o, p = _setup(zl, "sponza_light", 64, 36)
    rng = np.random.default_rng(11)
    n = 4096
    nl = s.info["nLightTriangles"]
    lid = rng.integers(0, nl, n)
    u = rng.random((n, 4), dtype=np.float32)
    g, r = _both(zl, s, o, p, "LIGHT_SAMPLE_LE", np.concatenate([_bits(lid).reshape(-1, 1), u], axis=1), 11)
    _assert_same_bits(g, r, "LIGHT_SAMPLE_LE")
    x = (rng.random((n, 3), dtype=np.float32) - 0.5) * np.array([30, 10, 8], np.float32) + np.array([0, 0, 4.5], np.float32)
    y = r[:, :3]
    wo = x - y
    wo /= np.linalg.norm(wo, axis=1, keepdims=True)
    g, r2 = _both(zl, s, o, p, "LIGHT_LE", np.concatenate([_bits(lid).reshape(-1, 1), y, wo, x], axis=1), 4)
    _assert_same_bits(g, r2, "lightLe(light, y, wo)")
    g, r2 = _both(zl, s, o, p, "LIGHT_LE", np.concatenate([_bits(lid).reshape(-1, 1), x, wo, y], axis=1), 4)
    _assert_same_bits(g, r2, "lightPdfLi(light, x, y)")
    assert (r2[:, 3] > 0).mean() > 0.9
    g, r3 = _both(zl, s, o, p, "SAMPLE_LIGHT_ENV", np.concatenate([x, rng.random((n, 5), dtype=np.float32)], axis=1), 7)
    _assert_same_bits(g, r3, "sampleLightAndEnv")
    assert (r3[:, 6] > 0).sum() > 100
