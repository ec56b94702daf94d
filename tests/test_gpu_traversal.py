"""GPU parity tests proper (through the C ABI): MTBVH closest-hit / any-hit traversal must
return the oracle's triangle ids, distances and visit counters BIT-EXACTLY."""
import numpy as np
import pytest

from conftest import get_scene, random_rays

pytestmark = pytest.mark.gpu

CASES = [("cornell", 64, 48), ("default", 128, 72), ("rungholt_small", 128, 72)]


def _upload(name, w, h):
    s, o = get_scene(name, w, h)
    if not s.device:
        s.upload()
    return s, o


@pytest.mark.parametrize("name,w,h", CASES)
def test_random_ray_set_ids_bit_exact(name, w, h, zl):
    s, o = _upload(name, w, h)
    rays = random_rays(s, 1 << 16, seed=20261017)
    ids, t, steps = zl.trace_rays(s, rays, steps=True)      # counting build: the reference's exact visit sequence
    rid, rt, rsteps = o.trace_rays(rays, steps=True)
    assert np.array_equal(ids, rid)
    assert np.array_equal(t.view(np.uint32), rt.view(np.uint32))
    assert np.array_equal(steps, rsteps)          # same nodes visited, same triangles tested
    pid, pt = zl.trace_rays(s, rays)                         # production kernel (with the ignored-slab rejection)
    assert np.array_equal(pid, rid) and np.array_equal(pt.view(np.uint32), rt.view(np.uint32))
    assert (ids >= 0).mean() > 0.2


@pytest.mark.parametrize("name,w,h", CASES)
def test_primary_rays_ids_bit_exact(name, w, h, zl):
    s, o = _upload(name, w, h)
    p = zl.ZlRenderParams()
    p.camera = s.camera(); p.camera.asp = w / h
    p.filmW, p.filmH = w, h
    rs = zl.RaySet.primary(p)
    rs.trace(s)
    ids, t = rs.download()
    rays = rs.rays()
    rid, rt = o.trace_rays(rays)
    assert np.array_equal(ids, rid)
    assert np.array_equal(t.view(np.uint32), rt.view(np.uint32))
    # instance id = material/texture word of the hit triangle (the reference flattens instances away)
    mt = s.array("matTexIndices")
    obj = (ids >= 0) & (ids < s.info["objPrimCount"])
    assert np.array_equal(mt[ids[obj]], mt[rid[obj]])
    assert (ids >= 0).mean() > 0.3


@pytest.mark.parametrize("name,w,h", CASES)
def test_shadow_rays_bit_exact(name, w, h, zl):
    s, o = _upload(name, w, h)
    rays = random_rays(s, 1 << 15, seed=7)
    _, t = o.trace_rays(rays)
    rng = np.random.default_rng(1)
    tmax = (np.where(t < 1e7, t, 10.0) * rng.choice([0.5, 0.999, 1.0, 1.001, 2.0], t.size)).astype(np.float32)
    occ, _ = zl.trace_rays(s, rays, anyhit=True, tmax=tmax)
    rocc, _ = o.trace_rays(rays, anyhit=True, tmax=tmax)
    assert np.array_equal(occ, rocc)
    assert 0.05 < occ.mean() < 0.95


@pytest.mark.parametrize("name,w,h", CASES + [("sponza", 64, 36)])
def test_octant_sorted_rays_take_the_specialised_walks_bit_exact(name, w, h, zl):
    """Warps whose 32 rays share one direction octant take traversePure<OCT> (no per-axis min / max, packed
    FADD2 / FMUL2 slab arithmetic; zl_traverse.cuh traverseWarp).  A ray set sorted by octant makes all but
    seven warps uniform, so all eight specialised walks run; ids, distances and occlusion must still be the
    oracle's bits.  (The unsorted sets of the tests above take the general packed walk.)"""
    s, o = _upload(name, w, h)
    rays = random_rays(s, 1 << 16, seed=4242)
    d = rays[:, 3:6]
    octant = (d[:, 0] < 0).astype(np.int64) | ((d[:, 1] < 0).astype(np.int64) << 1) | ((d[:, 2] < 0).astype(np.int64) << 2)
    a = np.abs(d)
    pure = np.all((a >= np.float32(1e-6)) & (a <= np.float32(1.0) - np.float32(1e-6)), axis=1)
    key = np.where(pure, octant, 8)                                       # the 10 % axis-parallel / near-zero rays go last (general walk)
    rays = np.ascontiguousarray(rays[np.argsort(key, kind="stable")])
    assert np.bincount(key, minlength=9)[:8].min() > 1000                 # every octant is populated
    rid, rt = o.trace_rays(rays)
    ids, t = zl.trace_rays(s, rays)
    assert np.array_equal(ids, rid) and np.array_equal(t.view(np.uint32), rt.view(np.uint32))
    tmax = (np.where(rt < 1e7, rt, 10.0) * np.random.default_rng(5).choice([0.5, 0.999, 1.0, 1.001, 2.0], rt.size)).astype(np.float32)
    occ, _ = zl.trace_rays(s, rays, anyhit=True, tmax=tmax)
    assert np.array_equal(occ, o.trace_rays(rays, anyhit=True, tmax=tmax)[0])
    # A/B switch: the same ray set through the general walk only
    import os
    os.environ["ZL_OCTANT_WALK"] = "0"
    try:
        gid, gt = zl.trace_rays(s, rays)
        gocc, _ = zl.trace_rays(s, rays, anyhit=True, tmax=tmax)
    finally:
        os.environ.pop("ZL_OCTANT_WALK", None)
    assert np.array_equal(gid, ids) and np.array_equal(gt.view(np.uint32), t.view(np.uint32)) and np.array_equal(gocc, occ)


def _special_rays(s, n, seed):
    """Every ray has one direction component in boxHit's |d| < 1e-6 branch (incl. exact zero)."""
    rng = np.random.default_rng(seed)
    r = random_rays(s, n, seed)
    k = rng.integers(0, 3, n)
    r[np.arange(n), 3 + k] = rng.choice([0.0, 1e-9, -3e-8, 3e-7, -5e-7, 9.9e-7, -9.99e-7], n)
    nrm = np.linalg.norm(r[:, 3:], axis=1, keepdims=True)
    r[:, 3:] /= np.where(nrm > 0, nrm, 1)
    return r.astype(np.float32)


@pytest.mark.parametrize("name,w,h", CASES + [("sponza", 64, 36)])
def test_near_zero_direction_rays(name, w, h, zl):
    """The reference ignores the slab of an axis with |d| < 1e-6 (App. B #2), which makes such rays
    visit every node overlapping them in the other two axes.  The production kernel adds a
    conservative rejection; ids, distances and occlusion must be unchanged."""
    s, o = _upload(name, w, h)
    rays = _special_rays(s, 1 << 15, seed=99)
    rid, rt, rsteps = o.trace_rays(rays, steps=True)
    ids, t = zl.trace_rays(s, rays)
    assert np.array_equal(ids, rid) and np.array_equal(t.view(np.uint32), rt.view(np.uint32))
    gid, gt, gsteps = zl.trace_rays(s, rays, steps=True)
    assert np.array_equal(gsteps, rsteps)
    tmax = (np.where(rt < 1e7, rt, 10.0) * np.random.default_rng(2).choice([0.5, 0.999, 1.0, 1.001, 2.0], rt.size)).astype(np.float32)
    assert np.array_equal(zl.trace_rays(s, rays, anyhit=True, tmax=tmax)[0], o.trace_rays(rays, anyhit=True, tmax=tmax)[0])


def test_ragged_and_degenerate_inputs(zl):
    s, o = _upload("cornell", 64, 48)
    one = np.array([[0, -3, 1, 0, 1, 0]], np.float32)                        # a single axis-parallel ray
    ids, t = zl.trace_rays(s, one)
    rid, rt = o.trace_rays(one)
    assert ids[0] == rid[0] and t[0] == rt[0]
    odd = random_rays(s, 33, seed=3)                                         # not a multiple of the warp size
    assert np.array_equal(zl.trace_rays(s, odd)[0], o.trace_rays(odd)[0])
    away = np.array([[0, -30, 1, 0, -1, 0], [50, 50, 50, 1, 0, 0]], np.float32)   # rays that leave the scene
    ids, t = zl.trace_rays(s, away)
    assert list(ids) == [-1, -1] and np.all(t == np.float32(1e8))
    zero = np.array([[0, 0, 1, 0, 0, 0]], np.float32)                        # zero direction: never hits, never hangs
    assert zl.trace_rays(s, zero)[0][0] == o.trace_rays(zero)[0][0]


def test_full_size_properties_sponza(zl):
    """BASELINE-size scene (262,144 triangles): size-independent properties instead of a CPU
    comparison of every ray — any-hit is consistent with closest-hit, and the hit distance
    really lies on the reported triangle's plane."""
    s, o = _upload("sponza", 256, 144)
    p = zl.ZlRenderParams()
    p.camera = s.camera(); p.camera.asp = 256 / 144
    p.filmW, p.filmH = 256, 144
    rs = zl.RaySet.primary(p)
    rs.trace(s)
    ids, t = rs.download()
    rays = rs.rays()
    assert (ids >= 0).mean() > 0.9
    sub = np.arange(0, ids.size, 7)
    rid, rt = o.trace_rays(rays[sub])
    assert np.array_equal(ids[sub], rid) and np.array_equal(t[sub], rt)
    hit = ids >= 0
    tmax_lo, tmax_hi = (t * 0.999).astype(np.float32), (t * 1.001).astype(np.float32)
    assert zl.trace_rays(s, rays[hit], anyhit=True, tmax=tmax_lo[hit])[0].sum() <= 0.001 * hit.sum()
    assert zl.trace_rays(s, rays[hit], anyhit=True, tmax=tmax_hi[hit])[0].mean() > 0.999
    v = s.array("vertices").reshape(-1, 3)[s.array("indices").reshape(-1, 3)[ids[hit]]].astype(np.float64)
    pt = rays[hit, :3].astype(np.float64) + rays[hit, 3:].astype(np.float64) * t[hit, None]
    nrm = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True) + 1e-30
    assert np.abs(np.einsum("ij,ij->i", pt - v[:, 0], nrm)).max() < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h", CASES + [("sponza", 32, 18), ("rungholt_small?nx=96&ny=64", 32, 32)])
def test_device_threaded_mtbvh_equals_host_flatten(name, w, h, zl):
    """the six threaded orderings computed on the device from bounds + sizeIndices (threadMtbvhKernel) must be
    the records BVH::buildHitTable (BVH.cpp:298-346) + the host re-pack produce, bit for bit, on every face."""
    host = zl.Scene.builtin(name, w, h)
    host.flatten()
    host.upload()
    dev = zl.Scene.builtin(name, w, h)
    dev.set_device_mtbvh(True)
    dev.flatten()
    assert dev.array("hitTable").size == 0
    dev.upload()
    n = host.info["bvhSize"]
    table = host.array("hitTable").reshape(6, n, 3)
    bounds = host.array("bounds").reshape(n, 6)
    for f in range(6):
        hb, hl = host.read_nodes(f)
        db, dl = dev.read_nodes(f)
        assert np.array_equal(hb.view(np.uint32), db.view(np.uint32)), f"face {f}: bounds differ"
        assert np.array_equal(hl, dl), f"face {f}: links differ"
        # and both are the reference's texels: uBounds[node], (prim, miss) of uHitTable
        assert np.array_equal(dl, table[f, :, 1:3]) and np.array_equal(db.view(np.uint32), bounds[table[f, :, 0]].view(np.uint32))
    # partial reads and argument checks of the accessor
    b, l = dev.read_nodes(5, first=max(n - 3, 0), count=min(3, n))
    assert np.array_equal(l, table[5, max(n - 3, 0):, 1:3])
    with pytest.raises(zl.ZillumError):
        dev.read_nodes(6)
    with pytest.raises(zl.ZillumError):
        dev.read_nodes(0, first=n, count=1)


def _same_floats(a, b):
    """Equal as floats; zeros may differ in sign (min / max of +0 and -0 depends on the host's visiting order)."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and bool(np.all((a == b) | ((a == 0) & (b == 0))))


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h", CASES + [("sponza", 32, 18), ("rungholt_small?nx=96&ny=64", 32, 32)])
def test_device_bvh_build_equals_host_build(name, w, h, zl):
    """BVH::build (16-bucket binned SAH, BVH.cpp:217-296) run level by level on the device must give the host's
    tree: the same pre-order sizeIndices (bit for bit) and the same node bounds."""
    s = zl.Scene.builtin(name, w, h)
    s.set_device_mtbvh(True)
    s.flatten()
    n = s.info["bvhSize"]
    bounds, sizes, levels = zl.build_bvh(s.array("vertices"), s.array("indices"))
    assert np.array_equal(sizes, s.array("sizeIndices"))
    assert _same_floats(bounds, s.array("bounds").reshape(n, 6))
    assert 1 <= levels <= 4 * int(np.ceil(np.log2(max(n, 2)))) + 8
    # a scene uploaded without any host tree renders from the device-built one: node records equal the host-threaded ones
    host = zl.Scene.builtin(name, w, h)
    host.flatten(); host.upload()
    dev = zl.Scene.builtin(name, w, h)
    dev.set_device_bvh(True)
    dev.flatten()
    assert dev.array("bounds").size == 0 and dev.array("sizeIndices").size == 0 and dev.array("hitTable").size == 0
    dev.upload()
    for f in range(6):
        hb, hl = host.read_nodes(f)
        db, dl = dev.read_nodes(f)
        assert np.array_equal(hl, dl), f"face {f}: links differ"
        assert _same_floats(hb, db), f"face {f}: bounds differ"
    rays = random_rays(host, 20000, 11)
    a, b = zl.trace_rays(host, rays), zl.trace_rays(dev, rays)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.gpu
def test_device_bvh_build_degenerate_inputs(zl, oracle):
    """One and two triangles (no split), coincident centroids (every primitive in bucket 0: the `pr--` split), duplicated
    triangles, a long thin strip — against the oracle's restatement of BVH::build."""
    rng = np.random.default_rng(9)
    def check(v, idx):
        v = np.ascontiguousarray(v, np.float32); idx = np.ascontiguousarray(idx, np.uint32)
        T = idx.shape[0]
        b, s, _ = zl.build_bvh(v, idx)
        ob, ot = oracle.build_bvh(v.reshape(-1), idx.reshape(-1))
        ob = ob.reshape(2 * T - 1, 6); ot = ot.reshape(6, 2 * T - 1, 3)
        # the oracle returns the threaded table: face 0 lists (node, prim, miss); rebuild sizeIndices from it
        node, prim, miss = ot[0, :, 0], ot[0, :, 1], ot[0, :, 2]
        size = np.empty(2 * T - 1, np.int32)
        span = miss - np.arange(2 * T - 1)
        size[node] = np.where(prim >= 0, prim | np.int32(-2**31), span)
        assert np.array_equal(s, size)
        assert _same_floats(b, ob)
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    check(tri, [[0, 1, 2]])
    check(np.concatenate([tri, tri + 2]), [[0, 1, 2], [3, 4, 5]])
    check(np.concatenate([tri + 2, tri]), [[0, 1, 2], [3, 4, 5]])                      # the pair is ordered by centroid
    check(np.concatenate([tri] * 7), np.arange(21).reshape(7, 3))                      # identical triangles: centroid box of zero extent
    v = rng.random((300, 3), dtype=np.float32); idx = rng.integers(0, 300, (500, 3))
    check(v, idx)
    strip = np.array([[[i, 0, 0], [i + 1, 0, 0], [i, 1e-3, 0]] for i in range(257)], np.float32).reshape(-1, 3)
    check(strip, np.arange(strip.shape[0]).reshape(-1, 3))
    # symmetric around the centre: many equal centroids along the split axis
    sym = np.concatenate([tri * 0.1 + [x, y, 0] for x in (-1, 0, 1) for y in (-1, 0, 1)]).astype(np.float32)
    check(sym, np.arange(sym.shape[0]).reshape(-1, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_device_bvh_build_random_soups(seed, zl, oracle):
    """The inputs of tests/test_ref_parity.py::test_bvh_random_soups_equal_reference (where the oracle is held to the reference's own
    BVH.cpp) through the device builder: same sizeIndices, same bounds."""
    from conftest import random_soup
    v, idx = random_soup(seed)
    T = idx.size // 3
    b, s, _ = zl.build_bvh(v, idx.reshape(-1, 3))
    ob, ot = oracle.build_bvh(v.reshape(-1), idx)
    ob = ob.reshape(2 * T - 1, 6); ot = ot.reshape(6, 2 * T - 1, 3)
    node, prim, miss = ot[0, :, 0], ot[0, :, 1], ot[0, :, 2]
    size = np.empty(2 * T - 1, np.int32)
    size[node] = np.where(prim >= 0, prim | np.int32(-2**31), miss - np.arange(2 * T - 1))
    assert np.array_equal(s, size)
    assert _same_floats(b, ob)


@pytest.mark.parametrize("name,w,h", [("cornell", 64, 48), ("default", 64, 36), ("rungholt_small", 64, 36), ("sponza_light", 48, 27)])
def test_bvh2_walk_switch_gives_identical_results(name, w, h, zl):
    """ZL_BVH2_WALK=1 (child boxes in the parent + per-lane stack, zl_traverse.cuh traverseBvh2): the reference's visit sequence from one
    copy of the builder's tree.  Ids, distances, any-hit flags and a wavefront film must equal the threaded walk's bit for bit."""
    import os
    s, o = get_scene(name, w, h)
    if not s.device:
        s.upload()
    rays = random_rays(s, 1 << 16, seed=31)
    oi, ot = o.trace_rays(rays)
    tm = np.where(oi >= 0, ot * np.float32(0.999), np.float32(1e8)).astype(np.float32)
    base = zl.NaivePathIntegrator(s, w, h)
    base.mParam.kernelVariant = 1
    for _ in range(3):
        base.renderOnePass()
    film0 = base.getFrame(1.0)
    os.environ["ZL_BVH2_WALK"] = "1"
    try:
        gi, gt = zl.trace_rays(s, rays)
        occ = zl.trace_rays(s, rays, anyhit=True, tmax=tm)[0]
        integ = zl.NaivePathIntegrator(s, w, h)
        integ.mParam.kernelVariant = 1
        for _ in range(3):
            integ.renderOnePass()
        film1 = integ.getFrame(1.0)
    finally:
        os.environ.pop("ZL_BVH2_WALK", None)
    assert np.array_equal(gi, oi) and np.array_equal(gt.view(np.uint32), ot.view(np.uint32))
    assert np.array_equal(occ, o.trace_rays(rays, anyhit=True, tmax=tm)[0])
    assert np.array_equal(film0.view(np.uint32), film1.view(np.uint32)) and film0[..., :3].max() > 0


def test_scene_create_refuses_inconsistent_descriptors(zl):
    """zl_scene_create validates what the kernels would otherwise index with: counts, required arrays, vertex indices (checked by the
    device-side gather), material / texture indices.  Every refusal is an error code + text, nothing is created, and the unmodified
    descriptor still works afterwards."""
    import copy
    import ctypes as C
    from zillumgl_b200 import _native as N
    s = zl.Scene.builtin("cornell", 32, 24)
    s.flatten()
    good = s.desc.contents

    def create(desc):
        h = C.c_void_p()
        rc = N.cuda.zl_scene_create(C.byref(desc), C.byref(h))
        if rc == 0:
            N.cuda.zl_scene_destroy(h)
        return rc, N.cuda.zl_last_error_string().decode(errors="replace")

    def variant(**kw):
        d = N.ZlSceneDesc()
        C.memmove(C.byref(d), C.byref(good), C.sizeof(d))
        for k, v in kw.items():
            setattr(d, k, v)
        return d

    assert create(variant())[0] == 0
    idx = s.array("indices").copy()
    idx[7] = good.numVertices + 5
    rc, msg = create(variant(indices=idx.ctypes.data))
    assert rc != 0 and "vertex index out of range" in msg
    mt = s.array("matTexIndices").copy()
    mt[0] = good.numMaterials
    rc, msg = create(variant(matTexIndices=mt.ctypes.data))
    assert rc != 0 and "material / texture index out of range" in msg
    for kw, text in ((dict(objPrimCount=good.objPrimCount + 1), "counts are inconsistent"), (dict(bvhSize=good.bvhSize + 1), "bvhSize"),
                     (dict(vertices=None), "missing required array"), (dict(lightPower=None), "light tables missing")):
        rc, msg = create(variant(**kw))
        assert rc != 0 and text in msg, (kw, msg)
    assert create(variant())[0] == 0
