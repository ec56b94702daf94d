"""Golden fixtures (tests/golden/*.npz, written by tools/make_golden.py by running THE REFERENCE: its own
shaders and host code compiled for the host, oracle/_ref).

The reference ships no golden vectors, so these committed files are its golden vectors: the CPU tests
below fail if the oracle or the host preparation (BVH build order, alias tables, Sobol, noise seeds, camera)
stops reproducing the reference's outputs, and the GPU tests check the CUDA path against the same files."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, get_scene, rel_mse

sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden as G  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TRAVERSAL = [("cornell", 64, 48), ("default", 64, 36), ("rungholt_small", 64, 36)]


def _load(name):
    return np.load(os.path.join(GOLD, name))


@pytest.mark.parametrize("name,w,h", TRAVERSAL)
def test_oracle_reproduces_traversal_golden(name, w, h):
    g = _load(f"traversal_{name}.npz")
    _, oracle = get_scene(name, w, h)
    ids, t, steps = oracle.trace_rays(g["rays"], steps=True)
    assert np.array_equal(ids, g["ids"]) and np.array_equal(t.view(np.uint32), g["t"].view(np.uint32))
    assert np.array_equal(steps, g["steps"])
    occ, _ = oracle.trace_rays(g["rays"], anyhit=True, tmax=g["tmax"])
    assert np.array_equal(occ, g["occluded"])
    assert 0.05 < (g["ids"] >= 0).mean() < 1.0 and 0 < g["occluded"].mean() < 1


@pytest.mark.parametrize("case", G.film_cases(), ids=lambda c: c[0])
def test_oracle_reproduces_film_golden(case, zl, oracle):
    tag, kind, name, w, h, passes, over = case
    gold = _load(f"film_{tag}.npz")["film"]
    film = G.render_film(zl, oracle, kind, name, w, h, passes, over)
    if kind == "path":                      # one owner per pixel: deterministic, bit for bit
        assert np.array_equal(film[..., :3].view(np.uint32), gold[..., :3].view(np.uint32))
    else:                                   # splats are summed by OpenMP threads in arrival order (the golden: in invocation order)
        assert np.allclose(film[..., :3], gold[..., :3], rtol=2e-5, atol=1e-6 * gold[..., :3].max())
    assert gold[..., :3].max() > 0


def test_kat_golden(zl, oracle):
    g = _load("kat.npz")
    assert all(oracle.lib.zo_hash(int(s)) == int(h) for s, h in zip(g["seeds"], g["hashes"]))
    from zillumgl_b200 import _native as N
    got = np.array([N.host.zh_sobol_sample(int(i), int(d)) for i, d in zip(g["sobol_index"], g["sobol_dim"])], np.uint32)
    assert np.array_equal(got, g["sobol"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h", TRAVERSAL)
def test_cuda_traversal_matches_golden(name, w, h, zl):
    g = _load(f"traversal_{name}.npz")
    s, _ = get_scene(name, w, h)
    if not s.device:
        s.upload()
    ids, t, steps = zl.trace_rays(s, g["rays"], steps=True)
    assert np.array_equal(ids, g["ids"]) and np.array_equal(t.view(np.uint32), g["t"].view(np.uint32))
    assert np.array_equal(steps, g["steps"])                       # zl_trace_rays' counters keep the reference visit sequence
    occ, _ = zl.trace_rays(s, g["rays"], anyhit=True, tmax=g["tmax"])
    assert np.array_equal(occ, g["occluded"])
    rs = zl.RaySet.from_host(g["rays"])                            # production kernel (conservative ignored-slab culling)
    rs.trace(s)
    pid, pt = rs.download()
    assert np.array_equal(pid, g["ids"]) and np.array_equal(pt.view(np.uint32), g["t"].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("case", G.film_cases(), ids=lambda c: c[0])
def test_cuda_film_matches_golden(case, zl):
    """A few passes at tiny resolution against the reference's film: bit for bit for the path tracer, up to the summation
    order of the float-atomic splats for the light and triple tracers."""
    tag, kind, name, w, h, passes, over = case
    gold = _load(f"film_{tag}.npz")["film"]
    s, _ = get_scene(name, w, h)
    if not s.device:
        s.upload()
    cls = {"path": zl.NaivePathIntegrator, "light": zl.LightPathIntegrator, "triple": zl.TriplePathIntegrator}[kind]
    integ = cls(s, w, h)
    if "russianRoulette" in over:
        integ.mParam.russianRoulette = over["russianRoulette"]
    if kind == "light":
        integ.mParam.threadBlocksOnePass = over["blocks"]
    if kind == "triple":
        integ.mParam.LPTBlocksOnePass = over["blocks"]
    for _ in range(passes):
        integ.renderOnePass()
    img = np.ascontiguousarray(integ.getFrame(1.0)[..., :3])
    g3 = np.ascontiguousarray(gold[..., :3])
    if kind == "path":
        assert np.array_equal(img.view(np.uint32), g3.view(np.uint32))
    else:
        assert np.allclose(img, g3, rtol=2e-5, atol=1e-6 * g3.max())
