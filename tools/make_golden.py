"""Generate the golden fixtures under tests/golden/ by running THE REFERENCE ITSELF (oracle/_ref:
the reference's own GLSL text and C++ host code compiled for the host, see oracle/Makefile `ref`).
Run in the container that holds /root/reference; the fixtures are committed and travel.

The reference ships no tests or golden vectors for this path, so these files ARE its
golden vectors: outputs of its own shaders on fixed seeded inputs, which the oracle (CPU tests) and
the CUDA path (GPU tests) must reproduce:
  traversal_<scene>.npz   fixed ray set -> bvhHit (triangle id, t), bvhTest flags, bvhDebug's counter (reference);
                          (node visits, triangle tests) per ray from the oracle's instrumented walk
  film_<integrator>.npz   accumulated film of a few passes of each integrator's shader (tiny resolution)
  kat.npz                 hash / Sobol known answers (random.glsl:5-13, Sampler.cpp:19-28 over SobolMatrices256x32.h)

  python tools/make_golden.py          # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
GOLD = os.path.join(ROOT, "tests", "golden")


def film_cases():
    # (file tag, integrator kind, scene, w, h, passes, param overrides)
    return [("path_cornell", "path", "cornell", 24, 18, 4, {}),
            ("path_default", "path", "default", 32, 18, 3, {}),
            ("path_rungholt_small_rr", "path", "rungholt_small", 24, 14, 3, dict(russianRoulette=1)),
            ("light_cornell", "light", "cornell", 24, 18, 4, dict(blocks=1)),
            ("triple_cornell", "triple", "cornell", 24, 18, 3, dict(blocks=1))]


def film_params(zl, scene, kind, w, h, i, kernel, over):
    """Uniforms of pass i exactly as the host Integrator classes produce them (defaults of Integrator.h)."""
    p = zl.ZlRenderParams()
    p.camera = scene.camera(); p.camera.asp = w / h
    p.filmW, p.filmH = w, h
    p.maxDepth, p.russianRoulette, p.sampleLight, p.lightEnvUniformSample, p.lightPortion = 4, int(over.get("russianRoulette", 0)), 1, 0, 0.5
    p.sampler = 1 if (kind in ("path", "triple") and kernel == 0) else 0
    p.spp, p.freeCounter = i, i + 1
    p.blocksOnePass, p.loopsPerPass, p.scale = 0, 1, 1.0
    if kind == "light" or kernel == 1:
        p.blocksOnePass = int(over.get("blocks", 1))
        p.scale = w * h / (p.blocksOnePass * 1536.0) if kind == "triple" else 1.0
    return p


def render_film(zl, O, kind, name, w, h, passes, over, reference=False):
    """the film of `passes` passes by the oracle (O = oracle_lib), or by the reference's shaders (reference=True, O = ref_lib)"""
    from conftest import get_scene
    scene, oracle = get_scene(name, w, h)
    if reference:
        oracle = O.RefScene(scene.desc)
    film = np.zeros((h, w, 4), np.float32)
    for i in range(passes):
        if kind == "path":
            oracle.path_pass(film_params(zl, scene, kind, w, h, i, 0, over), film)
        elif kind == "light":
            oracle.light_pass(film_params(zl, scene, kind, w, h, i, 0, over), film)
        else:
            oracle.triple_pt_pass(film_params(zl, scene, kind, w, h, i, 0, over), film)
            oracle.triple_lpt_pass(film_params(zl, scene, kind, w, h, i, 1, over), film)
    return film


def traversal_case(R, name, w, h, n=4096, seed=1234):
    from conftest import get_scene, random_rays
    scene, oracle = get_scene(name, w, h)
    ref = R.RefScene(scene.desc)
    rays = random_rays(scene, n, seed)
    ids, t, entered = ref.trace_rays(rays, steps=True)                       # bvhHit + bvhDebug of the reference
    tmax = np.where(ids >= 0, t * 0.9 + 0.05, 5.0).astype(np.float32)
    occ, _ = ref.trace_rays(rays, anyhit=True, tmax=tmax)                    # bvhTest of the reference
    _, _, steps = oracle.trace_rays(rays, steps=True)                        # (entries fetched, triangles tested): the reference has no such counter
    return dict(rays=rays, ids=ids, t=t, entered=entered, steps=steps, tmax=tmax, occluded=occ)


def kat_case(zl, R):
    from conftest import get_scene
    rng = np.random.default_rng(77)
    seeds = rng.integers(0, 2 ** 32, 256, dtype=np.uint64).astype(np.uint32)
    scene, _ = get_scene("cornell", 64, 48)
    ref = R.RefScene(scene.desc)
    hashes = ref.debug_eval(zl.ZlRenderParams(), zl.KAT["HASH"], seeds.view(np.float32).reshape(-1, 1), 1).view(np.uint32).reshape(-1)
    idx, dim = rng.integers(0, 131072, 256), rng.integers(0, 256, 256)
    sob = np.array([R.sobol_sample(int(i), int(d)) for i, d in zip(idx, dim)], np.uint32)
    return dict(seeds=seeds, hashes=hashes, sobol_index=idx.astype(np.int32), sobol_dim=dim.astype(np.int32), sobol=sob)


def main():
    import ref_lib as R
    import zillumgl_b200 as zl
    os.makedirs(GOLD, exist_ok=True)
    np.save(os.path.join(GOLD, "sobol_matrices_256x32.npy"), R.sobol_matrices().reshape(256, 32))     # SobolMatrices256x32.h as compiled
    R.set_threads(1)                          # splats in invocation order: the light / triple films are reproducible bit for bit
    for name, w, h in (("cornell", 64, 48), ("default", 64, 36), ("rungholt_small", 64, 36)):
        np.savez_compressed(os.path.join(GOLD, f"traversal_{name}.npz"), **traversal_case(R, name, w, h))
    for tag, kind, name, w, h, passes, over in film_cases():
        np.savez_compressed(os.path.join(GOLD, f"film_{tag}.npz"), film=render_film(zl, R, kind, name, w, h, passes, over, reference=True))
    np.savez_compressed(os.path.join(GOLD, "kat.npz"), **kat_case(zl, R))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
