import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def zl():
    import zillumgl_b200
    return zillumgl_b200


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib


_SCENES = {}


def get_scene(name, w, h):
    """Session cache of flattened builtin scenes (+ their oracle twins)."""
    import oracle_lib
    import zillumgl_b200 as zl
    key = (name, w, h)
    if key not in _SCENES:
        s = zl.Scene.builtin(name, w, h)
        s.flatten()
        _SCENES[key] = (s, oracle_lib.OracleScene(s.desc))
    return _SCENES[key]


@pytest.fixture(scope="session")
def scene_factory():
    return get_scene


def rel_mse(a, b):
    """relMSE of SURVEY.md §8(d): mean over pixels and channels of (a-b)^2 / (b^2 + 1e-2)."""
    a = np.asarray(a, np.float64)[..., :3]
    b = np.asarray(b, np.float64)[..., :3]
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def random_rays(scene, n, seed):
    """Fixed ray set of SURVEY.md §8(d): origins uniform in the (slightly grown) scene box,
    directions uniform on the sphere, plus 5 % axis-parallel and 5 % with one component
    below 1e-6 to reach boxHit's special branches."""
    rng = np.random.default_rng(seed)
    b = scene.array("bounds").reshape(-1, 6)
    lo, hi = b[0, :3], b[0, 3:]
    ext = np.maximum(hi - lo, 1e-3)
    o = (lo - 0.1 * ext) + rng.random((n, 3)) * (1.2 * ext)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = n // 20
    ax = rng.integers(0, 3, k)
    d[:k] = 0.0
    d[np.arange(k), ax] = rng.choice([-1.0, 1.0], k)
    small = rng.integers(0, 3, k)
    d[k:2 * k][np.arange(k), small] = rng.choice([0.0, 3e-7, -5e-7, 9.9e-7], k)
    d[k:2 * k] /= np.linalg.norm(d[k:2 * k], axis=1, keepdims=True)
    return np.concatenate([o, d], axis=1).astype(np.float32)


def random_soup(seed):
    """Randomised builder inputs (vertices (V, 3) float32, indices (3T,) uint32) for the binned-SAH split and its tie rules
    (BVH.cpp:217-296): clustered soups, vertices snapped to a coarse grid (equal centroids, bucket boundaries hit exactly),
    duplicated triangles and slivers along one axis, indexed height fields with shared vertices."""
    rng = np.random.default_rng(1000 + seed)
    T = int(rng.choice([3, 4, 5, 17, 64, 333, 1024, 2500]))
    kind = seed % 4
    if kind == 0:      # clusters of very different size
        c = rng.normal(size=(max(T // 40, 1), 3)) * 20
        base = c[rng.integers(0, c.shape[0], T)] + rng.normal(size=(T, 3)) * rng.choice([0.01, 1.0, 5.0], (T, 1))
        tri = base[:, None, :] + rng.normal(size=(T, 3, 3)) * 0.3
    elif kind == 1:    # snapped to a coarse grid
        tri = np.round(rng.random((T, 3, 3)) * 6) / 2
    elif kind == 2:    # duplicates + slivers along x
        tri = rng.random((T, 3, 3)); tri[:, :, 1:] *= 1e-3
        tri[T // 2:] = tri[:T - T // 2]
    else:              # a height field: indexed mesh with shared vertices
        n = int(np.ceil(np.sqrt(T / 2))) + 1
        gx, gy = np.meshgrid(np.arange(n, dtype=np.float32), np.arange(n, dtype=np.float32))
        v = np.stack([gx, gy, np.floor(rng.random((n, n)) * 3)], -1).reshape(-1, 3).astype(np.float32)
        q = (np.arange(n - 1)[:, None] * n + np.arange(n - 1)[None, :]).reshape(-1)
        idx = np.stack([q, q + 1, q + n, q + 1, q + n + 1, q + n], -1).reshape(-1).astype(np.uint32)[:3 * T]
        tri = None
    if tri is not None:
        v = tri.reshape(-1, 3).astype(np.float32); idx = rng.permutation(T).astype(np.uint32)
        idx = (3 * idx[:, None] + np.arange(3, dtype=np.uint32)[None, :]).reshape(-1).astype(np.uint32)
    return np.ascontiguousarray(v, np.float32), np.ascontiguousarray(idx, np.uint32)
