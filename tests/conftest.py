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
