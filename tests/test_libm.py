"""include/zl_libm.h — the transcendental functions every implementation of the shading path shares.
CPU: accuracy against binary64 libm (the bounds the header states) and the special values; oracle == reference shim bit for bit.
GPU: the device evaluates the same bits as the host (this is what makes shading bit-exact)."""
import numpy as np
import pytest

from conftest import get_scene

FN = {"sin": 0, "cos": 1, "atan2": 2, "asin": 3, "acos": 4, "log": 5, "pow": 6, "exp": 7}


def _inputs(fn, x, y=None):
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y if y is not None else np.zeros_like(x), np.float32)
    a = np.zeros((x.size, 3), np.float32)
    a[:, 0] = np.array([FN[fn]], np.int32).view(np.float32)[0]
    a[:, 1], a[:, 2] = x, y
    return a


def _eval(evaluator, zl, p, fn, x, y=None):
    return evaluator(p, zl.KAT["LIBM"], _inputs(fn, x, y), 1)[:, 0]


def _ulps(got, ref64):
    ref32 = ref64.astype(np.float32)
    return np.abs(got.astype(np.float64) - ref64) / np.spacing(np.abs(ref32)).astype(np.float64)


def _cases(rng, n):
    return [
        ("sin", rng.uniform(-12867, 12867, n), None, np.sin, 2.5), ("cos", rng.uniform(-12867, 12867, n), None, np.cos, 2.5),
        ("sin", rng.uniform(-7, 7, n), None, np.sin, 2.0), ("cos", rng.uniform(-7, 7, n), None, np.cos, 2.0),
        ("atan2", rng.normal(size=n), rng.normal(size=n), lambda x, y: np.arctan2(y, x), 3.5),
        ("asin", rng.uniform(-1, 1, n), None, np.arcsin, 3.0), ("acos", rng.uniform(-1, 1, n), None, np.arccos, 3.0),
        ("log", np.exp(rng.uniform(-80, 80, n)), None, np.log, 1.0),
        ("pow", np.exp(rng.uniform(-20, 20, n)), rng.uniform(-1.5, 1.5, n), lambda x, y: np.power(x, y), 0.501),
        ("pow", rng.uniform(0, 1, n), rng.uniform(0, 3, n), lambda x, y: np.power(x, y), 0.501),
        ("exp", rng.uniform(-87, 88, n), None, np.exp, 0.501),
    ]


def test_accuracy_and_special_values(zl, oracle):
    s, o = get_scene("cornell", 64, 48)
    p = zl.ZlRenderParams()
    rng = np.random.default_rng(1)
    for fn, x, y, ref, bound in _cases(rng, 200000):
        x32 = x.astype(np.float32)
        y32 = None if y is None else y.astype(np.float32)
        got = _eval(o.debug_eval, zl, p, fn, x32, y32)
        r = ref(x32.astype(np.float64)) if y is None else ref(x32.astype(np.float64), y32.astype(np.float64))
        ok = np.isfinite(r) & (np.abs(r) > 1e-37) & (np.abs(r) < 3e38)
        assert _ulps(got[ok], r[ok]).max() < bound, (fn, _ulps(got[ok], r[ok]).max())
    inf = np.float32(np.inf)
    assert np.array_equal(_eval(o.debug_eval, zl, p, "pow", [0, 0, 2, 4, inf, 0.5, 1, 7], [0, 1, 0.5, -0.5, 1, inf, np.nan, 0]),
                          np.array([1, 0, np.float32(2) ** np.float32(0.5), 0.5, inf, 0, 1, 1], np.float32))
    assert np.isnan(_eval(o.debug_eval, zl, p, "pow", [-1.0], [0.5])).all() and np.isnan(_eval(o.debug_eval, zl, p, "log", [-1.0])).all()
    assert np.array_equal(_eval(o.debug_eval, zl, p, "log", [0, inf, 1]), np.array([-inf, inf, 0], np.float32))
    assert np.array_equal(_eval(o.debug_eval, zl, p, "atan2", [0, -1, 1, inf, -inf], [0, 0, -0.0, inf, inf]),
                          np.array([0, np.pi, -0.0, np.pi / 4, 3 * np.pi / 4], np.float32))
    assert np.isnan(_eval(o.debug_eval, zl, p, "sin", [inf, np.nan])).all() and _eval(o.debug_eval, zl, p, "cos", [1e9])[0] == 1.0
    assert np.isnan(_eval(o.debug_eval, zl, p, "asin", [1.5])).all() and np.isnan(_eval(o.debug_eval, zl, p, "acos", [-1.5])).all()


def test_reference_shim_builtins_are_the_same_functions(zl):
    """sin / cos / atan / asin / acos / log / pow / exp as the reference's GLSL sees them (oracle/ref_shim/glsl_shim.h)"""
    ref_lib = pytest.importorskip("ref_lib")
    if not ref_lib.available():
        pytest.skip("oracle/_ref not present")
    s, o = get_scene("cornell", 64, 48)
    r = ref_lib.RefScene(s.desc)
    p = zl.ZlRenderParams()
    for fn, x, y, _, _ in _cases(np.random.default_rng(2), 50000):
        a = _inputs(fn, x.astype(np.float32), None if y is None else y.astype(np.float32))
        g, q = o.debug_eval(p, zl.KAT["LIBM"], a, 1), r.debug_eval(p, zl.KAT["LIBM"], a, 1)
        assert np.array_equal(g.view(np.uint32), q.view(np.uint32)), fn


@pytest.mark.gpu
def test_device_equals_host_bit_for_bit(zl):
    s, o = get_scene("cornell", 64, 48)
    if not s.device:
        s.upload()
    p = zl.ZlRenderParams()
    p.filmW, p.filmH = 64, 48
    rng = np.random.default_rng(3)
    special = np.array([0, -0.0, 1, -1, np.inf, -np.inf, np.nan, 1e-40, 3e38, 0.5, 2, 1e9, 12867, -12867], np.float32)
    for fn, x, y, _, _ in _cases(rng, 400000):
        x32 = np.concatenate([x.astype(np.float32), special, np.repeat(special, special.size)])
        y32 = np.concatenate([(y if y is not None else np.zeros_like(x)).astype(np.float32), special, np.tile(special, special.size)])
        a = _inputs(fn, x32, y32)
        g, q = zl.debug_eval(s, p, zl.KAT["LIBM"], a, 1), o.debug_eval(p, zl.KAT["LIBM"], a, 1)
        bad = (g.view(np.uint32) != q.view(np.uint32)) & ~(np.isnan(g) & np.isnan(q))
        assert not bad.any(), (fn, a[bad.any(axis=1)][:4], g[bad][:4], q[bad][:4])
