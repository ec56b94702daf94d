"""ctypes wrapper of oracle/liboracle.so — the CPU restatement of the reference shaders.
Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs, never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ODIR = os.path.join(_ROOT, "oracle")
_FP = C.POINTER(C.c_float)
_IP = C.POINTER(C.c_int32)


def _load():
    so = os.path.join(_ODIR, "liboracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make"], cwd=_ODIR)
    lib = C.CDLL(so)
    P = C.c_void_p
    lib.zo_scene_create.restype = P; lib.zo_scene_create.argtypes = [P]
    lib.zo_scene_destroy.restype = None; lib.zo_scene_destroy.argtypes = [P]
    lib.zo_get_threads.restype = C.c_int
    lib.zo_set_threads.argtypes = [C.c_int]
    for n in ("zo_path_pass", "zo_triple_pt_pass"):
        getattr(lib, n).restype = C.c_int
        getattr(lib, n).argtypes = [P, P, _FP, C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_int]
    for n in ("zo_light_pass", "zo_triple_lpt_pass"):
        getattr(lib, n).restype = C.c_int
        getattr(lib, n).argtypes = [P, P, _FP, C.POINTER(C.c_uint64), C.c_long, C.c_long]
    lib.zo_trace_rays.restype = C.c_int
    lib.zo_trace_rays.argtypes = [P, _FP, C.c_size_t, C.c_int, _FP, _IP, _FP, _IP]
    lib.zo_brute_force_two_nearest.restype = C.c_int
    lib.zo_brute_force_two_nearest.argtypes = [P, _FP, C.c_size_t, _IP, _FP]
    lib.zo_build_bvh.restype = C.c_int
    lib.zo_build_bvh.argtypes = [_FP, C.POINTER(C.c_uint32), C.c_int, _FP, _IP]
    lib.zo_alias_table.restype = None; lib.zo_alias_table.argtypes = [_FP, C.c_int, _IP, _FP]
    lib.zo_env_tables.restype = C.c_float; lib.zo_env_tables.argtypes = [_FP, C.c_int, C.c_int, _IP, _FP]
    lib.zo_sobol_sample.restype = C.c_uint32; lib.zo_sobol_sample.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.c_int]
    lib.zo_hash.restype = C.c_uint32; lib.zo_hash.argtypes = [C.c_uint32]
    lib.zo_camera_update.restype = None
    lib.zo_camera_update.argtypes = [_FP, _FP, C.c_float, C.c_float, C.c_float, C.c_float, P]
    lib.zo_light_table.restype = C.c_float
    lib.zo_light_table.argtypes = [_FP, C.POINTER(C.c_uint32), C.c_int, _IP, _IP, _FP, _FP, _FP]
    lib.zo_round_to_half.restype = C.c_float; lib.zo_round_to_half.argtypes = [C.c_float]
    lib.zo_post_proc.restype = None
    lib.zo_post_proc.argtypes = [_FP, C.c_size_t, C.c_float, C.c_int, _FP, C.POINTER(C.c_ubyte)]
    lib.zo_debug_eval.restype = C.c_int
    lib.zo_debug_eval.argtypes = [P, P, C.c_int, _FP, C.c_int, _FP, C.c_int, C.c_size_t]
    return lib


lib = _load()


def _fp(a):
    return a.ctypes.data_as(_FP)


def _ip(a):
    return a.ctypes.data_as(_IP)


def threads():
    return lib.zo_get_threads()


class OracleScene:
    """Built from the same ZlSceneDesc the product uploads (pointer from Scene.desc)."""

    def __init__(self, desc_ptr):
        self._h = lib.zo_scene_create(C.cast(desc_ptr, C.c_void_p))

    def __del__(self, _destroy=lib.zo_scene_destroy):
        if getattr(self, "_h", None):
            _destroy(self._h)
            self._h = None

    def _pass(self, fn, params, film, *rng):
        stats = (C.c_uint64 * 5)()
        fn(self._h, C.cast(C.byref(params), C.c_void_p), _fp(film), stats, *rng)
        return dict(rays=stats[0], nodeVisits=stats[1], triTests=stats[2], paths=stats[3], splats=stats[4])

    def path_pass(self, params, film, row_begin=0, row_end=-1, row_stride=1):
        return self._pass(lib.zo_path_pass, params, film, row_begin, row_end, row_stride)

    def triple_pt_pass(self, params, film, row_begin=0, row_end=-1, row_stride=1):
        return self._pass(lib.zo_triple_pt_pass, params, film, row_begin, row_end, row_stride)

    def light_pass(self, params, film, id_begin=0, id_end=-1):
        return self._pass(lib.zo_light_pass, params, film, id_begin, id_end)

    def triple_lpt_pass(self, params, film, id_begin=0, id_end=-1):
        return self._pass(lib.zo_triple_lpt_pass, params, film, id_begin, id_end)

    def trace_rays(self, rays, anyhit=False, tmax=None, steps=False, cull_ignored_slab=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        ids, t = np.empty(n, np.int32), np.empty(n, np.float32)
        st = np.empty((n, 2), np.int32) if steps else None
        tm = np.ascontiguousarray(tmax, np.float32) if tmax is not None else None
        lib.zo_trace_rays(self._h, _fp(rays), n, int(anyhit) | (2 if cull_ignored_slab else 0), _fp(tm) if tm is not None else None, _ip(ids), _fp(t),
                          _ip(st) if steps else None)
        return (ids, t, st) if steps else (ids, t)

    def brute_force_two_nearest(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        ids, t = np.empty((n, 2), np.int32), np.empty((n, 2), np.float32)
        lib.zo_brute_force_two_nearest(self._h, _fp(rays), n, _ip(ids), _fp(t))
        return ids, t

    def debug_eval(self, params, op, inputs, out_stride):
        inputs = np.ascontiguousarray(inputs, np.float32)
        n, stride = inputs.shape
        out = np.zeros((n, out_stride), np.float32)
        rc = lib.zo_debug_eval(self._h, C.cast(C.byref(params), C.c_void_p), op, _fp(inputs), stride, _fp(out), out_stride, n)
        assert rc == 0
        return out


def build_bvh(vertices, indices):
    vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1)
    indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
    T = indices.size // 3
    bounds, table = np.empty(6 * (2 * T - 1), np.float32), np.empty(18 * (2 * T - 1), np.int32)
    n = lib.zo_build_bvh(_fp(vertices), indices.ctypes.data_as(C.POINTER(C.c_uint32)), T, _fp(bounds), _ip(table))
    assert n == 2 * T - 1
    return bounds, table


def alias_table(pdf):
    pdf = np.ascontiguousarray(pdf, np.float32)
    alias, prob = np.empty(pdf.size, np.int32), np.empty(pdf.size, np.float32)
    lib.zo_alias_table(_fp(pdf), pdf.size, _ip(alias), _fp(prob))
    return alias, prob


def env_tables(rgb, w, h):
    rgb = np.ascontiguousarray(rgb, np.float32).reshape(-1)
    alias, prob = np.zeros((w + 1) * h, np.int32), np.zeros((w + 1) * h, np.float32)
    s = lib.zo_env_tables(_fp(rgb), w, h, _ip(alias), _fp(prob))
    return alias, prob, s


def sobol_sample(matrices, index, dim):
    matrices = np.ascontiguousarray(matrices, np.uint32)
    return lib.zo_sobol_sample(matrices.ctypes.data_as(C.POINTER(C.c_uint32)), index, dim)


def light_table(vertices, indices, first, count, power):
    vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1)
    indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
    first, count = np.ascontiguousarray(first, np.int32), np.ascontiguousarray(count, np.int32)
    power = np.ascontiguousarray(power, np.float32).reshape(-1)
    n = int(count.sum())
    lp, pdf = np.empty(3 * n, np.float32), np.empty(n, np.float32)
    s = lib.zo_light_table(_fp(vertices), indices.ctypes.data_as(C.POINTER(C.c_uint32)), first.size, _ip(first), _ip(count),
                           _fp(power), _fp(lp), _fp(pdf))
    return lp, pdf, s


def post_proc(film_rgba, scale, tone_mapper):
    """post_proc.glsl on the CPU: (rgba float32, rgb8 uint8) of a HxWx4 film."""
    film = np.ascontiguousarray(film_rgba, np.float32)
    h, w = film.shape[:2]
    out = np.empty((h, w, 4), np.float32)
    out8 = np.empty((h, w, 3), np.uint8)
    lib.zo_post_proc(_fp(film), h * w, float(scale), int(tone_mapper), _fp(out), out8.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return out, out8
