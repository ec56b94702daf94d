"""ctypes wrapper of oracle/_ref/libzillum_ref.so — the REFERENCE ITSELF built for the host: its own
C++ host code compiled unmodified against stand-in third-party headers, and its own GLSL text
compiled as C++ (recipe: oracle/Makefile target `ref`, oracle/ref_glsl2cpp.py, oracle/ref_shim/).
Test infrastructure only.  The library is built in the container that holds /root/reference and
travels to the GPU box as a prebuilt file; nothing here reads the reference tree at run time."""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ODIR = os.path.join(_ROOT, "oracle")
SO = os.path.join(_ODIR, "_ref", "libzillum_ref.so")
_FP = C.POINTER(C.c_float)
_IP = C.POINTER(C.c_int32)
_UP = C.POINTER(C.c_uint32)
REFERENCE = os.environ.get("ZILLUM_REFERENCE", "/root/reference")


def available():
    return os.path.exists(SO) or os.path.isdir(os.path.join(REFERENCE, "src", "shader"))


def _load():
    if os.path.isdir(os.path.join(REFERENCE, "src", "shader")):
        subprocess.check_call(["make", "ref", "REF=" + REFERENCE], cwd=_ODIR, stdout=subprocess.DEVNULL)   # up to date: no-op
    if not os.path.exists(SO):
        raise RuntimeError("oracle/_ref/libzillum_ref.so is missing and the reference tree is not here to build it from")
    lib = C.CDLL(SO)
    P = C.c_void_p
    lib.zr_scene_create.restype = P; lib.zr_scene_create.argtypes = [P]
    lib.zr_scene_destroy.argtypes = [P]
    lib.zr_get_threads.restype = C.c_int
    lib.zr_set_threads.argtypes = [C.c_int]
    for n in ("zr_path_pass", "zr_triple_pt_pass"):
        getattr(lib, n).restype = C.c_int
        getattr(lib, n).argtypes = [P, P, _FP, C.c_int, C.c_int, C.c_int]
    for n in ("zr_light_pass", "zr_triple_lpt_pass"):
        getattr(lib, n).restype = C.c_int
        getattr(lib, n).argtypes = [P, P, _FP, C.c_long, C.c_long]
    lib.zr_trace_rays.restype = C.c_int
    lib.zr_trace_rays.argtypes = [P, _FP, C.c_size_t, C.c_int, _FP, _IP, _FP, _IP]
    lib.zr_debug_eval.restype = C.c_int
    lib.zr_debug_eval.argtypes = [P, P, C.c_int, _FP, C.c_int, _FP, C.c_int, C.c_size_t]
    lib.zr_post_proc.restype = C.c_int
    lib.zr_post_proc.argtypes = [_FP, C.c_int, C.c_int, C.c_float, C.c_int, _FP]
    lib.zr_build_bvh.restype = C.c_int
    lib.zr_build_bvh.argtypes = [_FP, C.c_int, _UP, C.c_int, _FP, _IP]
    lib.zr_alias_table.restype = None; lib.zr_alias_table.argtypes = [_FP, C.c_int, _IP, _FP]
    lib.zr_sobol_sample.restype = C.c_uint32; lib.zr_sobol_sample.argtypes = [C.c_uint32, C.c_int]
    lib.zr_sobol_matrices.restype = None; lib.zr_sobol_matrices.argtypes = [_UP]
    lib.zr_env_tables.restype = C.c_float; lib.zr_env_tables.argtypes = [_FP, C.c_int, C.c_int, _IP, _FP, _FP]
    lib.zr_camera_update.restype = None
    lib.zr_camera_update.argtypes = [_FP, _FP, C.c_float, C.c_float, C.c_float, C.c_float, P]
    lib.zr_noise_texture.restype = None; lib.zr_noise_texture.argtypes = [C.c_int, C.c_int, _FP]
    lib.zr_register_image.restype = None
    lib.zr_register_image.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, _FP, C.POINTER(C.c_ubyte)]
    lib.zr_full_reset.restype = None
    lib.zr_full_register_model.restype = None; lib.zr_full_register_model.argtypes = [C.c_char_p]
    lib.zr_full_model_add_mesh.restype = None
    lib.zr_full_model_add_mesh.argtypes = [C.c_char_p, C.c_int, _FP, _FP, _FP, C.c_int, _UP, C.c_char_p, C.c_int]
    lib.zr_full_model_set_materials.restype = None; lib.zr_full_model_set_materials.argtypes = [C.c_char_p, C.c_int, _FP]
    lib.zr_full_scene_load.restype = P; lib.zr_full_scene_load.argtypes = [C.c_char_p, _FP]
    lib.zr_full_scene_destroy.argtypes = [P]
    lib.zr_full_scene_info.restype = None; lib.zr_full_scene_info.argtypes = [P, _IP, _FP]
    lib.zr_full_scene_array.restype = C.c_size_t; lib.zr_full_scene_array.argtypes = [P, C.c_char_p, P, C.c_size_t]
    lib.zr_full_scene_set.restype = None; lib.zr_full_scene_set.argtypes = [P, C.c_char_p, C.c_float]
    lib.zr_full_scene_camera.restype = None; lib.zr_full_scene_camera.argtypes = [P, P]
    lib.zr_full_integrator_create.restype = P; lib.zr_full_integrator_create.argtypes = [P, C.c_char_p, C.c_int, C.c_int]
    lib.zr_full_integrator_destroy.argtypes = [P]
    lib.zr_full_integrator_set.restype = C.c_int; lib.zr_full_integrator_set.argtypes = [P, C.c_char_p, C.c_double]
    lib.zr_full_integrator_render_one_pass.restype = None; lib.zr_full_integrator_render_one_pass.argtypes = [P]
    lib.zr_full_integrator_result_scale.restype = C.c_float; lib.zr_full_integrator_result_scale.argtypes = [P]
    lib.zr_full_integrator_get_frame.restype = C.c_size_t; lib.zr_full_integrator_get_frame.argtypes = [P, _FP, C.c_size_t]
    return lib


lib = _load()


def _fp(a):
    return a.ctypes.data_as(_FP)


def _ip(a):
    return a.ctypes.data_as(_IP)


def threads():
    return lib.zr_get_threads()


def set_threads(n):
    lib.zr_set_threads(int(n))


class RefScene:
    """The reference's shader programs over the arrays of a ZlSceneDesc (same construction as
    oracle_lib.OracleScene, so that either can stand behind the same test)."""

    def __init__(self, desc_ptr):
        self._h = lib.zr_scene_create(C.cast(desc_ptr, C.c_void_p))

    def __del__(self, _destroy=lib.zr_scene_destroy):
        if getattr(self, "_h", None):
            _destroy(self._h)
            self._h = None

    def path_pass(self, params, film, row_begin=0, row_end=-1, row_stride=1):
        lib.zr_path_pass(self._h, C.cast(C.byref(params), C.c_void_p), _fp(film), row_begin, row_end, row_stride)

    def triple_pt_pass(self, params, film, row_begin=0, row_end=-1, row_stride=1):
        lib.zr_triple_pt_pass(self._h, C.cast(C.byref(params), C.c_void_p), _fp(film), row_begin, row_end, row_stride)

    def light_pass(self, params, film, id_begin=0, id_end=-1):
        lib.zr_light_pass(self._h, C.cast(C.byref(params), C.c_void_p), _fp(film), id_begin, id_end)

    def triple_lpt_pass(self, params, film, id_begin=0, id_end=-1):
        lib.zr_triple_lpt_pass(self._h, C.cast(C.byref(params), C.c_void_p), _fp(film), id_begin, id_end)

    def trace_rays(self, rays, anyhit=False, tmax=None, steps=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        ids, t = np.empty(n, np.int32), np.empty(n, np.float32)
        st = np.empty(n, np.int32) if steps else None
        tm = np.ascontiguousarray(tmax, np.float32) if tmax is not None else None
        lib.zr_trace_rays(self._h, _fp(rays), n, int(anyhit), _fp(tm) if tm is not None else None, _ip(ids), _fp(t),
                          _ip(st) if steps else None)
        return (ids, t, st) if steps else (ids, t)

    def debug_eval(self, params, op, inputs, out_stride):
        inputs = np.ascontiguousarray(inputs, np.float32)
        n, stride = inputs.shape
        out = np.zeros((n, out_stride), np.float32)
        rc = lib.zr_debug_eval(self._h, C.cast(C.byref(params), C.c_void_p), op, _fp(inputs), stride, _fp(out), out_stride, n)
        assert rc == 0
        return out


def build_bvh(vertices, indices):
    vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1)
    indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
    T = indices.size // 3
    bounds, table = np.empty(6 * (2 * T - 1), np.float32), np.empty(18 * (2 * T - 1), np.int32)
    n = lib.zr_build_bvh(_fp(vertices), vertices.size // 3, indices.ctypes.data_as(_UP), T, _fp(bounds), _ip(table))
    assert n == 2 * T - 1
    return bounds, table


def alias_table(pdf):
    pdf = np.ascontiguousarray(pdf, np.float32)
    alias, prob = np.empty(pdf.size, np.int32), np.empty(pdf.size, np.float32)
    lib.zr_alias_table(_fp(pdf), pdf.size, _ip(alias), _fp(prob))
    return alias, prob


def sobol_sample(index, dim):
    return lib.zr_sobol_sample(int(index), int(dim))


def sobol_matrices():
    m = np.empty(256 * 32, np.uint32)
    lib.zr_sobol_matrices(m.ctypes.data_as(_UP))
    return m


def env_tables(rgb, w, h):
    rgb = np.ascontiguousarray(rgb, np.float32).reshape(-1)
    alias, prob = np.zeros((w + 1) * h, np.int32), np.zeros((w + 1) * h, np.float32)
    texels = np.zeros(3 * w * h, np.float32)
    s = lib.zr_env_tables(_fp(rgb), w, h, _ip(alias), _fp(prob), _fp(texels))
    return alias, prob, s, texels


def camera_update(zl_camera_cls, pos, angle, fov, aspect, lens, focal):
    out = zl_camera_cls()
    p, a = np.asarray(pos, np.float32), np.asarray(angle, np.float32)
    lib.zr_camera_update(_fp(p), _fp(a), fov, aspect, lens, focal, C.cast(C.byref(out), C.c_void_p))
    return out


def noise_texture(w, h):
    out = np.empty(2 * w * h, np.float32)
    lib.zr_noise_texture(w, h, _fp(out))
    return out


def post_proc(film_rgba, scale, tone_mapper):
    film = np.ascontiguousarray(film_rgba, np.float32)
    h, w = film.shape[:2]
    out = np.zeros((h, w, 4), np.float32)
    lib.zr_post_proc(_fp(film), w, h, float(scale), int(tone_mapper), _fp(out))
    return out


# ---- the reference's Scene / Integrator classes on a whole scene ----
_ARRAYS = {"vertices": np.float32, "normals": np.float32, "texcoords": np.float32, "indices": np.uint32, "bounds": np.float32,
           "hitTable": np.int32, "matTexIndices": np.int32, "materials": np.float32, "lightPower": np.float32, "lightAlias": np.int32,
           "lightProb": np.float32, "texUVScale": np.float32, "envMap": np.float32, "envAlias": np.int32, "envAliasProb": np.float32,
           "noise": np.float32, "texels": np.uint8}


def full_reset():
    lib.zr_full_reset()


def register_image(path, w, h, rgb_float=None, rgb8=None):
    f = np.ascontiguousarray(rgb_float, np.float32) if rgb_float is not None else None
    b = np.ascontiguousarray(rgb8, np.uint8) if rgb8 is not None else None
    lib.zr_register_image(path.encode(), w, h, 3, _fp(f) if f is not None else None,
                          b.ctypes.data_as(C.POINTER(C.c_ubyte)) if b is not None else None)


def register_model(path, meshes, materials):
    """meshes: list of dict(pos, nrm, tex, idx, texture, matIndex); materials: (n, 16) float32"""
    lib.zr_full_register_model(path.encode())
    for m in meshes:
        pos = np.ascontiguousarray(m["pos"], np.float32).reshape(-1)
        nrm = np.ascontiguousarray(m["nrm"], np.float32).reshape(-1)
        tex = np.ascontiguousarray(m["tex"], np.float32).reshape(-1)
        idx = np.ascontiguousarray(m["idx"], np.uint32).reshape(-1)
        lib.zr_full_model_add_mesh(path.encode(), pos.size // 3, _fp(pos), _fp(nrm), _fp(tex), idx.size, idx.ctypes.data_as(_UP),
                                   (m.get("texture") or "").encode(), int(m["matIndex"]))
    mats = np.ascontiguousarray(materials, np.float32).reshape(-1)
    lib.zr_full_model_set_materials(path.encode(), mats.size // 16, _fp(mats) if mats.size else None)


class FullScene:
    def __init__(self, xml_text, noise=None):
        n = np.ascontiguousarray(noise, np.float32) if noise is not None else None
        self._noise = n
        self._h = lib.zr_full_scene_load(xml_text.encode(), _fp(n) if n is not None else None)
        if not self._h:
            raise RuntimeError("reference Scene::load failed")

    def __del__(self, _destroy=lib.zr_full_scene_destroy):
        if getattr(self, "_h", None):
            _destroy(self._h)
            self._h = None

    @property
    def info(self):
        ints, floats = np.zeros(16, np.int32), np.zeros(4, np.float32)
        lib.zr_full_scene_info(self._h, _ip(ints), _fp(floats))
        keys = ("numVertices", "numTriangles", "bvhSize", "objPrimCount", "nLightTriangles", "numMaterials", "filmWidth", "filmHeight", "sampler",
                "numTextures", "envW", "envH")
        d = dict(zip(keys, (int(v) for v in ints)))
        d.update(lightSum=float(floats[0]), envSum=float(floats[1]), envRotation=float(floats[2]))
        return d

    def array(self, name):
        n = lib.zr_full_scene_array(self._h, name.encode(), None, 0)
        out = np.empty(n, np.uint8)
        if n:
            lib.zr_full_scene_array(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), n)
        return out.view(_ARRAYS[name])

    def set(self, name, value):
        lib.zr_full_scene_set(self._h, name.encode(), float(value))

    def camera(self, zl_camera_cls):
        out = zl_camera_cls()
        lib.zr_full_scene_camera(self._h, C.cast(C.byref(out), C.c_void_p))
        return out


class FullIntegrator:
    """NaivePathIntegrator / LightPathIntegrator / TriplePathIntegrator of the reference, driven like Application.cpp."""

    def __init__(self, scene, kind, w, h):
        self.scene, self.w, self.h = scene, w, h
        self._h = lib.zr_full_integrator_create(scene._h, kind.encode(), w, h)
        assert self._h

    def __del__(self, _destroy=lib.zr_full_integrator_destroy):
        if getattr(self, "_h", None):
            _destroy(self._h)
            self._h = None

    def set(self, name, value):
        assert lib.zr_full_integrator_set(self._h, name.encode(), float(value)) == 0, name

    def renderOnePass(self):
        lib.zr_full_integrator_render_one_pass(self._h)

    def resultScale(self):
        return lib.zr_full_integrator_result_scale(self._h)

    def getFrame(self):
        out = np.zeros((self.h, self.w, 4), np.float32)
        lib.zr_full_integrator_get_frame(self._h, _fp(out), out.nbytes)
        return out
