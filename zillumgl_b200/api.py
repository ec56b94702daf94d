"""Python mirror of the host interface, for the test / benchmark harness.

The product is the native code: libzillum_cuda.so (kernels behind include/zillum_cuda.h)
and libzillum_host.so (the C++ Scene / Integrator classes that mirror the reference's
src/core/Scene.h and src/core/Integrator.h).  These classes only forward to them, with
the reference's names: Scene.load / createGLContext, NaivePathIntegrator.renderOnePass,
LightPathIntegrator, TriplePathIntegrator, mParam fields.
"""
import ctypes as C

import numpy as np

from . import _native as N
from ._native import ZlCamera, ZlRenderParams, ZlSceneDesc, ZillumError, check  # noqa: F401

_FP = C.POINTER(C.c_float)
_IP = C.POINTER(C.c_int32)


def _fptr(a):
    return a.ctypes.data_as(_FP)


def _iptr(a):
    return a.ctypes.data_as(_IP)


def device_count():
    n = C.c_int32(0)
    rc = N.cuda.zl_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def set_device(i):
    check(N.cuda.zl_set_device(i), "zl_set_device")


def synchronize():
    check(N.cuda.zl_device_synchronize(), "zl_device_synchronize")


def launch_count():
    return int(N.cuda.zl_launch_count())


class Scene:
    """src/core/Scene.h: load(path) + createGLContext(); here createGLContext = flatten + upload."""

    INFO = ("numVertices", "numTriangles", "bvhSize", "objPrimCount", "nLightTriangles", "numMaterials",
            "filmWidth", "filmHeight", "sampler", "numTextures", "envW", "envH", "numLightMeshes")

    def __init__(self):
        self._h = N.host.zh_scene_create()
        self._flattened = False

    def __del__(self, _destroy=N.host.zh_scene_destroy):
        if getattr(self, "_h", None):
            _destroy(self._h)
            self._h = None

    @classmethod
    def builtin(cls, name, width, height):
        s = cls()
        if N.host.zh_scene_load_builtin(s._h, name.encode(), width, height) != 0:
            raise ZillumError(f"unknown builtin scene {name!r}")
        return s

    @classmethod
    def from_file(cls, path):
        s = cls()
        if N.host.zh_scene_load(s._h, str(path).encode()) != 0:
            raise ZillumError(f"cannot load scene {path!r}")
        return s

    @classmethod
    def from_xml(cls, text):
        s = cls()
        if N.host.zh_scene_load_xml_text(s._h, text.encode()) != 0:
            raise ZillumError("cannot parse scene xml")
        return s

    def load(self, path):
        if N.host.zh_scene_load(self._h, str(path).encode()) != 0:
            raise ZillumError(f"cannot load scene {path!r}")
        self._flattened = False
        return True

    def flatten(self):
        empty = N.host.zh_scene_flatten(self._h) != 0
        self._flattened = True
        if empty:
            raise ZillumError("the scene has no triangles (the reference's BVH::build cannot handle that either, BVH.cpp:120)")
        return self

    def upload(self):
        if not self._flattened:
            self.flatten()
        check(N.host.zh_scene_upload(self._h), "Scene.upload")
        return self

    def createGLContext(self, resetTextures=True):
        self.flatten()
        return self.upload()

    @property
    def info(self):
        a = np.zeros(16, np.int32)
        N.host.zh_scene_info(self._h, _iptr(a))
        return dict(zip(self.INFO, (int(x) for x in a)))

    @property
    def times(self):
        t = (C.c_double * 3)()
        N.host.zh_scene_times(self._h, t)
        return {"bvh_build_s": t[0], "mtbvh_flatten_s": t[1], "flatten_s": t[2]}

    @property
    def desc(self):
        """Pointer to the ZlSceneDesc of the flattened scene (host arrays owned by the Scene)."""
        if not self._flattened:
            self.flatten()
        return N.host.zh_scene_desc(self._h)

    @property
    def device(self):
        return N.host.zh_scene_device(self._h)

    def array(self, name):
        """Copy of one flattened host array as numpy (for tests)."""
        d = self.desc.contents
        spec = {
            "vertices": (np.float32, 3 * d.numVertices), "normals": (np.float32, 3 * d.numVertices),
            "texcoords": (np.float32, 2 * d.numTexcoords), "indices": (np.uint32, 3 * d.numTriangles),
            "bounds": (np.float32, 6 * d.bvhSize), "hitTable": (np.int32, 18 * d.bvhSize), "sizeIndices": (np.int32, d.bvhSize),
            "matTexIndices": (np.int32, d.objPrimCount), "materials": (np.float32, 16 * d.numMaterials),
            "lightPower": (np.float32, 3 * d.numLightTriangles), "lightAlias": (np.int32, d.numLightTriangles),
            "lightProb": (np.float32, d.numLightTriangles), "envMap": (np.float32, 3 * d.envW * d.envH),
            "envAlias": (np.int32, (d.envW + 1) * d.envH), "envAliasProb": (np.float32, (d.envW + 1) * d.envH),
            "noise": (np.float32, 2 * d.noiseW * d.noiseH), "sobolMatrices": (np.uint32, 256 * 32),
            "texels": (np.uint8, 3 * d.numTextures * d.texMaxW * d.texMaxH), "texUVScale": (np.float32, 2 * d.numTextures),
        }[name]
        ptr = getattr(d, name)
        if not ptr or spec[1] == 0:
            return np.zeros(0, spec[0])
        buf = (C.c_char * (spec[1] * np.dtype(spec[0]).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=spec[0]).copy()

    def builtin_xml(self, name, width, height):
        return N.host.zh_builtin_scene_xml(name.encode(), width, height).decode()

    def models(self):
        """The scene BEFORE flattening (objects, then lights): per model instance its path, TRS, light power, materials
        and meshes in model space.  Test accessor: tests/test_ref_parity.py feeds these to the reference's own Scene."""
        out = []
        for m in range(N.host.zh_scene_num_models(self._h)):
            info, trs, power = np.zeros(3, np.int32), np.zeros(9, np.float32), np.zeros(3, np.float32)
            path = C.create_string_buffer(512)
            N.host.zh_scene_model_info(self._h, m, _iptr(info), _fptr(trs), _fptr(power), path, 512)
            mats = np.zeros((int(info[2]), 16), np.float32)
            if info[2]:
                N.host.zh_scene_model_materials(self._h, m, _fptr(mats))
            meshes = []
            for k in range(int(info[1])):
                cnt = np.zeros(4, np.int32)
                N.host.zh_scene_model_mesh_counts(self._h, m, k, _iptr(cnt))
                pos, nrm, tex = np.zeros((cnt[0], 3), np.float32), np.zeros((cnt[0], 3), np.float32), np.zeros((cnt[0], 2), np.float32)
                idx = np.zeros(cnt[1], np.uint32)
                N.host.zh_scene_model_mesh_data(self._h, m, k, _fptr(pos), _fptr(nrm), _fptr(tex), idx.ctypes.data_as(C.POINTER(C.c_uint32)))
                meshes.append(dict(pos=pos, nrm=nrm, tex=tex, idx=idx, texIndex=int(cnt[2]), matIndex=int(cnt[3])))
            out.append(dict(path=path.value.decode(), isLight=bool(info[0]), trs=trs, power=power, materials=mats, meshes=meshes))
        return out

    @staticmethod
    def images():
        """Resource::getAllImages(): the pooled 8-bit RGB albedo images, in pool order."""
        res = []
        for i in range(N.host.zh_num_images()):
            w, h = C.c_int(), C.c_int()
            N.host.zh_image(i, C.byref(w), C.byref(h), None)
            px = np.zeros((h.value, w.value, 3), np.uint8)
            N.host.zh_image(i, C.byref(w), C.byref(h), px.ctypes.data_as(C.POINTER(C.c_ubyte)))
            res.append(px)
        return res

    def light_meshes(self):
        n = self.info["numLightMeshes"]
        first, num, power = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(3 * n, np.float32)
        if n:
            N.host.zh_scene_light_meshes(self._h, _iptr(first), _iptr(num), _fptr(power))
        return first, num, power.reshape(n, 3)

    def set_camera(self, pos, angle, fov=45.0, lens_radius=0.0, focal_dist=1.0):
        p, a = np.asarray(pos, np.float32), np.asarray(angle, np.float32)
        N.host.zh_scene_set_camera(self._h, _fptr(p), _fptr(a), fov, lens_radius, focal_dist)

    def camera(self):
        c = ZlCamera()
        N.host.zh_scene_camera(self._h, C.byref(c))
        return c

    def set_device_mtbvh(self, on=True):
        """Skip the host MTBVH flatten; the device threads the six orderings at upload (call before flatten())."""
        N.host.zh_scene_set_device_mtbvh(self._h, 1 if on else 0)
        self._flattened = False
        return self

    def device_prep_times(self):
        """Device-side preparation done at upload: {"bvh_build_ms", "mtbvh_thread_ms", "bvh_levels"} (0 where the host did it)."""
        a, b, lv = C.c_double(0), C.c_double(0), np.zeros(1, np.int32)
        check(N.cuda.zl_scene_prep_times(self.device, C.byref(a), C.byref(b), _iptr(lv)), "device_prep_times")
        return {"bvh_build_ms": a.value, "mtbvh_thread_ms": b.value, "bvh_levels": int(lv[0]),
                "cuda_init_ms": float(N.cuda.zl_scene_cuda_init_ms(self.device))}

    def set_device_bvh(self, on=True):
        """No host BVH at all: zl_scene_create builds the reference's tree on the device and threads it (call before flatten())."""
        N.host.zh_scene_set_device_bvh(self._h, 1 if on else 0)
        self._flattened = False
        return self

    def read_nodes(self, face, first=0, count=None):
        """Threaded node records of one MTBVH face as uploaded: (bounds (count, 6) float32, links (count, 2) int32 = prim|-1, miss)."""
        n = self.info["bvhSize"]
        count = n - first if count is None else count
        b, l = np.empty((count, 6), np.float32), np.empty((count, 2), np.int32)
        check(N.cuda.zl_scene_read_nodes(self.device, face, first, count, _fptr(b), _iptr(l)), "read_nodes")
        return b, l

    def update_materials(self, first, materials):
        """glContext.material->write(...) of the material editor (src/gui/Editor.cpp:73): overwrite `count` 64-byte records of the
        UPLOADED scene starting at record `first` (zl_scene_update_materials); materials: (count, 16) float32."""
        m = np.ascontiguousarray(materials, np.float32).reshape(-1, 16)
        if not self.device:
            raise ZillumError("update_materials: the scene is not uploaded")
        check(N.cuda.zl_scene_update_materials(self.device, int(first), int(m.shape[0]), _fptr(m)), "zl_scene_update_materials")

    def set_sampler(self, sampler):
        N.host.zh_scene_set_sampler(self._h, int(sampler))

    def set_env_rotation(self, radians):
        N.host.zh_scene_set_env_rotation(self._h, float(radians))

    def memory(self):
        total, nodes = C.c_size_t(0), C.c_size_t(0)
        check(N.cuda.zl_scene_memory(self.device, C.byref(total), C.byref(nodes)), "zl_scene_memory")
        return total.value, nodes.value


class _ParamProxy:
    def __init__(self, integ):
        object.__setattr__(self, "_i", integ)

    def __getattr__(self, name):
        v = N.host.zh_integrator_get(self._i._h, name.encode())
        if v < -1e299:
            raise AttributeError(name)
        return v

    def __setattr__(self, name, value):
        if N.host.zh_integrator_set(self._i._h, name.encode(), float(value)) != 0:
            raise AttributeError(name)


class Integrator:
    """src/core/Integrator.h:25-52."""
    TYPE = None

    def __init__(self, scene, width, height, external_film_ptr=None, stream=None, host_only=False):
        # host_only: drive the C++ bookkeeping (pass indices, uniforms, sample shards) without a
        # device; renderOnePass() then launches nothing (Integrator::setDryRun).
        # For CPU tests of the host logic only: nothing is rendered on this path.
        if not scene.device and not host_only:
            scene.createGLContext()
        self.scene, self.width, self.height = scene, width, height
        self._h = N.host.zh_integrator_create(self.TYPE.encode(), scene._h, width, height, external_film_ptr, stream)
        if not self._h:
            raise ZillumError("integrator creation failed")
        self.mParam = _ParamProxy(self)
        if host_only:
            N.host.zh_integrator_set(self._h, b"dryRun", 1.0)

    def __del__(self, _destroy=N.host.zh_integrator_destroy):
        if getattr(self, "_h", None):
            _destroy(self._h)
            self._h = None

    def renderOnePass(self):
        check(N.host.zh_integrator_render_one_pass(self._h), "Integrator.renderOnePass")

    def reset(self):
        N.host.zh_integrator_reset(self._h)

    def snapshotAsync(self, dst_device_ptr):
        """consistent device-side copy of the film into caller memory (Integrator::snapshotAsync)"""
        check(N.host.zh_integrator_snapshot_async(self._h, dst_device_ptr), "Integrator.snapshotAsync")

    def setSampleShard(self, first, stride):
        N.host.zh_integrator_set_sample_shard(self._h, first, stride)

    def params(self, kernel=0):
        p = ZlRenderParams()
        N.host.zh_integrator_params(self._h, kernel, C.byref(p))
        return p

    def resultScale(self):
        return N.host.zh_integrator_result_scale(self._h)

    def trueScale(self):
        return N.host.zh_integrator_true_scale(self._h)

    @property
    def curSample(self):
        return N.host.zh_integrator_cur_sample(self._h)

    @property
    def film(self):
        return N.host.zh_integrator_film(self._h)

    def getFrame(self, scale=None):
        """(H, W, 4) float32, row 0 = bottom; scale=None -> sum / true sample count."""
        out = np.empty((self.height, self.width, 4), np.float32)
        check(N.host.zh_integrator_get_frame(self._h, -1.0 if scale is None else float(scale), _fptr(out)), "getFrame")
        return out


    def getFrameAsync(self, pinned_ptr, scale=None, channels=4):
        """Pipelined read-back into page-locked host memory (address as int); overlaps the next passes.
        channels=3: packed RGB (H*W*3 floats) instead of RGBA with its constant alpha."""
        fn = N.host.zh_integrator_get_frame_rgb_async if channels == 3 else N.host.zh_integrator_get_frame_async
        check(fn(self._h, -1.0 if scale is None else float(scale), C.cast(pinned_ptr, C.POINTER(C.c_float))), "getFrameAsync")

    def flush(self):
        """kernelVariant 2: make the integrator's stream wait for the passes in flight on internal streams."""
        check(N.host.zh_integrator_flush(self._h), "flush")

    def waitFrame(self):
        check(N.host.zh_integrator_wait_frame(self._h), "waitFrame")

    TONE_MAPPERS = {"none": 0, "filmic": 1, "aces": 2}

    def postProcess(self, toneMapper="filmic", scale=None):
        """Display stage (post_proc.glsl): returns (rgba float32 HxWx4, rgb8 uint8 HxWx3), rows in film order."""
        tm = self.TONE_MAPPERS[toneMapper] if isinstance(toneMapper, str) else int(toneMapper)
        rgba = np.empty((self.height, self.width, 4), np.float32)
        rgb8 = np.empty((self.height, self.width, 3), np.uint8)
        check(N.host.zh_integrator_post_process(self._h, -1.0 if scale is None else float(scale), tm, _fptr(rgba),
                                                rgb8.ctypes.data_as(C.POINTER(C.c_ubyte))), "postProcess")
        return rgba, rgb8


class ExternalFilm:
    """A ZlFilm over caller-owned device memory (zl_film_create_external), e.g. the reduced row slice of a multi-GPU frame:
    downloadRgbAsync() resolves (film * scale, packed RGB) and copies it to pinned host memory; wait() completes the oldest."""

    def __init__(self, width, height, device_ptr):
        self._h = C.c_void_p()
        check(N.cuda.zl_film_create_external(width, height, device_ptr, C.byref(self._h)), "zl_film_create_external")

    def __del__(self, _destroy=N.cuda.zl_film_destroy):
        if getattr(self, "_h", None):
            _destroy(self._h)
            self._h = None

    def downloadRgbAsync(self, pinned_ptr, scale, stream=None):
        check(N.cuda.zl_film_download_rgb_async(self._h, scale, C.cast(pinned_ptr, _FP), stream), "zl_film_download_rgb_async")

    def wait(self):
        check(N.cuda.zl_film_download_wait(self._h), "zl_film_download_wait")


class NaivePathIntegrator(Integrator):
    TYPE = "path"


class LightPathIntegrator(Integrator):
    TYPE = "light"


class TriplePathIntegrator(Integrator):
    TYPE = "triple"


def trace_rays(scene, rays, anyhit=False, tmax=None, steps=False):
    """zl_trace_rays: rays (n,6) float32 on the host -> (ids, t[, steps])."""
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    n = rays.shape[0]
    ids, t = np.empty(n, np.int32), np.empty(n, np.float32)
    st = np.empty((n, 2), np.int32) if steps else None
    tm = np.ascontiguousarray(tmax, np.float32) if tmax is not None else None
    check(N.cuda.zl_trace_rays(scene.device, _fptr(rays), n, int(anyhit), _fptr(tm) if tm is not None else None,
                               _iptr(ids), _fptr(t), _iptr(st) if steps else None), "zl_trace_rays")
    return (ids, t, st) if steps else (ids, t)


class RaySet:
    """Device-resident ray buffer (zl_rayset_*), for throughput measurements."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_host(cls, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        h = N.P()
        check(N.cuda.zl_rayset_create(_fptr(rays), rays.shape[0], C.byref(h)), "zl_rayset_create")
        return cls(h)

    @classmethod
    def primary(cls, params):
        h = N.P()
        check(N.cuda.zl_rayset_create_primary(C.byref(params), C.byref(h)), "zl_rayset_create_primary")
        return cls(h)

    def __len__(self):
        return int(N.cuda.zl_rayset_size(self._h))

    def trace(self, scene, anyhit=False, variant=0, stream=None):
        check(N.cuda.zl_rayset_trace(scene.device, self._h, int(anyhit), variant, stream), "zl_rayset_trace")

    def unique_sectors(self, scene):
        """zl_rayset_unique_sectors: dict(lane_nodes, lane_tris, warp_nodes, warp_tris) over the closest-hit walks of this set"""
        o = (C.c_ulonglong * 4)()
        check(N.cuda.zl_rayset_unique_sectors(scene.device, self._h, o), "zl_rayset_unique_sectors")
        return dict(lane_nodes=int(o[0]), lane_tris=int(o[1]), warp_nodes=int(o[2]), warp_tris=int(o[3]))

    def download(self):
        n = len(self)
        ids, t = np.empty(n, np.int32), np.empty(n, np.float32)
        check(N.cuda.zl_rayset_download(self._h, _iptr(ids), _fptr(t)), "zl_rayset_download")
        return ids, t

    def rays(self):
        out = np.empty((len(self), 6), np.float32)
        check(N.cuda.zl_rayset_download_rays(self._h, _fptr(out)), "zl_rayset_download_rays")
        return out

    def __del__(self, _destroy=N.cuda.zl_rayset_destroy):
        if getattr(self, "_h", None):
            _destroy(self._h)
            self._h = None


COUNTER_NAMES = ("rays", "nodes", "tris", "shades", "splats", "paths")


def counted_pass(scene, film, params, kind):
    """zl_counted_pass: one instrumented pass (kind 0 path, 1 light, 2 triple-PT, 3 triple-LPT)
    accumulated into `film`; returns the visit counters as a dict."""
    c = (C.c_ulonglong * 6)()
    check(N.cuda.zl_counted_pass(scene.device, film, C.byref(params), kind, c), "zl_counted_pass")
    out = dict(zip(COUNTER_NAMES, (int(x) for x in c)))
    u = (C.c_ulonglong * 3)()
    check(N.cuda.zl_counted_pass_untraced(u), "zl_counted_pass_untraced")
    # the part of rays / nodes / tris that belongs to shadow rays the production pass does not trace (path tracer only)
    out.update(untraced_rays=int(u[0]), untraced_nodes=int(u[1]), untraced_tris=int(u[2]))
    return out


def algorithmic_bytes(counters, film_rmw_paths=0):
    """SURVEY.md §8(d) byte model: 36 B per hit-table entry visited (12 B link + 24 B AABB),
    48 B per leaf triangle test (12 B indices + 36 B positions), 108 B per shading point,
    32 B film read-modify-write per camera path, 12 B per splat."""
    return (36 * counters["nodes"] + 48 * counters["tris"] + 108 * counters["shades"] + 12 * counters["splats"]
            + 32 * film_rmw_paths)


def debug_eval(scene, params, op, inputs, out_stride):
    inputs = np.ascontiguousarray(inputs, np.float32)
    n, stride = inputs.shape
    out = np.zeros((n, out_stride), np.float32)
    check(N.cuda.zl_debug_eval(scene.device, C.byref(params), op, _fptr(inputs), stride, _fptr(out), out_stride, n), "zl_debug_eval")
    return out


STAGE_NAMES = ("generate", "shade", "sort", "trace", "resolve", "megakernel")   # ZL_STAGE_* of zillum_cuda.h


def stage_timing_enable(on=True):
    check(N.cuda.zl_stage_timing_enable(1 if on else 0), "zl_stage_timing_enable")


def stage_timing_read():
    """Summed device milliseconds and launch counts per stage since the last read: {stage: (ms, launches)}."""
    ms = (C.c_double * len(STAGE_NAMES))()
    n = (C.c_ulonglong * len(STAGE_NAMES))()
    check(N.cuda.zl_stage_timing_read(ms, n), "zl_stage_timing_read")
    return {name: (float(ms[i]), int(n[i])) for i, name in enumerate(STAGE_NAMES)}


def measure_read_bandwidth(nbytes, iters=20):
    v = C.c_double(0)
    check(N.cuda.zl_measure_read_bandwidth(nbytes, iters, C.byref(v)), "zl_measure_read_bandwidth")
    return v.value


def write_pfm(path, rgba):
    rgba = np.ascontiguousarray(rgba, np.float32)
    return N.host.zh_write_pfm(str(path).encode(), _fptr(rgba), rgba.shape[1], rgba.shape[0]) == 0


def write_exr(path, rgba):
    rgba = np.ascontiguousarray(rgba, np.float32)
    return N.host.zh_write_exr(str(path).encode(), _fptr(rgba), rgba.shape[1], rgba.shape[0]) == 0


def build_bvh(vertices, indices):
    """BVH::build on the device (zl_build_bvh): (bounds (2T-1, 6) float32, sizeIndices (2T-1,) int32, levels)."""
    v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
    idx = np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
    n = 2 * idx.shape[0] - 1
    b, s, lv = np.empty((n, 6), np.float32), np.empty(n, np.int32), np.zeros(1, np.int32)
    check(N.cuda.zl_build_bvh(_fptr(v), v.shape[0], idx.ctypes.data_as(C.POINTER(C.c_uint32)), idx.shape[0], _fptr(b), _iptr(s), _iptr(lv)), "build_bvh")
    return b, s, int(lv[0])


def load_byte_image(path):
    """The texture loader (stbi_load(path, ..., 3) of src/core/Image.cpp:10-34): (h, w, 3) uint8, row 0 = top; None if unreadable."""
    w, h = C.c_int(0), C.c_int(0)
    if N.host.zh_load_byte_image(str(path).encode(), C.byref(w), C.byref(h), None) != 0:
        return None
    out = np.empty((h.value, w.value, 3), np.uint8)
    if N.host.zh_load_byte_image(str(path).encode(), C.byref(w), C.byref(h), out.ctypes.data_as(C.POINTER(C.c_ubyte))) != 0:
        return None
    return out


def write_png(path, rgb8):
    """8-bit RGB, rows in film order (row 0 = bottom); flipped on write like the reference's screenshot."""
    rgb8 = np.ascontiguousarray(rgb8, np.uint8)
    return N.host.zh_write_png(str(path).encode(), rgb8.ctypes.data_as(C.POINTER(C.c_ubyte)), rgb8.shape[1], rgb8.shape[0]) == 0


KAT = {name: i for i, name in enumerate((
    "HASH", "SOBOL", "CUBEMAP_FACE", "BOXHIT", "TRIANGLE", "SURFACE", "CAMERA_RAY", "CAMERA_II", "CAMERA_PDF",
    "BSDF_EVAL", "BSDF_SAMPLE", "ENV_LE", "ENV_SAMPLE", "LIGHT_LE", "LIGHT_SAMPLE_LE", "SAMPLE_LIGHT_ENV", "LIBM"))}
