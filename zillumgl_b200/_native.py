"""ctypes bindings of the two native libraries.  There is no Python/CPU fallback: if the
libraries are missing they are built in-tree (nvcc / g++), and if that fails the import
raises."""
import ctypes as C
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))


class ZlCamera(C.Structure):
    _fields_ = [("F", C.c_float * 3), ("R", C.c_float * 3), ("U", C.c_float * 3), ("matInv", C.c_float * 9),
                ("pos", C.c_float * 3), ("tanFOV", C.c_float), ("asp", C.c_float), ("lensRadius", C.c_float),
                ("focalDist", C.c_float)]


class ZlRenderParams(C.Structure):
    _fields_ = [("camera", ZlCamera), ("filmW", C.c_int32), ("filmH", C.c_int32), ("maxDepth", C.c_int32),
                ("russianRoulette", C.c_int32), ("sampleLight", C.c_int32), ("lightEnvUniformSample", C.c_int32),
                ("lightPortion", C.c_float), ("sampler", C.c_int32), ("envRotation", C.c_float), ("spp", C.c_int32),
                ("freeCounter", C.c_int32), ("blocksOnePass", C.c_int32), ("loopsPerPass", C.c_int32),
                ("scale", C.c_float)]

    def copy(self):
        other = ZlRenderParams()
        C.memmove(C.byref(other), C.byref(self), C.sizeof(self))
        return other


class ZlSceneDesc(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "vertices", "normals", "texcoords", "indices", "bounds", "hitTable", "matTexIndices", "materials",
        "lightPower", "lightAlias", "lightProb", "texels", "texUVScale", "envMap", "envAlias", "envAliasProb",
        "noise", "sobolMatrices")] + [(n, C.c_int32) for n in (
        "numVertices", "numTexcoords", "numTriangles", "bvhSize", "objPrimCount", "numMaterials",
        "numLightTriangles", "numTextures", "texMaxW", "texMaxH", "envW", "envH", "noiseW", "noiseH")] + [
        ("lightSum", C.c_float), ("envSum", C.c_float), ("sizeIndices", C.c_void_p)]


class _NoDevice:
    """Stand-in for libzillum_cuda.so in the host-preparation-only mode (ZILLUM_HOST_PREP_ONLY=1): any call raises."""

    def __getattr__(self, name):
        def missing(*_a, **_k):
            raise ZillumError(f"{name}: host-preparation-only mode (ZILLUM_HOST_PREP_ONLY=1), libzillum_cuda.so is not loaded")
        return missing


HOST_PREP_ONLY = os.environ.get("ZILLUM_HOST_PREP_ONLY") == "1"


def _load():
    if HOST_PREP_ONLY:
        prep = os.path.join(_HERE, "host", "libzillum_hostprep.so")
        if not os.path.exists(prep):
            _build.build_hostprep()
        return _NoDevice(), C.CDLL(prep, mode=C.RTLD_GLOBAL)
    cuda_path = os.path.join(_HERE, "csrc", "libzillum_cuda.so")
    host_path = os.path.join(_HERE, "host", "libzillum_host.so")
    if not (os.path.exists(cuda_path) and os.path.exists(host_path)):
        _build.build_cuda()
        _build.build_host()
    cuda = C.CDLL(cuda_path, mode=C.RTLD_GLOBAL)
    host = C.CDLL(host_path, mode=C.RTLD_GLOBAL)
    return cuda, host


cuda, host = _load()

P = C.c_void_p
_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int32)


def _sig(lib, name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


# ---- include/zillum_cuda.h ----
_sig(cuda, "zl_abi_version", C.c_int)
_sig(cuda, "zl_last_error_string", C.c_char_p)
_sig(cuda, "zl_device_count", C.c_int, _i)
_sig(cuda, "zl_set_device", C.c_int, C.c_int)
_sig(cuda, "zl_device_synchronize", C.c_int)
_sig(cuda, "zl_scene_create", C.c_int, C.POINTER(ZlSceneDesc), C.POINTER(P))
_sig(cuda, "zl_scene_destroy", C.c_int, P)
_sig(cuda, "zl_scene_update_materials", C.c_int, P, C.c_int, C.c_int, _f)
_sig(cuda, "zl_build_bvh", C.c_int, _f, C.c_int, C.POINTER(C.c_uint32), C.c_int, _f, _i, _i)
_sig(cuda, "zl_scene_prep_times", C.c_int, P, C.POINTER(C.c_double), C.POINTER(C.c_double), _i)
_sig(cuda, "zl_scene_read_nodes", C.c_int, P, C.c_int, C.c_size_t, C.c_size_t, _f, _i)
_sig(cuda, "zl_scene_memory", C.c_int, P, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t))
_sig(cuda, "zl_film_create", C.c_int, C.c_int, C.c_int, C.POINTER(P))
_sig(cuda, "zl_film_create_external", C.c_int, C.c_int, C.c_int, P, C.POINTER(P))
_sig(cuda, "zl_film_destroy", C.c_int, P)
_sig(cuda, "zl_film_clear", C.c_int, P, P)
_sig(cuda, "zl_film_device_ptr", P, P)
_sig(cuda, "zl_film_download", C.c_int, P, C.c_float, _f, P)
_sig(cuda, "zl_film_download_async", C.c_int, P, C.c_float, _f, P)
_sig(cuda, "zl_film_download_rgb_async", C.c_int, P, C.c_float, _f, P)
_sig(cuda, "zl_film_download_wait", C.c_int, P)
_sig(cuda, "zl_film_flush", C.c_int, P, P)
_sig(cuda, "zl_film_postprocess", C.c_int, P, C.c_float, C.c_int, _f, C.POINTER(C.c_ubyte), P)
_sig(cuda, "zl_film_allreduce", C.c_int, P, P, P)
_sig(cuda, "zl_film_snapshot_async", C.c_int, P, P, P)
_sig(cuda, "zl_launch_path_pass", C.c_int, P, P, C.POINTER(ZlRenderParams), C.c_int, P)
_sig(cuda, "zl_launch_light_pass", C.c_int, P, P, C.POINTER(ZlRenderParams), C.c_int, P)
_sig(cuda, "zl_launch_triple_pt_pass", C.c_int, P, P, C.POINTER(ZlRenderParams), C.c_int, P)
_sig(cuda, "zl_launch_triple_lpt_pass", C.c_int, P, P, C.POINTER(ZlRenderParams), C.c_int, P)
_sig(cuda, "zl_scene_cuda_init_ms", C.c_double, P)
_sig(cuda, "zl_counted_pass", C.c_int, P, P, C.POINTER(ZlRenderParams), C.c_int, C.POINTER(C.c_ulonglong))
_sig(cuda, "zl_counted_pass_untraced", C.c_int, C.POINTER(C.c_ulonglong))
_sig(cuda, "zl_trace_rays", C.c_int, P, _f, C.c_size_t, C.c_int, _f, _i, _f, _i)
_sig(cuda, "zl_rayset_create", C.c_int, _f, C.c_size_t, C.POINTER(P))
_sig(cuda, "zl_rayset_destroy", C.c_int, P)
_sig(cuda, "zl_rayset_trace", C.c_int, P, P, C.c_int, C.c_int, P)
_sig(cuda, "zl_rayset_download", C.c_int, P, _i, _f)
_sig(cuda, "zl_rayset_create_primary", C.c_int, C.POINTER(ZlRenderParams), C.POINTER(P))
_sig(cuda, "zl_rayset_set_tmax", C.c_int, P, _f)
_sig(cuda, "zl_rayset_size", C.c_size_t, P)
_sig(cuda, "zl_rayset_download_rays", C.c_int, P, _f)
_sig(cuda, "zl_debug_eval", C.c_int, P, C.POINTER(ZlRenderParams), C.c_int, _f, C.c_int, _f, C.c_int, C.c_size_t)
_sig(cuda, "zl_rayset_unique_sectors", C.c_int, P, P, C.POINTER(C.c_ulonglong))
_sig(cuda, "zl_launch_count", C.c_ulonglong)
_sig(cuda, "zl_stage_timing_enable", C.c_int, C.c_int)
_sig(cuda, "zl_stage_timing_read", C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong))
_sig(cuda, "zl_measure_read_bandwidth", C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_double))

# ---- include/zillum_host.h ----
_sig(host, "zh_scene_create", P)
_sig(host, "zh_scene_destroy", None, P)
_sig(host, "zh_scene_load", C.c_int, P, C.c_char_p)
_sig(host, "zh_scene_load_builtin", C.c_int, P, C.c_char_p, C.c_int, C.c_int)
_sig(host, "zh_scene_load_xml_text", C.c_int, P, C.c_char_p)
_sig(host, "zh_scene_flatten", C.c_int, P)
_sig(host, "zh_scene_upload", C.c_int, P)
_sig(host, "zh_scene_desc", C.POINTER(ZlSceneDesc), P)
_sig(host, "zh_scene_device", P, P)
_sig(host, "zh_scene_info", None, P, _i)
_sig(host, "zh_scene_times", None, P, C.POINTER(C.c_double))
_sig(host, "zh_scene_light_meshes", None, P, _i, _i, _f)
_sig(host, "zh_scene_set_camera", None, P, _f, _f, C.c_float, C.c_float, C.c_float)
_sig(host, "zh_scene_camera", None, P, C.POINTER(ZlCamera))
_sig(host, "zh_scene_set_sampler", None, P, C.c_int)
_sig(host, "zh_scene_set_device_mtbvh", None, P, C.c_int)
_sig(host, "zh_scene_set_device_bvh", None, P, C.c_int)
_sig(host, "zh_scene_set_env_rotation", None, P, C.c_float)
_sig(host, "zh_builtin_scene_xml", C.c_char_p, C.c_char_p, C.c_int, C.c_int)
_sig(host, "zh_scene_num_models", C.c_int, P)
_sig(host, "zh_scene_model_info", None, P, C.c_int, _i, _f, _f, C.c_char_p, C.c_int)
_sig(host, "zh_scene_model_mesh_counts", None, P, C.c_int, C.c_int, _i)
_sig(host, "zh_scene_model_mesh_data", None, P, C.c_int, C.c_int, _f, _f, _f, C.POINTER(C.c_uint32))
_sig(host, "zh_scene_model_materials", None, P, C.c_int, _f)
_sig(host, "zh_num_images", C.c_int)
_sig(host, "zh_image", None, C.c_int, _i, _i, C.POINTER(C.c_ubyte))
_sig(host, "zh_integrator_create", P, C.c_char_p, P, C.c_int, C.c_int, P, P)
_sig(host, "zh_integrator_destroy", None, P)
_sig(host, "zh_integrator_snapshot_async", C.c_int, P, P)
_sig(host, "zh_integrator_set", C.c_int, P, C.c_char_p, C.c_double)
_sig(host, "zh_integrator_get", C.c_double, P, C.c_char_p)
_sig(host, "zh_integrator_set_sample_shard", None, P, C.c_int, C.c_int)
_sig(host, "zh_integrator_render_one_pass", C.c_int, P)
_sig(host, "zh_integrator_reset", None, P)
_sig(host, "zh_integrator_params", None, P, C.c_int, C.POINTER(ZlRenderParams))
_sig(host, "zh_integrator_film", P, P)
_sig(host, "zh_integrator_result_scale", C.c_float, P)
_sig(host, "zh_integrator_true_scale", C.c_float, P)
_sig(host, "zh_integrator_cur_sample", C.c_int, P)
_sig(host, "zh_integrator_get_frame", C.c_int, P, C.c_float, _f)
_sig(host, "zh_integrator_get_frame_async", C.c_int, P, C.c_float, _f)
_sig(host, "zh_integrator_flush", C.c_int, P)
_sig(host, "zh_integrator_get_frame_rgb_async", C.c_int, P, C.c_float, _f)
_sig(host, "zh_integrator_wait_frame", C.c_int, P)
_sig(host, "zh_build_bvh", C.c_int, _f, C.c_int, C.POINTER(C.c_uint32), C.c_int, _f, _i, C.POINTER(C.c_double))
_sig(host, "zh_alias_table", None, _f, C.c_int, _i, _f)
_sig(host, "zh_env_tables", C.c_float, _f, C.c_int, C.c_int, _i, _f)
_sig(host, "zh_sobol_sample", C.c_uint32, C.c_uint32, C.c_int)
_sig(host, "zh_noise_texture", None, C.c_int, C.c_int, _f)
_sig(host, "zh_write_pfm", C.c_int, C.c_char_p, _f, C.c_int, C.c_int)
_sig(host, "zh_write_exr", C.c_int, C.c_char_p, _f, C.c_int, C.c_int)
_sig(host, "zh_write_png", C.c_int, C.c_char_p, C.POINTER(C.c_ubyte), C.c_int, C.c_int)
_sig(host, "zh_load_byte_image", C.c_int, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_ubyte))
_sig(host, "zh_integrator_post_process", C.c_int, P, C.c_float, C.c_int, _f, C.POINTER(C.c_ubyte))


class ZillumError(RuntimeError):
    pass


def check(rc, what="zillum"):
    if rc != 0:
        raise ZillumError(f"{what} failed ({rc}): {cuda.zl_last_error_string().decode(errors='replace')}")
