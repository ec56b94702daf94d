"""zillumgl_b200 — B200-native (sm_100a) implementation of ZillumGL's rendering hot path.

Native code: csrc/ (CUDA kernels + C ABI, include/zillum_cuda.h) and host/ (C++ Scene, BVH,
Integrator classes mirroring the reference, include/zillum_host.h).  This package is the
Python harness over them.
"""
from .api import (COUNTER_NAMES, KAT, ExternalFilm, Integrator, LightPathIntegrator, NaivePathIntegrator, RaySet, Scene,  # noqa: F401
                  TriplePathIntegrator, ZillumError, ZlCamera, ZlRenderParams, ZlSceneDesc, algorithmic_bytes, counted_pass, debug_eval,
                  device_count, launch_count, measure_read_bandwidth, set_device, stage_timing_enable, stage_timing_read,
                  STAGE_NAMES, synchronize, trace_rays,
                  write_exr, write_pfm, write_png, load_byte_image, build_bvh)
