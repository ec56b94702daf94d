"""Times the device BVH build (zl_build_bvh / zl_scene_create with bounds = NULL) per scene, several times in one process."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import zillumgl_b200 as zl
for name, w, h in (("sponza", 64, 36), ("rungholt", 64, 36), ("sponza", 64, 36), ("default", 64, 36), ("cornell", 64, 36)):
    s = zl.Scene.builtin(name, w, h)
    s.flatten()
    v, i = s.array("vertices"), s.array("indices")
    for rep in range(3):
        t0 = time.perf_counter()
        b, si, lv = zl.build_bvh(v, i)
        t1 = time.perf_counter()
        print(f"{name:10s} T={i.size // 3:8d} zl_build_bvh call {rep}: {1e3 * (t1 - t0):8.1f} ms wall, {lv} levels", flush=True)
    s2 = zl.Scene.builtin(name, w, h)
    s2.set_device_bvh(True)
    s2.flatten(); s2.upload()
    print("   scene upload with device build:", s2.device_prep_times(), flush=True)
