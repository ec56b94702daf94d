"""GPU sweep over tuning switches given as environment settings: times one pass of a workload per
configuration (CUDA events around K passes after warm-up) and compares every film bit for bit with
the megakernel's (tuning switches must not change results).

  python tools/sweep_env.py --workload rungholt --configs "ZL_WF_TRACE_LOOP=0;ZL_WF_TRACE_LOOP=2,ZL_WF_TRACE_MINB=8"
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch
import zillumgl_b200 as zl
import bench as B

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="rungholt")
ap.add_argument("--configs", required=True, help="';'-separated configurations, each a ','-separated list of NAME=VALUE")
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--no-megakernel", action="store_true")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_env.json"))
a = ap.parse_args()
scene, w, h, kind, desc, _ = B.build_scene(zl, a.workload, 0, 0)
res, ref = {}, None
configs = ([] if a.no_megakernel else ["MEGAKERNEL"]) + [c for c in a.configs.split(";") if c.strip() != ""]
touched = set()
for cfg in configs:
    for name in touched:
        os.environ.pop(name, None)
    variant = 1
    if cfg == "MEGAKERNEL":
        variant = 0
    elif cfg != "default":
        for kv in cfg.split(","):
            name, _, val = kv.partition("=")
            os.environ[name.strip()] = val.strip()
            touched.add(name.strip())
    integ = B.make_integrator(zl, scene, kind, w, h, None, variant)
    ppp = B.paths_per_pass(kind, integ, w, h)
    for _ in range(3):
        integ.renderOnePass()
    integ.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(a.steps):
        integ.renderOnePass()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    frame = integ.getFrame(1.0)
    zl.stage_timing_enable(True)
    for _ in range(a.steps):
        integ.renderOnePass()
    stages = {k: round(v[0] / a.steps, 3) for k, v in zl.stage_timing_read().items() if v[1] > 0}
    zl.stage_timing_enable(False)
    if ref is None:
        ref = frame
    same = bool(np.array_equal(ref.view(np.uint32), frame.view(np.uint32)))
    rel = float(np.mean((frame[..., :3].astype(np.float64) - ref[..., :3]) ** 2 / (ref[..., :3].astype(np.float64) ** 2 + 1e-2)))
    res[cfg] = {"stage_ms": stages, "ms_per_pass": ms, "msamples_per_s": ppp / ms / 1e3, "bit_identical_to_first": same, "relmse_vs_first": rel}
    print(f"{cfg:70s} {ms:8.3f} ms/pass {ppp/ms/1e3:8.1f} Msamples/s", "identical" if same else f"DIFFERENT relMSE={rel:.3e}", stages, flush=True)
    del integ
json.dump({"workload": desc, "results": res}, open(a.out, "w"), indent=1)
