"""GPU sweep: time one wavefront path pass of a workload per trace-kernel tuning variant
(ZL_WF_TRACE_VARIANT), CUDA events around K passes after warm-up.  Films are compared bit for bit
across variants (tuning switches must not change results)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch
import zillumgl_b200 as zl
import bench as B

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="rungholt")
ap.add_argument("--variants", default="0,0s,6,6s", help="ZL_WF_TRACE_SIMPLE masks (bit0: plain loop for camera rays, bit1: for other bounces); suffix s = with ray sorting")
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--rays-per-lane", default="1")
ap.add_argument("--env", default="", help="semicolon-separated list of extra env settings to sweep, e.g. 'ZL_WF_SORT_MODE=1;ZL_WF_SORT_MODE=2'")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_wf.json"))
a = ap.parse_args()
scene, w, h, kind, desc, _ = B.build_scene(zl, a.workload, 0, 0)
res, ref = {}, None
extra = a.env.split(";") if a.env else [""]
for tag in ["-1"] + [f"{v}r{r}|{e}" for e in extra for r in a.rays_per_lane.split(",") for v in a.variants.split(",")]:
    tag, _, env = tag.partition("|")
    for kv in (a.env.split(";") if a.env else []):
        os.environ.pop(kv.split("=")[0], None)
    if env:
        os.environ[env.split("=")[0]] = env.split("=")[1]
        tag_full = tag + " " + env
    else:
        tag_full = tag
    if "r" in tag:
        os.environ["ZL_WF_RAYS_PER_LANE"] = tag.split("r")[1]
    v = int(tag.split("r")[0].rstrip("s"))
    if v >= 0:
        os.environ["ZL_WF_TRACE_SIMPLE"] = str(v)
        os.environ["ZL_WF_SORT"] = "1" if tag.split("r")[0].endswith("s") else "0"
    integ = zl.NaivePathIntegrator(scene, w, h)
    integ.mParam.kernelVariant = 0 if v < 0 else 1
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
    if ref is None:
        ref = frame
    same = bool(np.array_equal(ref.view(np.uint32), frame.view(np.uint32)))
    res["megakernel" if v < 0 else f"wf_{tag_full}"] = {"ms_per_pass": ms, "msamples_per_s": w * h / ms / 1e3, "bit_identical_to_megakernel": same}
    print(tag_full, f"{ms:.3f} ms/pass", f"{w*h/ms/1e3:.1f} Msamples/s", "identical" if same else "DIFFERENT", flush=True)
    del integ
json.dump({"workload": desc, "results": res}, open(a.out, "w"), indent=1)
