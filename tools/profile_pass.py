"""One pass of a bench workload between cudaProfilerStart/Stop, for `ncu --profile-from-start off`:

  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wfTrace \\
      -o gpurun_out/trace_full python tools/profile_pass.py --workload rungholt

Warm-up passes run unprofiled, so the capture holds exactly the launches of one steady-state pass."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import zillumgl_b200 as zl
import bench as B

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="rungholt")
ap.add_argument("--variant", type=int, default=1)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--passes", type=int, default=1)
a = ap.parse_args()
scene, w, h, kind, desc, _ = B.build_scene(zl, a.workload, 0, 0)
integ = B.make_integrator(zl, scene, kind, w, h, None, a.variant)
for _ in range(a.warmup):
    integ.renderOnePass()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.passes):
    integ.renderOnePass()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled", a.passes, "pass(es) of", desc)
