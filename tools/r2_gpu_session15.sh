#!/bin/bash
# round 2, session 15: full GPU suite after the device-side triangle gather + descriptor validation test; bench line (N = 1)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.log
tail -3 gpurun_out/r2_bench_c.log
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_c.json"))
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), d.get("pcie"), d.get("host_binding"))
print(d["scene_prep"]); print(d["strong_scaling"]["seconds"], d["strong_scaling"]["render_and_reduce_seconds"], d["strong_scaling"]["scene_prep"])
print(d["roofline"]["whole_step"])
PY
