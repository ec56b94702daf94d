#!/bin/bash
# round 2, session 41: dense (56-register) stage kernels in the pipelined path tracer on films >= 2^20 pixels: full GPU suite + bench lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for wl in rungholt default; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --strong-spp 0 --steps 32 > gpurun_out/r2_dense_$wl.json 2> gpurun_out/r2_dense_$wl.log
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_dense_$wl.json"))
print("$wl value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
PY
done
