#!/bin/bash
# round 2, session 13: full GPU suite with the scene-chosen two-rays-per-lane trace kernel on Cornell-class scenes, then C2 / C1 bench lines
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for wl in cornell default; do
  for v in 1 2; do
    timeout 300 python bench.py --workload $wl --variant $v --steps 40 --warmup 5 --no-cpu-baseline --strong-spp 0 > gpurun_out/r2_dual_${wl}_v$v.json 2> gpurun_out/r2_dual_${wl}_v$v.log
    python - <<PY
import json
d=json.load(open("gpurun_out/r2_dual_${wl}_v$v.json"))
print("$wl v$v value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), "launches", d.get("gpu_launches"), d["roofline"]["stage_ms_per_step"])
PY
  done
done
