#!/bin/bash
# round 2, session 17: full GPU suite with the 96-register build, then one bench line per BASELINE workload
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for wl in rungholt sponza sponza_triple cornell default; do
  extra="--no-cpu-baseline --strong-spp 0"; [ $wl = rungholt ] && extra=""
  timeout 600 python bench.py --workload $wl $extra > gpurun_out/r2_bench_d_$wl.json 2> gpurun_out/r2_bench_d_$wl.log
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_d_$wl.json"))
print("$wl value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), d["roofline"]["stage_ms_per_step"], "frac", round(d["roofline"]["frac"],3), d["config"]["kernel_variant"][:30])
PY
done
