#!/bin/bash
# round 2, session 28b: 10 (new default with two passes in flight) against 12 resident trace CTAs per SM, per workload
mkdir -p gpurun_out
for wl in rungholt sponza default "sponza_triple --variant 2"; do
  tag=$(echo $wl | cut -d' ' -f1)
  for c in default 12; do
    if [ $c = default ]; then unset ZL_WF_TRACE_CTAS_PER_SM; else export ZL_WF_TRACE_CTAS_PER_SM=$c; fi
    timeout 300 python bench.py --workload $wl --no-cpu-baseline --strong-spp 0 --steps 24 > gpurun_out/r2_ctas_${tag}_$c.json 2> gpurun_out/r2_ctas_${tag}_$c.log
    python - <<PY
import json
d=json.load(open("gpurun_out/r2_ctas_${tag}_$c.json"))
print("$tag ctas $c value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
PY
  done
done
