#!/bin/bash
# round 2, session 29: sequential (1) against pipelined (2) schedule for the splatting integrators, device-timed and end to end, with the new read-back path
mkdir -p gpurun_out
for wl in cornell sponza_triple; do
  for v in 1 2; do
    timeout 300 python bench.py --workload $wl --variant $v --no-cpu-baseline --strong-spp 0 --steps 32 > gpurun_out/r2_v_${wl}_$v.json 2> gpurun_out/r2_v_${wl}_$v.log
    python - <<PY
import json
d=json.load(open("gpurun_out/r2_v_${wl}_$v.json"))
print("$wl variant $v value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["e2e"]["host_ms_per_step"])
PY
  done
done
