#!/bin/bash
# round 2, session 35: triple tracer with three pass pairs in flight (ZL_WF_TRIPLE_DEPTH=3)
mkdir -p gpurun_out
ZL_WF_TRIPLE_DEPTH=3 timeout 600 python -m pytest tests/test_gpu_integrators.py -x -q -m gpu -k "triple_tracer_pipelined" 2>&1 | tail -2
for d in 2 3; do
  ZL_WF_TRIPLE_DEPTH=$d timeout 300 python bench.py --workload sponza_triple --no-cpu-baseline --strong-spp 0 --steps 24 > gpurun_out/r2_tripledepth_$d.json 2> gpurun_out/r2_tripledepth_$d.log
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_tripledepth_$d.json"))
print("triple depth $d value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["film_mean_radiance"], d["e2e"]["last_frame_mean_radiance"])
PY
done
