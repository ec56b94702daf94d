#!/bin/bash
# round 2, session 34: ring of 2 / 3 / 4 passes in flight for the pipelined light tracer
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_integrators.py -x -q -m gpu -k "light or cornell" 2>&1 | tail -2
for d in 2 3 4; do
  ZL_WF_PIPE_DEPTH=$d timeout 300 python bench.py --workload cornell --no-cpu-baseline --strong-spp 0 --steps 32 > gpurun_out/r2_lightdepth_$d.json 2> gpurun_out/r2_lightdepth_$d.log
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_lightdepth_$d.json"))
print("cornell depth $d value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["film_mean_radiance"], d["e2e"]["last_frame_mean_radiance"])
PY
done
