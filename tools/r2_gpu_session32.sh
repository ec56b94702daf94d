#!/bin/bash
# round 2, session 32b: two / three / four passes in flight for the pipelined path tracer (ZL_WF_PIPE_DEPTH)
mkdir -p gpurun_out
ZL_WF_PIPE_DEPTH=4 timeout 600 python -m pytest tests/test_gpu_integrators.py -x -q -m gpu -k "pipelined_passes_are or two_readbacks or external_film" 2>&1 | tail -2
for wl in rungholt default sponza; do
  for d in 2 3 4; do
    ZL_WF_PIPE_DEPTH=$d timeout 300 python bench.py --workload $wl --no-cpu-baseline --strong-spp 0 --steps 24 > gpurun_out/r2_depth_${wl}_$d.json 2> gpurun_out/r2_depth_${wl}_$d.log
    python - <<PY
import json
d=json.load(open("gpurun_out/r2_depth_${wl}_$d.json"))
print("$wl depth $d value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
PY
  done
done
