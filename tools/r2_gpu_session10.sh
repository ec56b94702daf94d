#!/bin/bash
# round 2, session 10: CUDA-graph replay of a wavefront pass (kernelVariant 3) — parity tests, then C1 / C2 / 720p with variants 1, 2, 3
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_integrators.py -x -q -m gpu -k "graph_replayed" 2>&1 | tail -5
for wl in default cornell; do
  for v in 1 2 3; do
    timeout 300 python bench.py --workload $wl --variant $v --steps 40 --warmup 5 --no-cpu-baseline --strong-spp 0 > gpurun_out/r2_graph_${wl}_v$v.json 2> gpurun_out/r2_graph_${wl}_v$v.log
    python - <<PY
import json
d=json.load(open("gpurun_out/r2_graph_${wl}_v$v.json"))
print("$wl v$v value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), "launches", d.get("gpu_launches"))
PY
  done
done
for v in 2 3; do
  timeout 300 python bench.py --workload rungholt --variant $v --steps 10 --warmup 3 --no-cpu-baseline --strong-spp 0 > gpurun_out/r2_graph_rungholt_v$v.json 2> gpurun_out/r2_graph_rungholt_v$v.log
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_graph_rungholt_v$v.json"))
print("rungholt v$v value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), "launches", d.get("gpu_launches"))
PY
done
