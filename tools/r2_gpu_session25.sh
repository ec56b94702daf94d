#!/bin/bash
# round 2, session 25: lean trace kernel at 10 / 11 / 12 CTAs per SM (48 / 40 / 40 registers); explicit-ray-set traversal with the lean walks
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_traversal.py -x -q -m gpu 2>&1 | tail -2
for wl in rungholt sponza; do
  python tools/sweep_env.py --workload $wl --steps 8 --no-megakernel --configs "default;ZL_WF_TRACE_MINB=10;ZL_WF_TRACE_MINB=11" --out gpurun_out/r2_sweep_leanminb_$wl.json 2>&1 | grep -v "^\[" | tail -3
done
timeout 300 python bench.py --no-cpu-baseline --strong-spp 0 --steps 12 > gpurun_out/r2_bench_h.json 2> gpurun_out/r2_bench_h.log
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_h.json"))
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "traversal Mrays/s", round(d["traversal_mrays_per_s"],1))
PY
