#!/bin/bash
# round 2, verification of the final tree on one B200: full GPU suite, smoke(), the default bench line, the reference arm, the ncu launch list
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.log
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_final.json"))
r=d["roofline"]
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), "frac", round(r["frac"],3), "dram_frac", r["dram_frac"], "l2_frac", r["l2_frac"], "bound", r["bound"], "launches", d["gpu_launches"])
print(r["stage_ms_per_step"], r["share_of_step"]); print(d["cpu_baseline"]); print(d["e2e"].get("host_ms_per_step"), d.get("pcie")); print(d["strong_scaling"]["seconds"], d["strong_scaling"]["render_and_reduce_seconds"]); print(d["clocks"])
PY
for wl in sponza sponza_triple cornell default; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --strong-spp 0 --steps 24 > gpurun_out/r2_bench_final_$wl.json 2> gpurun_out/r2_bench_final_$wl.log
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_final_$wl.json"))
print("$wl value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3))
PY
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_final_reference.json 2> gpurun_out/r2_bench_final_reference.log
cat gpurun_out/r2_bench_final_reference.json | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --strong-spp 0 > gpurun_out/r2_launches_final.log 2>&1
wc -l gpurun_out/r2_launches_final.csv
