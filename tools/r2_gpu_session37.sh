#!/bin/bash
# round 2: the default bench command of the final tree (timing of the whole command included)
mkdir -p gpurun_out
S=$(date +%s.%N); python bench.py > gpurun_out/r2_bench_default_cmd.json 2> gpurun_out/r2_bench_default_cmd.log; E=$(date +%s.%N); echo "Elapsed (wall clock) $(echo "$E - $S" | bc) s"
true
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_default_cmd.json"))
print("steps", d["steps"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), d["cpu_baseline"]["value"], d["strong_scaling"]["seconds"])
PY
