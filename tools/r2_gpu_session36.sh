#!/bin/bash
# round 2, last check of the final tree: full GPU suite + smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --workload sponza_triple --no-cpu-baseline --strong-spp 0 --steps 24 > gpurun_out/r2_bench_final_sponza_triple.json 2> /dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_final_sponza_triple.json"))
print("sponza_triple value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1))
PY
