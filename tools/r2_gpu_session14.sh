#!/bin/bash
# round 2, session 14: full ncu capture of EVERY kernel of one steady-state pass: headline workload (rungholt), sponza, cornell (two rays per lane now);
# raw page exported as csv on the box.  Then the sort-bits test and the C2 bench line with the per-integrator default variant.
set -x
mkdir -p gpurun_out
for wl in rungholt cornell sponza; do
  timeout 1200 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r2_pass_full_$wl python tools/profile_pass.py --workload $wl > gpurun_out/r2_ncu_pass_$wl.log 2>&1
  tail -1 gpurun_out/r2_ncu_pass_$wl.log
  ncu -i /tmp/r2_pass_full_$wl.ncu-rep --page raw --csv > gpurun_out/r2_pass_full_$wl.raw.csv
  rm -f /tmp/r2_pass_full_$wl.ncu-rep
done
ls -la gpurun_out/*.csv
timeout 600 python -m pytest tests/test_gpu_integrators.py -x -q -m gpu -k "sort_bits" 2>&1 | tail -3
timeout 300 python bench.py --workload cornell --steps 40 --warmup 5 --no-cpu-baseline --strong-spp 0 > gpurun_out/r2_c2.json 2> gpurun_out/r2_c2.log
python - <<PY
import json
d=json.load(open("gpurun_out/r2_c2.json"))
print("cornell value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), d.get("pcie"), d.get("host_binding"))
PY
