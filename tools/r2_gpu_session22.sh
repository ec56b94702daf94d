#!/bin/bash
# round 2, session 22: read-back tests after deferring the D2H copy call behind the next pass launch; bench diagnostics
set -x
mkdir -p gpurun_out
true
T=$(date +%s)
ZL_DEBUG_DOWNLOAD_TIMING=1 timeout 300 python bench.py --no-cpu-baseline --strong-spp 0 --steps 32 > gpurun_out/r2_diag_$T.json 2> gpurun_out/r2_diag_$T.log
python - <<PY
import json
d=json.load(open("gpurun_out/r2_diag_$T.json"))
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["e2e"].get("host_ms_per_step"), d["e2e"].get("host_loadavg"), d.get("pcie"))
PY
grep "download timing" gpurun_out/r2_diag_$T.log | tail -2
for wl in default cornell; do
timeout 300 python bench.py --workload $wl --no-cpu-baseline --strong-spp 0 --steps 40 > gpurun_out/r2_diag_${wl}_$T.json 2> gpurun_out/r2_diag_${wl}_$T.log
python - <<PY
import json
d=json.load(open("gpurun_out/r2_diag_${wl}_$T.json"))
print("$wl value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["e2e"].get("host_ms_per_step"))
PY
done
