#!/bin/bash
# round 2, session 21: end-to-end diagnostics on whatever box this lands on (run several times: the pod's boxes differ)
mkdir -p gpurun_out
T=$(date +%s)
cat /proc/loadavg; nproc
ZL_DEBUG_DOWNLOAD_TIMING=1 timeout 300 python bench.py --no-cpu-baseline --strong-spp 0 --steps 32 > gpurun_out/r2_diag_$T.json 2> gpurun_out/r2_diag_$T.log
python - <<PY
import json
d=json.load(open("gpurun_out/r2_diag_$T.json"))
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["e2e"].get("host_ms_per_step"), d["e2e"].get("host_loadavg"), d.get("pcie"))
PY
cat /proc/loadavg; grep "download timing" gpurun_out/r2_diag_$T.log | tail -3
