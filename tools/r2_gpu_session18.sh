#!/bin/bash
# round 2, session 18: where the end-to-end loop spends its time on this box (host_ms_per_step, pcie.d2h_gbs_while_rendering)
set -x
mkdir -p gpurun_out
nproc; cat /proc/loadavg; numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -8
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --strong-spp 0 --steps 20 > gpurun_out/r2_bench_e$i.json 2> gpurun_out/r2_bench_e$i.log
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_e$i.json"))
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), d["e2e"]["ms_per_step"], d["e2e"].get("host_ms_per_step"), d.get("pcie"), d.get("host_binding"))
PY
done
timeout 300 python bench.py --workload default --no-cpu-baseline --strong-spp 0 --steps 40 > gpurun_out/r2_bench_e_default.json 2> gpurun_out/r2_bench_e_default.log
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_e_default.json"))
print("default value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), d["e2e"]["ms_per_step"], d["e2e"].get("host_ms_per_step"), d.get("pcie"))
PY
cat /proc/loadavg
