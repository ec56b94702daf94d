#!/bin/bash
# round 2, session 40b: stage kernels at 9 CTAs per SM (56 registers) against 6 (80) on the other workloads
mkdir -p gpurun_out
L=zillumgl_b200/csrc/libzillum_cuda.so
cp $L /tmp/libzillum_cuda_default.so
for m in default s9; do
  if [ $m = default ]; then cp /tmp/libzillum_cuda_default.so $L; else cp zillumgl_b200/csrc/alt/libzillum_cuda_$m.so $L; fi
  for wl in sponza sponza_triple cornell default; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --strong-spp 0 --steps 32 > gpurun_out/r2_stageminb_${wl}_$m.json 2> gpurun_out/r2_stageminb_${wl}_$m.log
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_stageminb_${wl}_$m.json"))
print("$wl stage minb $m value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
PY
  done
done
cp /tmp/libzillum_cuda_default.so $L
