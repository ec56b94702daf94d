#!/bin/bash
# round 2, session 16: non-Lambertian shade kernels at 4 / 5 / 6 resident CTAs per SM (128 / 96 / 80 registers for them and for the out-of-line BSDF
# functions they call: -DZL_WF_SHADE_MINB_OTHER=m -maxrregcount=r builds in csrc/alt/), against the default build (compiler's choice: 119-142 registers)
set -x
mkdir -p gpurun_out
L=zillumgl_b200/csrc/libzillum_cuda.so
cp $L /tmp/libzillum_cuda_default.so
for m in default 4 5 6; do
  if [ $m = default ]; then cp /tmp/libzillum_cuda_default.so $L; else cp zillumgl_b200/csrc/alt/libzillum_cuda_m$m.so $L; fi
  for wl in sponza sponza_triple rungholt; do
    python tools/sweep_env.py --workload $wl --steps 8 --no-megakernel --configs "default" --out gpurun_out/r2_sweep_shademinb_${m}_$wl.json 2>&1 | grep -v "^\[" | tail -1
  done
done
cp /tmp/libzillum_cuda_default.so $L
python - <<PY
import time, os, sys
sys.path.insert(0, ".")
import zillumgl_b200 as zl
s = zl.Scene.builtin("rungholt", 3840, 2160); s.set_device_bvh(True)
for i in range(2):
    t = time.perf_counter(); s.flatten(); print("flatten", round(time.perf_counter() - t, 3))
PY
