#!/bin/bash
# round 2, session 24: lean instantiations of the default trace kernel (A/B walks compiled out, the scene's step fixed)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_integrators.py -x -q -m gpu -k "wavefront_variant or bvh2 or random_ray or shadow or near_zero or pipelined_passes_are" 2>&1 | tail -3
for wl in rungholt sponza sponza_triple default; do
  python tools/sweep_env.py --workload $wl --steps 8 --no-megakernel --configs "default;ZL_NODE_POLICY=1" --out gpurun_out/r2_sweep_lean_$wl.json 2>&1 | grep -v "^\[" | tail -2
done
