#!/bin/bash
# round 2, session 31: 4-way unrolled sort scatter; the new A/B configurations in the bit-identity test
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_integrators.py -x -q -m gpu -k "wavefront_variant or sort_bits or pipelined_passes_are or full_size" 2>&1 | tail -3
for wl in rungholt sponza; do
  python tools/sweep_env.py --workload $wl --steps 8 --no-megakernel --configs "default" --out gpurun_out/r2_sweep_scatter4_$wl.json 2>&1 | grep -v "^\[" | tail -1
done
