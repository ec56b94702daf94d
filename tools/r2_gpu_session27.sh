#!/bin/bash
# round 2, session 27: PIPE without the L2 prefetch of the next chunk's path state (claim + queue entries pipelined only)
set -x
mkdir -p gpurun_out
L=zillumgl_b200/csrc/libzillum_cuda.so
cp $L /tmp/libzillum_cuda_default.so
cp zillumgl_b200/csrc/alt/libzillum_cuda_nopf.so $L
for wl in rungholt sponza; do
  python tools/sweep_env.py --workload $wl --steps 8 --no-megakernel --configs "default;ZL_WF_TRACE_PIPE=1" --out gpurun_out/r2_sweep_pipe_nopf_$wl.json 2>&1 | grep -v "^\[" | tail -2
done
cp /tmp/libzillum_cuda_default.so $L
