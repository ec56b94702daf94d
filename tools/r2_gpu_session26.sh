#!/bin/bash
# round 2, session 26: software-pipelined chunk heads in the trace kernel (ZL_WF_TRACE_PIPE=1)
set -x
mkdir -p gpurun_out
for wl in rungholt sponza sponza_triple default; do
  python tools/sweep_env.py --workload $wl --steps 8 --no-megakernel --configs "default;ZL_WF_TRACE_PIPE=1" --out gpurun_out/r2_sweep_pipe_$wl.json 2>&1 | grep -v "^\[" | tail -2
done
