#!/bin/bash
# round 2, session 12: the trace-loop forms (defer 1-3, refill 4, dual 5, compaction 7) on the small scenes, where the ncu capture shows the trace
# kernel issue-bound at 10-11 live lanes per instruction (profiles/r2_pass_full_cornell.csv)
set -x
mkdir -p gpurun_out
CFG="default;ZL_WF_TRACE_LOOP=1;ZL_WF_TRACE_LOOP=2;ZL_WF_TRACE_LOOP=3;ZL_WF_TRACE_LOOP=4,ZL_WF_REFILL_FROM=0;ZL_WF_TRACE_LOOP=4,ZL_WF_REFILL_FROM=0,ZL_WF_REFILL_AT=16;ZL_WF_TRACE_LOOP=4,ZL_WF_REFILL_FROM=0,ZL_WF_REFILL_AT=24,ZL_WF_ROUND_STEPS=8;ZL_WF_TRACE_LOOP=5"
python tools/sweep_env.py --workload cornell --steps 8 --no-megakernel --configs "$CFG" --out gpurun_out/r2_sweep_loops_cornell.json 2>&1 | grep -v "^\[" | tail -12
python tools/sweep_env.py --workload default --steps 8 --no-megakernel --configs "$CFG" --out gpurun_out/r2_sweep_loops_default.json 2>&1 | grep -v "^\[" | tail -12
