#!/bin/bash
# round 2: ncu --set full of the (lean) trace kernel launches of one steady-state pass of the headline workload, final tree
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --profile-from-start off -k regex:wfTrace -f -o /tmp/r2_trace_full_rungholt_final python tools/profile_pass.py --workload rungholt > gpurun_out/r2_ncu_trace_rungholt_final.log 2>&1
tail -1 gpurun_out/r2_ncu_trace_rungholt_final.log
ncu -i /tmp/r2_trace_full_rungholt_final.ncu-rep --page raw --csv > gpurun_out/r2_trace_full_rungholt_final.raw.csv
ls -la gpurun_out/r2_trace_full_rungholt_final.raw.csv
