#!/bin/bash
# round 2: ncu --set full of the trace kernels of one steady-state pass of the triple tracer (the one workload without an ncu_traffic entry)
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --profile-from-start off -k regex:wfTrace -f -o /tmp/r2_trace_full_sponza_triple python tools/profile_pass.py --workload sponza_triple > gpurun_out/r2_ncu_trace_sponza_triple.log 2>&1
tail -1 gpurun_out/r2_ncu_trace_sponza_triple.log
ncu -i /tmp/r2_trace_full_sponza_triple.ncu-rep --page raw --csv > gpurun_out/r2_trace_full_sponza_triple.raw.csv
ls -la gpurun_out/r2_trace_full_sponza_triple.raw.csv
