#!/bin/bash
# round 2, session 11: full ncu capture of EVERY kernel of one steady-state pass on the small-film workloads (C1 default, C2 cornell);
# the raw page is exported as csv on the box (the .ncu-rep files exceed what gpurun copies back)
set -x
mkdir -p gpurun_out
for wl in cornell default; do
  timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r2_pass_full_$wl python tools/profile_pass.py --workload $wl > gpurun_out/r2_ncu_pass_$wl.log 2>&1
  tail -1 gpurun_out/r2_ncu_pass_$wl.log
  ncu -i /tmp/r2_pass_full_$wl.ncu-rep --page raw --csv > gpurun_out/r2_pass_full_$wl.raw.csv
done
ls -la gpurun_out/*.csv
