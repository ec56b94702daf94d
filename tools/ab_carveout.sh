# A/B: preferred shared-memory carve-out of the default trace kernel (ZL_WF_L1_CARVEOUT, percent; unset = driver default)
for c in none 6 12 25 50 75 none; do
  if [ $c = none ]; then unset ZL_WF_L1_CARVEOUT; else export ZL_WF_L1_CARVEOUT=$c; fi
  echo "carveout=$c"; python tools/sweep_env.py --workload ${1:-rungholt} --no-megakernel --steps 8 --configs "default" --out gpurun_out/sweep_carve_${1:-rungholt}_$c.json 2>&1 | tail -1
done
