set -x
CFG="default;ZL_NODE_POLICY=1;ZL_STATE_POLICY=1;ZL_NODE_POLICY=1,ZL_STATE_POLICY=1;ZL_WF_TRACE_LOOP=6;ZL_WF_TRACE_LOOP=6,ZL_NODE_POLICY=1,ZL_STATE_POLICY=1"
python tools/sweep_env.py --workload rungholt --steps 6 --configs "$CFG" --out gpurun_out/r2_sweep_policy_rungholt.json 2>&1 | grep -v "^\[" | tail -8
python tools/sweep_env.py --workload sponza --steps 8 --configs "$CFG" --out gpurun_out/r2_sweep_policy_sponza.json 2>&1 | grep -v "^\[" | tail -8
for wl in rungholt sponza; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wfTraceSimpleKernel -f -o gpurun_out/r2_trace_full_$wl python tools/profile_pass.py --workload $wl > gpurun_out/r2_ncu_$wl.log 2>&1
  tail -2 gpurun_out/r2_ncu_$wl.log
done
ls -la gpurun_out/*.ncu-rep
