CFG="default;ZL_WF_TRACE_LOOP=7;ZL_WF_TRACE_LOOP=7,ZL_WF_TRACE_MINB=10"
python tools/sweep_env.py --workload rungholt --steps 6 --no-megakernel --configs "$CFG" --out gpurun_out/r2_sweep_compact_rungholt.json 2>&1 | grep -v "^\[" | tail -4
python tools/sweep_env.py --workload sponza --steps 8 --configs "$CFG" --out gpurun_out/r2_sweep_compact_sponza.json 2>&1 | grep -v "^\[" | tail -5
python tools/sweep_env.py --workload sponza_triple --steps 6 --configs "$CFG" --out gpurun_out/r2_sweep_compact_sponza_triple.json 2>&1 | grep -v "^\[" | tail -5
python tools/sweep_env.py --workload cornell --steps 8 --configs "$CFG" --out gpurun_out/r2_sweep_compact_cornell.json 2>&1 | grep -v "^\[" | tail -5
