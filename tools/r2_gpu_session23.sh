#!/bin/bash
# round 2, session 23 (2 GPUs): N = 2 bench line after software-pipelining the read-back of the reduce-scattered frames
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_g_2gpu.json 2> gpurun_out/r2_bench_g_2gpu.log
tail -2 gpurun_out/r2_bench_g_2gpu.log
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_g_2gpu.json"))
print("N=2 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), d["e2e"]["ms_per_step"], d["e2e"]["last_frame_mean_radiance"], d["film_mean_radiance"], d["strong_scaling"]["seconds"])
PY
