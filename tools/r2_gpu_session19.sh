#!/bin/bash
# round 2, session 19 (2 GPUs): the N = 2 bench line through torchrun, reference arm under torchrun, generator timing
set -x
mkdir -p gpurun_out
python - <<PY
import time, sys
sys.path.insert(0, ".")
import zillumgl_b200 as zl
for i in range(2):
    t = time.perf_counter(); s = zl.Scene.builtin("rungholt", 3840, 2160); print("builtin", round(time.perf_counter() - t, 3))
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/r2_bench_f_2gpu.json 2> gpurun_out/r2_bench_f_2gpu.log
tail -2 gpurun_out/r2_bench_f_2gpu.log
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_f_2gpu.json"))
print("N=2 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), d["multi_gpu_breakdown"], d["strong_scaling"]["seconds"], d["strong_scaling"]["render_and_reduce_seconds"], d["strong_scaling"]["scene_prep"])
PY
