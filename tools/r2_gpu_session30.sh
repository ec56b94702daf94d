#!/bin/bash
# round 2, session 30 (8 GPUs): the scaling run of the final tree, N = 8, 4, 2, 1 back to back on one box
set -x
mkdir -p gpurun_out
for n in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_scale_${n}gpu.json 2> gpurun_out/r2_scale_${n}gpu.log
done
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale_1gpu.json 2> gpurun_out/r2_scale_1gpu.log
python - <<PY
import json
for n in (1,2,4,8):
    d=json.load(open(f"gpurun_out/r2_scale_{n}gpu.json"))
    print("N=%d value %.1f e2e %.1f ms %.3f strong %.2f s (render+reduce %.2f)" % (n, d["value"], d["e2e"]["value"], d["ms_per_step"], d["strong_scaling"]["seconds"], d["strong_scaling"]["render_and_reduce_seconds"]), d["multi_gpu_breakdown"])
PY
