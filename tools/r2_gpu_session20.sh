#!/bin/bash
# round 2, session 20 (8 GPUs): N = 8 and N = 4 bench lines through torchrun
set -x
mkdir -p gpurun_out
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_bench_f_${n}gpu.json 2> gpurun_out/r2_bench_f_${n}gpu.log
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_f_${n}gpu.json"))
print("N=$n value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],3), d["multi_gpu_breakdown"], d["strong_scaling"]["seconds"], d["strong_scaling"]["render_and_reduce_seconds"], d["strong_scaling"]["scene_prep"], d.get("pcie"))
PY
done
