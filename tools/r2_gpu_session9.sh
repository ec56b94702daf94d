CFG="ZL_OCTANT_WALK=1;ZL_OCTANT_WALK=2;ZL_OCTANT_WALK=0"
for wl in sponza rungholt sponza_triple default cornell; do
python tools/sweep_env.py --workload $wl --steps 8 --no-megakernel --configs "$CFG" --out gpurun_out/r2_sweep_scalar_$wl.json 2>&1 | grep -v "^\[" | tail -3
done
python -m pytest tests/test_gpu_traversal.py -m gpu -x -q 2>&1 | tail -3
