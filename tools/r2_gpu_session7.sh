python -m pytest tests/test_gpu_traversal.py tests/test_golden.py tests/test_ref_parity.py -m gpu -x -q 2>&1 | tail -5
python -m pytest tests/test_gpu_integrators.py -m gpu -x -q -k "bit_identical or single_pass or full_size" 2>&1 | tail -4
CFG="ZL_BVH2_WALK=0;ZL_BVH2_WALK=1"
for wl in rungholt sponza sponza_triple cornell default; do
python tools/sweep_env.py --workload $wl --steps 6 --configs "$CFG" --out gpurun_out/r2_sweep_bvh2_$wl.json 2>&1 | grep -v "^\[" | tail -3
done
