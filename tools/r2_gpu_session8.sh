CFG="ZL_BVH2_WALK=0;ZL_BVH2_WALK=1;ZL_BVH2_WALK=1,ZL_BVH2_MINB=10;ZL_BVH2_WALK=1,ZL_BVH2_MINB=8"
for wl in rungholt sponza; do
python tools/sweep_env.py --workload $wl --steps 6 --no-megakernel --configs "$CFG" --out gpurun_out/r2_sweep_bvh2c_$wl.json 2>&1 | grep -v "^\[" | tail -4
done
