#!/bin/bash
# The five BASELINE.json configurations at their stated sample counts through the headless CLI (C++ host classes -> C ABI),
# one JSON line each (seconds include every pass and the final flush; scene preparation is reported separately by the CLI).
#   C1 default scene, path, 1280x720, 64 spp            C2 Cornell box, light tracer, 1920x1080, 1024 spp-equivalent
#   C3 Sponza-class + HDR env, path, 1920x1080 (256)    C4 Sponza-class, triple tracer, 1920x1080, 4096 spp
#   C5 Rungholt-class, path, 3840x2160 (256)
set -e
cd "$(dirname "$0")/.."
R=zillumgl_b200/host/zillum_render
O=${1:-gpurun_out}
mkdir -p "$O"
$R builtin:default      --integrator path   --size 1280x720  --spp 64   --out $O/c1_default.pfm
$R builtin:cornell      --integrator light  --size 1920x1080 --spp 1024 --out $O/c2_cornell_light.pfm
$R builtin:sponza       --integrator path   --size 1920x1080 --spp 256  --out $O/c3_sponza_path.pfm
$R builtin:sponza_light --integrator triple --size 1920x1080 --spp 4096 --out $O/c4_sponza_triple.pfm --png $O/c4_sponza_triple.png
$R builtin:rungholt     --integrator path   --size 3840x2160 --spp 256  --out $O/c5_rungholt_path.pfm
