// zl_instrumented.cu — the four pass kernels compiled a second time with visit counters
// (ZL_INSTRUMENT): identical arithmetic, plus atomic event counts into DScene::counters.
// Used once per configuration, outside any timed region, to obtain the ALGORITHMIC byte
// count of a pass (36 B per hit-table entry visited, 48 B per leaf triangle test, 108 B per
// shading point, film traffic) for the roofline figure bench.py reports.
#define ZL_INSTRUMENT 1
#define zl zlc   // every symbol of this translation unit lives in its own namespace (relocatable device code links both)
#include "zl_kernels.cuh"

namespace zl {

int launchCountedPass(int kind, const DScene& S, const ZlRenderParams& U, float4* film, cudaStream_t stream) {
    dim3 grid((U.filmW + kTileW - 1) / kTileW, (U.filmH + kTileH - 1) / kTileH);
    long long total = (long long)ZL_LIGHT_GROUP_SIZE * U.blocksOnePass;
    unsigned blocks = (unsigned)((total + kLightBlock - 1) / kLightBlock);
    switch (kind) {
    case 0: pathPassKernelCounted<<<grid, kPixelBlock, 0, stream>>>(S, U, film); break;
    case 1: if (total > 0) lightPassKernelCounted<<<blocks, kLightBlock, 0, stream>>>(S, U, film, total); break;
    case 2: triplePtPassKernelCounted<<<grid, kPixelBlock, 0, stream>>>(S, U, film); break;
    case 3: if (total > 0) tripleLptPassKernelCounted<<<blocks, kLightBlock, 0, stream>>>(S, U, film, total); break;
    default: return 1;
    }
    return 0;
}

}  // namespace zl
