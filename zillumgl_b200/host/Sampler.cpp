#include "Sampler.h"

namespace zillum {
namespace Sampler {

const uint32_t SobolMatrices[SobolMatricesDim * SobolMatricesSize] = {
#include "sobol_matrices.inc"
};

uint32_t sobolSample(uint32_t index, int dim, uint32_t scramble) {
    uint32_t r = scramble;
    const uint32_t* col = SobolMatrices + dim * SobolMatricesSize;
    for (; index != 0; index >>= 1, col++)
        if (index & 1u) r ^= *col;
    return r;
}

std::vector<float> genNoiseTexture(int width, int height, uint64_t seed) {
    std::vector<float> data((size_t)width * height * 2);
    uint64_t state = seed;
    for (auto& v : data) {
        state += 0x9e3779b97f4a7c15ull;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        z ^= z >> 31;
        v = (float)(z >> 40) * (1.0f / 16777216.0f);   // 24 random bits -> [0,1)
    }
    return data;
}

}  // namespace Sampler
}  // namespace zillum
