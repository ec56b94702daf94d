// Sobol generator + per-pixel seed image (reference: src/core/Sampler.{h,cpp}).
// The reference precomputes a 131072x256 uint32 table (128 MiB, Sampler.cpp:48-64); here
// the generator matrices (32 KiB) are handed to the device, which evaluates the same
// sobolSample() for the one row a pass needs.
#pragma once
#include <cstdint>
#include <vector>

namespace zillum {
namespace Sampler {

const int SobolMatricesDim = 256;
const int SobolMatricesSize = 32;
extern const uint32_t SobolMatrices[SobolMatricesDim * SobolMatricesSize];

uint32_t sobolSample(uint32_t index, int dim, uint32_t scramble = 0);   // Sampler.cpp:19-28
// RG32F seed image in [0,1): the reference fills it from std::default_random_engine
// (Sampler.cpp:66-80), which is implementation-defined; we use a fixed SplitMix64 stream
// (seed 0) so that every build and the oracle see the same image.
std::vector<float> genNoiseTexture(int width, int height, uint64_t seed = 0);

}  // namespace Sampler
}  // namespace zillum
