// zl_scene.cuh — device-resident scene in its B200 layout (see DESIGN.md "Data layout in HBM").
//
// The reference binds 13 buffer textures and walks them with dependent texelFetch chains
// (hit table -> bounds, indices -> vertices).  Here everything the hot loops
// touch is re-packed once at upload into 16-byte vectors with the indirections resolved:
//
//   nodes   [6 faces][bvhSize] x 2 float4 : {pMin.x, pMin.y, pMax.x, pMax.y}, {pMin.z, pMax.z, bits(prim|-1), bits(missLink)}
//           one 32-byte record per THREADED entry, in traversal order: the hit link (k+1)
//           is the next record in memory, the miss link is in the record itself.  The order of the six
//           bounds is the one the packed-FP32 slab test wants (zl_traverse.cuh): the 256-bit load leaves
//           {pMin.x, pMin.y}, {pMax.x, pMax.y}, {pMin.z, pMax.z} in aligned register pairs (FADD2 / FMUL2 operands).
//           Every writer goes through packNodeRecord, every reader through unpackNodeRecord / loadNode.
//   triPos  [T] x 3 float4 : {a.xyz, ta.x}, {b.xyz, tb.x}, {c.xyz, tc.x}     (intersection + shading)
//   triNrm  [T] x 3 float4 : {na.xyz, ta.y}, {nb.xyz, tb.y}, {nc.xyz, tc.y}  (shading only)
//           vertices are gathered per triangle (no index fetch); the uv pair rides in the
//           w lanes that float3 padding would waste.
#pragma once
#include <cuda_fp16.h>
#include "zl_math.cuh"
#include "../../include/zillum_cuda.h"

namespace zl {

// the one definition of the threaded node record (32 bytes)
__host__ __device__ inline void packNodeRecord(float lox, float loy, float loz, float hix, float hiy, float hiz, int prim, int miss, float4& r0, float4& r1) {
    union { int i; float f; } p, m;
    p.i = prim; m.i = miss;
    r0 = make_float4(lox, loy, hix, hiy);
    r1 = make_float4(loz, hiz, p.f, m.f);
}
// -> the {pMin.xyz, bits(prim)}, {pMax.xyz, bits(miss)} view the generic code works with
__host__ __device__ inline void unpackNodeRecord(const float4& r0, const float4& r1, float4& lo, float4& hi) {
    lo = make_float4(r0.x, r0.y, r1.x, r1.z);
    hi = make_float4(r0.z, r0.w, r1.y, r1.w);
}

struct DScene {
    const float4* __restrict__ nodes;
    const float4* __restrict__ triPos;
    const float4* __restrict__ bvh2;          // child-boxes-in-the-parent records of the builder's tree, 64 B per interior node (zl_traverse.cuh traverseBvh2); nullptr: threaded walk only
    float3 rootLo, rootHi;                    // the root's own box (first step of the walk)
    const float4* __restrict__ top;           // top kTopDepth levels of the six orderings, compact with explicit links (zl_traverse.cuh, buildStagedTopKernel)
    const float4* __restrict__ triNrm;
    const int*    __restrict__ matTex;        // objPrimCount
    const float4* __restrict__ materials;     // 4 per material
    const float4* __restrict__ lightPowProb;  // {power.rgb, aliasProb}
    const int*    __restrict__ lightAlias;
    const uchar4* __restrict__ texels;        // sRGB8 layers, texMaxW x texMaxH each
    const float2* __restrict__ texScale;
    const float*  __restrict__ srgbLut;       // 256 entries
    const ushort4* __restrict__ env;          // RGB16F texels (w unused), envW x envH
    const int2*   __restrict__ envAlias;      // {alias, bits(prob)}, (envW+1) x envH
    const float2* __restrict__ noise;         // noiseW x noiseH
    const uint32_t* __restrict__ sobol;       // 256 x 32 generator matrices
    unsigned long long* counters;             // instrumented build only (NULL otherwise): rays, nodes, tris, shades, splats, paths
    int bvhSize, numTriangles, objPrimCount, numLightTriangles, numMaterials;
    int numTextures, texMaxW, texMaxH, envW, envH, noiseW, noiseH;
    float lightSum, envSum;
    int nodePolicy;                           // node-record loads: 0 = default cache priorities, 1 = evict_last in L1 and L2 (A/B switch ZL_NODE_POLICY)
    int statePolicy;                          // trace kernel's path-state / queue accesses: 0 = default, 1 = streaming (evict_first) (A/B switch ZL_STATE_POLICY)
    int octantWalk;                           // 1 (default): octant-uniform warps take the specialised walks of traverseWarp; 0: always the general walk (A/B switch ZL_OCTANT_WALK)
};

// GL LINEAR + REPEAT footprint: texel centres at (i + 0.5) / size (Texture.cpp:131)
struct Bilerp { int i0, i1; float f; };
ZL_DEV Bilerp bilerpRepeat(float u, int size) {
    float x = u * (float)size - 0.5f;
    float fl = floorf(x);
    Bilerp b;
    b.f = x - fl;
    int m = (int)fl % size;
    if (m < 0) m += size;
    b.i0 = m;
    b.i1 = (m + 1 == size) ? 0 : m + 1;
    return b;
}

ZL_DEV float3 envTexel(const DScene& S, int x, int y) {
    ushort4 h = __ldg(&S.env[(size_t)y * S.envW + x]);
    return f3(__half2float(__ushort_as_half(h.x)), __half2float(__ushort_as_half(h.y)), __half2float(__ushort_as_half(h.z)));
}
// texture(uEnvMap, uv): bilinear in FP32 over the fp16 texels
ZL_DEV float3 sampleEnv(const DScene& S, float2 uv) {
    Bilerp bx = bilerpRepeat(uv.x, S.envW), by = bilerpRepeat(uv.y, S.envH);
    float3 a = envTexel(S, bx.i0, by.i0) * (1.0f - bx.f) + envTexel(S, bx.i1, by.i0) * bx.f;
    float3 b = envTexel(S, bx.i0, by.i1) * (1.0f - bx.f) + envTexel(S, bx.i1, by.i1) * bx.f;
    return a * (1.0f - by.f) + b * by.f;
}
ZL_DEV float2 sampleNoise(const DScene& S, float2 uv) {
    Bilerp bx = bilerpRepeat(uv.x, S.noiseW), by = bilerpRepeat(uv.y, S.noiseH);
    float2 t00 = __ldg(&S.noise[(size_t)by.i0 * S.noiseW + bx.i0]), t10 = __ldg(&S.noise[(size_t)by.i0 * S.noiseW + bx.i1]);
    float2 t01 = __ldg(&S.noise[(size_t)by.i1 * S.noiseW + bx.i0]), t11 = __ldg(&S.noise[(size_t)by.i1 * S.noiseW + bx.i1]);
    float2 a = t00 * (1.0f - bx.f) + t10 * bx.f;
    float2 b = t01 * (1.0f - bx.f) + t11 * bx.f;
    return a * (1.0f - by.f) + b * by.f;
}
ZL_DEV float3 albedoTexel(const DScene& S, int layer, int x, int y) {
    uchar4 t = __ldg(&S.texels[((size_t)layer * S.texMaxH + y) * S.texMaxW + x]);
    return f3(__ldg(&S.srgbLut[t.x]), __ldg(&S.srgbLut[t.y]), __ldg(&S.srgbLut[t.z]));
}
// texture2DArray(uTextures, vec3(uv, layer)): sRGB decode per texel, then bilinear (Texture.cpp:146-161)
ZL_DEV float3 sampleAlbedo(const DScene& S, float2 uv, int layer) {
    if (layer < 0 || layer >= S.numTextures || S.texels == nullptr) return f3(0.0f);
    Bilerp bx = bilerpRepeat(uv.x, S.texMaxW), by = bilerpRepeat(uv.y, S.texMaxH);
    float3 a = albedoTexel(S, layer, bx.i0, by.i0) * (1.0f - bx.f) + albedoTexel(S, layer, bx.i1, by.i0) * bx.f;
    float3 b = albedoTexel(S, layer, bx.i0, by.i1) * (1.0f - bx.f) + albedoTexel(S, layer, bx.i1, by.i1) * bx.f;
    return a * (1.0f - by.f) + b * by.f;
}

}  // namespace zl
