// ORACLE — test infrastructure only (see zo_vec.h header).  Pinned to oracle/_ref by tests/test_ref_parity.py.
// zo_scene.h — the data contract of SURVEY.md App. A held in host vectors, plus the GL
// sampling semantics (texelFetch / texture LINEAR+REPEAT / sRGB array) the shaders rely on.
#pragma once
#include <vector>
#include "zo_vec.h"
#include "../include/zillum_cuda.h"

namespace zo {

struct Scene {
    std::vector<float> vertices, normals, texcoords, bounds, materials, lightPower, lightProb;
    std::vector<float> texUVScale, envMap /* rounded to fp16 */, envAliasProb, noise;
    std::vector<uint32_t> indices, sobolMatrices;
    std::vector<int32_t> hitTable, matTexIndices, lightAlias, envAlias;
    std::vector<uint8_t> texels;
    float srgbLut[256];
    int numVertices = 0, numTexcoords = 0, numTriangles = 0, bvhSize = 0, objPrimCount = 0;
    int numMaterials = 0, numLightTriangles = 0, numTextures = 0, texMaxW = 0, texMaxH = 0;
    int envW = 1, envH = 1, noiseW = 1, noiseH = 1;
    float lightSum = 0.0f, envSum = 0.0f;

    explicit Scene(const ZlSceneDesc& d);

    // texelFetch(samplerBuffer, i): exact element; out of range returns 0 (robust GL).
    vec3 fetchVertex(int i) const { return fetch3(vertices, i, numVertices); }
    vec3 fetchNormal(int i) const { return fetch3(normals, i, numVertices); }
    vec2 fetchTexCoord(int i) const {
        if (i < 0 || i >= numTexcoords) return vec2(0.0f);
        return vec2(texcoords[2 * i], texcoords[2 * i + 1]);
    }
    int fetchIndex(int i) const { return (int)indices[i]; }
    vec3 fetchBound(int i) const { return vec3(bounds[3 * i], bounds[3 * i + 1], bounds[3 * i + 2]); }
    vec4 fetchMaterial(int i) const {
        return vec4(materials[4 * i], materials[4 * i + 1], materials[4 * i + 2], materials[4 * i + 3]);
    }
    // uMatTypes: the same buffer read through an isamplerBuffer -> raw bits (material_loader.glsl:5)
    int fetchMatTypeBits(int texel, int comp) const { return (int)floatBits(materials[4 * texel + comp]); }
    vec3 fetchLightPower(int i) const { return vec3(lightPower[3 * i], lightPower[3 * i + 1], lightPower[3 * i + 2]); }

    // texture(sampler2D, uv) with LINEAR filter + REPEAT wrap (Texture.cpp:131), 3 / 2 channels
    vec3 sampleEnv(vec2 uv) const;
    vec2 sampleNoise(vec2 uv) const;
    // texture2DArray(uTextures, vec3(uv, layer)): sRGB decode per texel then bilinear (Texture.cpp:146-161)
    vec3 sampleAlbedo(vec2 uv, int layer) const;

private:
    static vec3 fetch3(const std::vector<float>& a, int i, int n) {
        if (i < 0 || i >= n) return vec3(0.0f);
        return vec3(a[3 * i], a[3 * i + 1], a[3 * i + 2]);
    }
};

// Bilinear footprint of GL LINEAR + REPEAT: texel centres at (i + 0.5) / size.
struct Bilerp { int i0, i1; float f; };
inline Bilerp bilerpRepeat(float u, int size) {
    float x = u * (float)size - 0.5f;
    float fl = std::floor(x);
    Bilerp b;
    b.f = x - fl;
    int i = (int)fl;
    int m = i % size; if (m < 0) m += size;
    b.i0 = m;
    b.i1 = (m + 1 == size) ? 0 : m + 1;
    return b;
}

inline Scene::Scene(const ZlSceneDesc& d) {
    numVertices = d.numVertices; numTexcoords = d.numTexcoords; numTriangles = d.numTriangles;
    bvhSize = d.bvhSize; objPrimCount = d.objPrimCount; numMaterials = d.numMaterials;
    numLightTriangles = d.numLightTriangles; numTextures = d.numTextures;
    texMaxW = d.texMaxW; texMaxH = d.texMaxH;
    lightSum = d.lightSum; envSum = d.envSum;
    vertices.assign(d.vertices, d.vertices + 3 * (size_t)numVertices);
    normals.assign(d.normals, d.normals + 3 * (size_t)numVertices);
    if (d.texcoords && numTexcoords > 0) texcoords.assign(d.texcoords, d.texcoords + 2 * (size_t)numTexcoords);
    indices.assign(d.indices, d.indices + 3 * (size_t)numTriangles);
    if (d.bounds) bounds.assign(d.bounds, d.bounds + 6 * (size_t)bvhSize);              // else: built by zo_scene_create (zo_api.cpp)
    if (d.hitTable) hitTable.assign(d.hitTable, d.hitTable + 18 * (size_t)bvhSize);     // else: zo_scene_create threads it itself (zo_api.cpp)
    if (objPrimCount > 0) matTexIndices.assign(d.matTexIndices, d.matTexIndices + objPrimCount);
    materials.assign(d.materials, d.materials + 16 * (size_t)numMaterials);
    if (numLightTriangles > 0) {
        lightPower.assign(d.lightPower, d.lightPower + 3 * (size_t)numLightTriangles);
        lightAlias.assign(d.lightAlias, d.lightAlias + numLightTriangles);
        lightProb.assign(d.lightProb, d.lightProb + numLightTriangles);
    }
    if (numTextures > 0 && d.texels) {
        texels.assign(d.texels, d.texels + (size_t)numTextures * texMaxW * texMaxH * 3);
        texUVScale.assign(d.texUVScale, d.texUVScale + 2 * (size_t)numTextures);
    }
    if (d.envMap && d.envW > 0 && d.envH > 0) {
        envW = d.envW; envH = d.envH;
        envMap.resize(3 * (size_t)envW * envH);
        for (size_t i = 0; i < envMap.size(); i++) envMap[i] = roundToHalf(d.envMap[i]);
        envAlias.assign(d.envAlias, d.envAlias + (size_t)(envW + 1) * envH);
        envAliasProb.assign(d.envAliasProb, d.envAliasProb + (size_t)(envW + 1) * envH);
    } else {  // the reference always binds an env map; a missing one behaves as 1x1 black
        envW = envH = 1;
        envMap.assign(3, 0.0f);
        envAlias.assign(2, 0);
        envAliasProb.assign(2, 1.0f);
        envSum = 0.0f;
    }
    if (d.noise && d.noiseW > 0 && d.noiseH > 0) {
        noiseW = d.noiseW; noiseH = d.noiseH;
        noise.assign(d.noise, d.noise + 2 * (size_t)noiseW * noiseH);
    } else {
        noiseW = noiseH = 1; noise.assign(2, 0.5f);
    }
    sobolMatrices.assign(d.sobolMatrices, d.sobolMatrices + 256 * 32);
    for (int i = 0; i < 256; i++) {  // GL_SRGB decode, evaluated in double then rounded
        double c = i / 255.0;
        double l = (c <= 0.04045) ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4);
        srgbLut[i] = (float)l;
    }
}

inline vec3 Scene::sampleEnv(vec2 uv) const {
    Bilerp bx = bilerpRepeat(uv.x, envW), by = bilerpRepeat(uv.y, envH);
    auto T = [&](int x, int y) {
        const float* p = &envMap[3 * ((size_t)y * envW + x)];
        return vec3(p[0], p[1], p[2]);
    };
    vec3 a = T(bx.i0, by.i0) * (1.0f - bx.f) + T(bx.i1, by.i0) * bx.f;
    vec3 b = T(bx.i0, by.i1) * (1.0f - bx.f) + T(bx.i1, by.i1) * bx.f;
    return a * (1.0f - by.f) + b * by.f;
}

inline vec2 Scene::sampleNoise(vec2 uv) const {
    Bilerp bx = bilerpRepeat(uv.x, noiseW), by = bilerpRepeat(uv.y, noiseH);
    auto T = [&](int x, int y) {
        const float* p = &noise[2 * ((size_t)y * noiseW + x)];
        return vec2(p[0], p[1]);
    };
    vec2 a = T(bx.i0, by.i0) * (1.0f - bx.f) + T(bx.i1, by.i0) * bx.f;
    vec2 b = T(bx.i0, by.i1) * (1.0f - bx.f) + T(bx.i1, by.i1) * bx.f;
    return a * (1.0f - by.f) + b * by.f;
}

inline vec3 Scene::sampleAlbedo(vec2 uv, int layer) const {
    if (layer < 0 || layer >= numTextures || texels.empty()) return vec3(0.0f);
    Bilerp bx = bilerpRepeat(uv.x, texMaxW), by = bilerpRepeat(uv.y, texMaxH);
    auto T = [&](int x, int y) {
        const uint8_t* p = &texels[3 * (((size_t)layer * texMaxH + y) * texMaxW + x)];
        return vec3(srgbLut[p[0]], srgbLut[p[1]], srgbLut[p[2]]);
    };
    vec3 a = T(bx.i0, by.i0) * (1.0f - bx.f) + T(bx.i1, by.i0) * bx.f;
    vec3 b = T(bx.i0, by.i1) * (1.0f - bx.f) + T(bx.i1, by.i1) * bx.f;
    return a * (1.0f - by.f) + b * by.f;
}

}  // namespace zo
