// ORACLE — test infrastructure only.  Part of the recipe that builds oracle/_ref/libzillum_ref.so.
//
// ref_harness.cpp — extern "C" driver of the reference's OWN shader programs (their text,
// rewritten by ref_glsl2cpp.py and compiled against ref_shim/glsl_shim.h).  It plays the part of
// the reference's host glue for one dispatch: it binds the arrays of a ZlSceneDesc to the
// samplers the shaders declare and sets their uniforms BY NAME, exactly the list of
// NaivePathIntegrator::updateUniforms (src/integrator/NaivePath.cpp:12-66,97-100),
// LightPathIntegrator (LightPath.cpp:17-66,99-104) and TriplePathIntegrator
// (TriplePath.cpp:17-77,117-126), then runs main() once per global invocation.
// The API mirrors oracle/zo_api.cpp so that tests can put either behind the same checks.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "glsl_shim.h"
#include "../include/zillum_cuda.h"

namespace glsl {
Program*& programList() { static Program* head = nullptr; return head; }
}
using namespace glsl;

// defined in ref_host.cpp (the reference's own C++: Sampler::sobolSample over its own SobolMatrices table)
extern "C" uint32_t zr_sobol_sample(uint32_t index, int dim);

namespace {

Program* findProgram(const char* name) {
    for (Program* p = programList(); p; p = p->next)
        if (!std::strcmp(p->name, name)) return p;
    std::fprintf(stderr, "[zillum_ref] no program %s\n", name);
    std::abort();
}
const UniformEntry* findUniform(Program* p, const char* name) {
    for (int i = 0; i < p->numUniforms; i++)
        if (!std::strcmp(p->uniforms[i].name, name)) return &p->uniforms[i];
    return nullptr;     // the GL behaviour: glGetUniformLocation = -1, the set is ignored (Shader.cpp:143-158)
}
void set1i(Program* p, const char* n, int v) {
    const UniformEntry* u = findUniform(p, n); if (!u) return;
    switch (u->type) {
    case U_INT: *(int*)u->ptr = v; break;
    case U_UINT: *(uint*)u->ptr = (uint)v; break;
    case U_BOOL: *(bool*)u->ptr = v != 0; break;
    default: std::abort();
    }
}
void set1f(Program* p, const char* n, float v) { const UniformEntry* u = findUniform(p, n); if (u) { if (u->type != U_FLOAT) std::abort(); *(float*)u->ptr = v; } }
void setVec3(Program* p, const char* n, const float* v) { const UniformEntry* u = findUniform(p, n); if (u) { if (u->type != U_VEC3) std::abort(); *(vec3*)u->ptr = vec3(v[0], v[1], v[2]); } }
void setVec2i(Program* p, const char* n, int a, int b) { const UniformEntry* u = findUniform(p, n); if (u) { if (u->type != U_IVEC2) std::abort(); *(ivec2*)u->ptr = ivec2(a, b); } }
void setMat3(Program* p, const char* n, const float* m) {
    const UniformEntry* u = findUniform(p, n); if (!u) return;
    if (u->type != U_MAT3) std::abort();
    *(mat3*)u->ptr = mat3(vec3(m[0], m[1], m[2]), vec3(m[3], m[4], m[5]), vec3(m[6], m[7], m[8]));
}
void setTexture(Program* p, const char* n, const TexBinding& t) {
    const UniformEntry* u = findUniform(p, n); if (!u) return;
    if (u->type < U_SAMPLER_BUFFER || u->type > U_SAMPLER_2D_ARRAY) std::abort();
    *(TexBinding*)u->ptr = t;      // every sampler type is a TexBinding
}
void setImage(Program* p, const char* n, float* data, int w, int h, int comps) {
    const UniformEntry* u = findUniform(p, n); if (!u) return;
    if (u->type != U_IMAGE_2D) std::abort();
    image2D* im = (image2D*)u->ptr; im->data = data; im->w = w; im->h = h; im->comps = comps;
}
TexBinding buf(const void* data, int comps, size_t count, int kind) {
    TexBinding t; t.data = data; t.comps = comps; t.w = (int)count; t.kind = kind; return t;
}
TexBinding tex2d(const void* data, int comps, int w, int h, int kind) {
    TexBinding t; t.data = data; t.comps = comps; t.w = w; t.h = h; t.kind = kind; return t;
}

uint32_t f2u(float f) { return zl_f2u(f); }
// IEEE binary16 round trip: the RGB16F storage of the environment map (EnvironmentMap.cpp:13, TextureFormat::Col3x16f)
float roundToHalf(float f) {
    uint32_t x = f2u(f), sign = x & 0x80000000u, ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return f;
    if (ax >= 0x477ff000u) return zl_u2f(sign | 0x7f800000u);
    if (ax < 0x33000001u) return zl_u2f(sign);
    if (ax < 0x38800000u) {
        float q = zl_u2f(ax) * 16777216.0f;
        float r = std::nearbyint(q);
        return zl_u2f(sign | f2u(r * (1.0f / 16777216.0f)));
    }
    uint32_t rem = ax & 0x1fffu, base = ax & ~0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (base & 0x2000u))) base += 0x2000u;
    return zl_u2f(sign | base);
}

uint32_t sobolGenerator(int flat) {          // content of uSobolSeq: Sampler::genSobolSeqTexture (Sampler.cpp:48-64), texel [i * 256 + j]
    if (flat < 0) return 0u;
    return zr_sobol_sample((uint32_t)(flat / 256), flat % 256);
}

struct RefScene {
    std::vector<float> vertices, normals, texcoords, bounds, materials, lightPower, lightProb, texUVScale, envMap, envAliasProb, noise;
    std::vector<uint32_t> indices;
    std::vector<int32_t> hitTable, matTexIndices, lightAlias, envAlias;
    std::vector<uint8_t> texels;
    float srgbLut[256];
    int numVertices = 0, numTexcoords = 0, numTriangles = 0, bvhSize = 0, objPrimCount = 0, numMaterials = 0, numLightTriangles = 0;
    int numTextures = 0, texMaxW = 0, texMaxH = 0, envW = 1, envH = 1, noiseW = 1, noiseH = 1;
    float lightSum = 0.0f, envSum = 0.0f;
};

}  // namespace

extern "C" int zr_build_bvh(const float* vertices, int numVertices, const uint32_t* indices, int numTriangles,
                            float* boundsOut, int32_t* hitTableOut);

extern "C" {

void* zr_scene_create(const ZlSceneDesc* d) {
    RefScene* s = new RefScene;
    s->numVertices = d->numVertices; s->numTexcoords = d->numTexcoords; s->numTriangles = d->numTriangles;
    s->bvhSize = d->bvhSize; s->objPrimCount = d->objPrimCount; s->numMaterials = d->numMaterials;
    s->numLightTriangles = d->numLightTriangles; s->numTextures = d->numTextures; s->texMaxW = d->texMaxW; s->texMaxH = d->texMaxH;
    s->lightSum = d->lightSum; s->envSum = d->envSum;
    s->vertices.assign(d->vertices, d->vertices + 3 * (size_t)s->numVertices);
    s->normals.assign(d->normals, d->normals + 3 * (size_t)s->numVertices);
    if (d->texcoords && s->numTexcoords > 0) s->texcoords.assign(d->texcoords, d->texcoords + 2 * (size_t)s->numTexcoords);
    s->indices.assign(d->indices, d->indices + 3 * (size_t)s->numTriangles);
    if (d->bounds && d->hitTable) {
        s->bounds.assign(d->bounds, d->bounds + 6 * (size_t)s->bvhSize);
        s->hitTable.assign(d->hitTable, d->hitTable + 18 * (size_t)s->bvhSize);
    } else {                                         // no host tree in the descriptor: the reference's own BVH::build makes one
        s->bvhSize = 2 * s->numTriangles - 1;
        s->bounds.resize(6 * (size_t)s->bvhSize); s->hitTable.resize(18 * (size_t)s->bvhSize);
        zr_build_bvh(d->vertices, s->numVertices, d->indices, s->numTriangles, s->bounds.data(), s->hitTable.data());
    }
    if (s->objPrimCount > 0) s->matTexIndices.assign(d->matTexIndices, d->matTexIndices + s->objPrimCount);
    s->materials.assign(d->materials, d->materials + 16 * (size_t)s->numMaterials);
    if (s->numLightTriangles > 0) {
        s->lightPower.assign(d->lightPower, d->lightPower + 3 * (size_t)s->numLightTriangles);
        s->lightAlias.assign(d->lightAlias, d->lightAlias + s->numLightTriangles);
        s->lightProb.assign(d->lightProb, d->lightProb + s->numLightTriangles);
    }
    if (s->numTextures > 0 && d->texels) {
        s->texels.assign(d->texels, d->texels + (size_t)s->numTextures * s->texMaxW * s->texMaxH * 3);
        s->texUVScale.assign(d->texUVScale, d->texUVScale + 2 * (size_t)s->numTextures);
    }
    if (d->envMap && d->envW > 0 && d->envH > 0) {
        s->envW = d->envW; s->envH = d->envH;
        s->envMap.resize(3 * (size_t)s->envW * s->envH);
        for (size_t i = 0; i < s->envMap.size(); i++) s->envMap[i] = roundToHalf(d->envMap[i]);   // the RGB16F upload
        s->envAlias.assign(d->envAlias, d->envAlias + (size_t)(s->envW + 1) * s->envH);
        s->envAliasProb.assign(d->envAliasProb, d->envAliasProb + (size_t)(s->envW + 1) * s->envH);
    } else {                                         // the reference always binds an env map; none = 1x1 black
        s->envMap.assign(3, 0.0f); s->envAlias.assign(2, 0); s->envAliasProb.assign(2, 1.0f); s->envSum = 0.0f;
    }
    if (d->noise && d->noiseW > 0 && d->noiseH > 0) {
        s->noiseW = d->noiseW; s->noiseH = d->noiseH;
        s->noise.assign(d->noise, d->noise + 2 * (size_t)s->noiseW * s->noiseH);
    } else s->noise.assign(2, 0.5f);
    for (int i = 0; i < 256; i++) {                  // GL_SRGB texel decode (GL 4.5 spec §8.24), evaluated in binary64
        double c = i / 255.0;
        s->srgbLut[i] = (float)((c <= 0.04045) ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
    }
    return s;
}
void zr_scene_destroy(void* s) { delete (RefScene*)s; }

int zr_get_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void zr_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

}  // extern "C"

namespace {

// The uniform list shared by the three integrators (NaivePath.cpp:20-66, LightPath.cpp:25-66, TriplePath.cpp:27-77)
void bindScene(Program* p, const RefScene& s, const ZlRenderParams& U) {
    setTexture(p, "uVertices", buf(s.vertices.data(), 3, s.numVertices, 0));
    setTexture(p, "uNormals", buf(s.normals.data(), 3, s.numVertices, 0));
    setTexture(p, "uTexCoords", buf(s.texcoords.data(), 2, s.numTexcoords, 0));
    setTexture(p, "uIndices", buf(s.indices.data(), 1, s.indices.size(), 1));
    setTexture(p, "uBounds", buf(s.bounds.data(), 3, 2 * (size_t)s.bvhSize, 0));
    setTexture(p, "uHitTable", buf(s.hitTable.data(), 3, 6 * (size_t)s.bvhSize, 1));
    setTexture(p, "uMatTexIndices", buf(s.matTexIndices.data(), 1, s.matTexIndices.size(), 1));
    setTexture(p, "uEnvMap", tex2d(s.envMap.data(), 3, s.envW, s.envH, 0));
    setTexture(p, "uEnvAliasTable", tex2d(s.envAlias.data(), 1, s.envW + 1, s.envH, 1));
    setTexture(p, "uEnvAliasProb", tex2d(s.envAliasProb.data(), 1, s.envW + 1, s.envH, 0));
    setTexture(p, "uMaterials", buf(s.materials.data(), 4, 4 * (size_t)s.numMaterials, 0));
    setTexture(p, "uMatTypes", buf(s.materials.data(), 4, 4 * (size_t)s.numMaterials, 1));   // same buffer, integer view (NaivePath.cpp:30-31)
    setTexture(p, "uLightPower", buf(s.lightPower.data(), 3, s.numLightTriangles, 0));
    setTexture(p, "uLightAlias", buf(s.lightAlias.data(), 1, s.numLightTriangles, 1));
    setTexture(p, "uLightProb", buf(s.lightProb.data(), 1, s.numLightTriangles, 0));
    TexBinding arr; arr.data = s.texels.empty() ? nullptr : s.texels.data(); arr.comps = 3; arr.w = s.texMaxW; arr.h = s.texMaxH;
    arr.layers = s.numTextures; arr.kind = 2; arr.lut = s.srgbLut;
    setTexture(p, "uTextures", arr);
    setTexture(p, "uTexUVScale", buf(s.texUVScale.data(), 2, s.numTextures, 0));
    TexBinding sob; sob.comps = 1; sob.w = 131072 * 256; sob.kind = 1; sob.generator = sobolGenerator;
    setTexture(p, "uSobolSeq", sob);
    setTexture(p, "uNoiseTex", tex2d(s.noise.data(), 2, s.noiseW, s.noiseH, 0));
    set1i(p, "uNumLightTriangles", s.numLightTriangles);
    set1f(p, "uLightSum", s.lightSum);
    set1f(p, "uEnvSum", s.envSum);
    set1i(p, "uObjPrimCount", s.objPrimCount);
    set1i(p, "uBvhSize", s.bvhSize);
    set1i(p, "uSampleDim", 256);                      // Scene.h:64-65
    set1i(p, "uSampleNum", 131072);
    set1f(p, "uEnvRotation", U.envRotation);
    set1i(p, "uSampler", U.sampler);
    setVec3(p, "uCamF", U.camera.F);
    setVec3(p, "uCamR", U.camera.R);
    setVec3(p, "uCamU", U.camera.U);
    setMat3(p, "uCamMatInv", U.camera.matInv);
    setVec3(p, "uCamPos", U.camera.pos);
    set1f(p, "uTanFOV", U.camera.tanFOV);
    set1f(p, "uCamAsp", U.camera.asp);
    set1f(p, "uLensRadius", U.camera.lensRadius);
    set1f(p, "uFocalDist", U.camera.focalDist);
    setVec2i(p, "uFilmSize", U.filmW, U.filmH);
    set1i(p, "uRussianRoulette", U.russianRoulette);
    set1i(p, "uMaxDepth", U.maxDepth);
    set1i(p, "uSampleLight", U.sampleLight);
    set1i(p, "uLightEnvUniformSample", U.lightEnvUniformSample);
    set1f(p, "uLightSamplePortion", U.lightPortion);
    set1i(p, "uSpp", U.spp);
    set1i(p, "uFreeCounter", U.freeCounter);
}

// caller's film is W*H*4 (rgb sums); the light / triple programs address an r32f image of 3W x H
// (LightPath.cpp:9-14).  Present the film in that layout, run, and copy back: the additions then
// happen on the same running sums, in the order the invocations run.
struct R32fView {
    std::vector<float> px; float* film; int W, H;
    R32fView(float* f, int w, int h) : px((size_t)3 * w * h), film(f), W(w), H(h) {
        for (size_t i = 0; i < (size_t)w * h; i++) for (int c = 0; c < 3; c++) px[3 * i + c] = f[4 * i + c];
    }
    ~R32fView() { for (size_t i = 0; i < (size_t)W * H; i++) for (int c = 0; c < 3; c++) film[4 * i + c] = px[3 * i + c]; }
};

}  // namespace

extern "C" {

// path_integ_naive.glsl over rows rowBegin, rowBegin + rowStride, ... < rowEnd (< 0: all) — NaivePath.cpp:94-100
int zr_path_pass(void* scene, const ZlRenderParams* U, float* film, int rowBegin, int rowEnd, int rowStride) {
    Program* p = findProgram("path_integ_naive.glsl");
    bindScene(p, *(RefScene*)scene, *U);
    setImage(p, "uFrame", film, U->filmW, U->filmH, 4);
    if (rowEnd < 0) rowEnd = U->filmH;
    if (rowStride < 1) rowStride = 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = rowBegin; y < rowEnd; y += rowStride)
        for (int x = 0; x < U->filmW; x++) p->invoke((uint)x, (uint)y, 0u);
    return 0;
}
int zr_triple_pt_pass(void* scene, const ZlRenderParams* U, float* film, int rowBegin, int rowEnd, int rowStride) {
    Program* p = findProgram("triple_path_pass_pt.glsl");
    bindScene(p, *(RefScene*)scene, *U);
    R32fView v(film, U->filmW, U->filmH);
    setImage(p, "uFrame", v.px.data(), 3 * U->filmW, U->filmH, 1);
    if (rowEnd < 0) rowEnd = U->filmH;
    if (rowStride < 1) rowStride = 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = rowBegin; y < rowEnd; y += rowStride)
        for (int x = 0; x < U->filmW; x++) p->invoke((uint)x, (uint)y, 0u);
    return 0;
}
// light_path_integ.glsl: global ids idBegin..idEnd of the 1536 * blocksOnePass invocations (LightPath.cpp:99-104)
int zr_light_pass(void* scene, const ZlRenderParams* U, float* film, long idBegin, long idEnd) {
    Program* p = findProgram("light_path_integ.glsl");
    bindScene(p, *(RefScene*)scene, *U);
    set1i(p, "uSampler", 0);                          // LightPath.cpp:48-49: the light tracer always uses the independent sampler
    set1i(p, "uBlocksOnePass", U->blocksOnePass);
    R32fView v(film, U->filmW, U->filmH);
    setImage(p, "uFrame", v.px.data(), 3 * U->filmW, U->filmH, 1);
    long total = (long)p->localSize[0] * U->blocksOnePass;
    if (idEnd < 0) idEnd = total;
#pragma omp parallel for schedule(dynamic, 256)
    for (long id = idBegin; id < idEnd; id++) p->invoke((uint)id, 0u, 0u);
    return 0;
}
int zr_triple_lpt_pass(void* scene, const ZlRenderParams* U, float* film, long idBegin, long idEnd) {
    Program* p = findProgram("triple_path_pass_lpt.glsl");
    bindScene(p, *(RefScene*)scene, *U);
    set1i(p, "uSampler", 0);                          // TriplePath.cpp:72
    set1i(p, "uBlocksOnePass", U->blocksOnePass);
    set1i(p, "uLoopsPerPass", U->loopsPerPass);
    set1f(p, "uScale", U->scale);
    R32fView v(film, U->filmW, U->filmH);
    setImage(p, "uFrame", v.px.data(), 3 * U->filmW, U->filmH, 1);
    long total = (long)p->localSize[0] * U->blocksOnePass;
    if (idEnd < 0) idEnd = total;
#pragma omp parallel for schedule(dynamic, 256)
    for (long id = idBegin; id < idEnd; id++) p->invoke((uint)id, 0u, 0u);
    return 0;
}

int zr_trace_rays(void* scene, const float* rays, size_t n, int anyhit, const float* tMax, int32_t* outIds, float* outT, int32_t* outSteps) {
    Program* p = findProgram("path_integ_naive.glsl");
    ZlRenderParams dummy{};
    bindScene(p, *(RefScene*)scene, dummy);
    return p->trace(rays, n, anyhit, tMax, outIds, outT, outSteps);
}
int zr_debug_eval(void* scene, const ZlRenderParams* U, int op, const float* in, int inStride, float* out, int outStride, size_t n) {
    Program* p = findProgram("path_integ_naive.glsl");
    bindScene(p, *(RefScene*)scene, *U);
    return p->kat(op, in, inStride, out, outStride, n);
}

// post_proc.glsl over a W*H*4 film: uIn = the rgba32f frame, uOut = the rgba32f result (Application.cpp:644-663)
int zr_post_proc(const float* film, int w, int h, float scale, int toneMapper, float* outRgba) {
    Program* p = findProgram("post_proc.glsl");
    setTexture(p, "uIn", tex2d(film, 4, w, h, 0));
    setImage(p, "uOut", outRgba, w, h, 4);
    set1i(p, "uToneMapper", toneMapper);
    set1f(p, "uResultScale", scale);
    setVec2i(p, "uFilmSize", w, h);
    set1i(p, "uPreview", 0);
    set1i(p, "uPreviewScale", 1);
#pragma omp parallel for
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) p->invoke((uint)x, (uint)y, 0u);
    return 0;
}

}  // extern "C"
