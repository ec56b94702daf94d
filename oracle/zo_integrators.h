// ORACLE — test infrastructure only (see zo_vec.h header).  Pinned to oracle/_ref by tests/test_ref_parity.py.
// zo_integrators.h — restatement of the four integrator kernels:
//   src/shader/path_integ_naive.glsl, light_path_integ.glsl,
//   triple_path_pass_pt.glsl, triple_path_pass_lpt.glsl
// and of the per-pass dispatch of src/integrator/{NaivePath,LightPath,TriplePath}.cpp.
// Film layout: W*H float4, row 0 = bottom of the image (App. A `uFrame`); rgb are running
// sums, w is unused.  The 3WxH r32f film + img_copy pass of the reference collapse into it.
#pragma once
#include "zo_shade.h"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace zo {

struct Stats { uint64_t rays = 0, nodeVisits = 0, triTests = 0, paths = 0, splats = 0; };

struct Film {
    float* px; int W, H;
    void add(int x, int y, vec3 v) {            // owner-writes accumulate (path / triple-PT)
        float* p = px + 4 * ((size_t)y * W + x);
        p[0] += v.x; p[1] += v.y; p[2] += v.z;
    }
    void atomicAdd(int x, int y, vec3 v) {      // imageAtomicAdd(r32f) x3 (light / LPT)
        float* p = px + 4 * ((size_t)y * W + x);
#pragma omp atomic
        p[0] += v.x;
#pragma omp atomic
        p[1] += v.y;
#pragma omp atomic
        p[2] += v.z;
    }
};

// ------------------------------------------------------------------------------------------
// path_integ_naive.glsl:35-143
// ------------------------------------------------------------------------------------------
inline vec3 pathIntegTrace(Shader& sh, Ray ray, int& s) {
    const Scene& S = sh.S; const ZlRenderParams& U = sh.U;
    float primDist;
    int id = sh.bvhHit(ray, primDist);
    vec3 pos = rayPoint(ray, primDist);
    if (id == -1) return sh.envLe(ray.dir);
    else if (id - S.objPrimCount >= 0) return sh.lightLe(id - S.objPrimCount, pos, -ray.dir);

    vec3 wo = -ray.dir;
    vec3 result = vec3(0.0f);
    vec3 throughput = vec3(1.0f);

    for (int bounce = 1; bounce <= U.maxDepth; bounce++) {
        SurfaceInfo surf = sh.triangleSurfaceInfo(id, pos);
        int matTexId = S.matTexIndices[id];
        int matId = matTexId & 0x0000ffff;
        int texId = matTexId >> 16;
        uint32_t matType = sh.loadMaterialType(matId);
        if (matType != Dielectric && matType != ThinDielectric) {
            if (dot(surf.ns, wo) < 0) { surf.ns = -surf.ns; surf.ng = -surf.ng; }
        }
        BSDFParam matParam = sh.loadMaterial(matType, matId, texId, surf.uv);

        if (U.sampleLight) {
            float ud = sh.sample1D(s);
            vec4 us = sh.sample4D(s);
            LightLiSample samp = sh.sampleLightAndEnv(pos, ud, us);
            if (samp.pdf > 0.0f) {
                vec4 bsdfAndPdf = sh.materialBSDFAndPdf(matType, matParam, wo, samp.wi, surf.ns, Radiance);
                float weight = biHeuristic(samp.pdf, bsdfAndPdf.w);
                result += bsdfAndPdf.xyz() * throughput * satDot(surf.ns, samp.wi) * samp.coef * weight;
            }
        }

        BSDFSample samp = sh.materialSample(matType, matParam, surf.ns, wo, Radiance, sh.sample3D(s));
        vec3 wi = samp.wi;
        float bsdfPdf = samp.pdf;
        vec3 bsdf = samp.bsdf;
        uint32_t flag = samp.flag;
        bool deltaBsdf = (flag == SpecRefl || flag == SpecTrans);
        if (bsdfPdf < 1e-8f) break;
        throughput *= bsdf / bsdfPdf * (deltaBsdf ? 1.0f : absDot(surf.ns, wi));

        ray = rayOffseted(pos, wi);
        float dist;
        int nextId = sh.bvhHit(ray, dist);
        int lightId = nextId - S.objPrimCount;
        vec3 nextPos = rayPoint(ray, dist);

        if (nextId == -1) {
            vec3 radiance = sh.envLe(wi);
            float weight = 1.0f;
            if (U.sampleLight && !deltaBsdf) {
                float envPdf = sh.envPdfLi(wi) * sh.pdfSelectEnv();
                weight = (envPdf <= 0.0f) ? 0.0f : biHeuristic(bsdfPdf, envPdf);
            }
            result += radiance * throughput * weight;
            break;
        } else if (lightId >= 0) {
            vec3 radiance = sh.lightLe(lightId, nextPos, -wi);
            float weight = 1.0f;
            if (U.sampleLight && !deltaBsdf) {
                float lightPdf = sh.lightPdfLi(lightId, pos, nextPos) * sh.pdfSelectLight(lightId);
                weight = (lightPdf <= 0.0f) ? 0.0f : biHeuristic(bsdfPdf, lightPdf);
            }
            result += radiance * throughput * weight;
            break;
        }
        if (U.russianRoulette) {
            float continueProb = gmin(maxComponent(bsdf / bsdfPdf), 0.95f);
            if (sh.sample1D(s) >= continueProb) break;
            throughput /= continueProb;
        }
        id = nextId;
        pos = nextPos;
        wo = -wi;
    }
    return result;
}

// Seeds shared by the two per-pixel kernels (path_integ_naive.glsl:156-164,
// triple_path_pass_pt.glsl:208-216).
inline void seedPixel(Shader& sh, vec2 scrCoord) {
    const ZlRenderParams& U = sh.U;
    vec2 texSize = vec2((float)U.filmW, (float)U.filmH);
    vec2 noiseCoord = sh.S.sampleNoise(scrCoord);
    noiseCoord = sh.S.sampleNoise(noiseCoord);
    vec2 texCoord = texSize * noiseCoord;
    sh.randSeed = ((uint32_t)texCoord.x * (uint32_t)U.freeCounter) + (uint32_t)texCoord.y;
    sh.sampleSeed = (uint32_t)texCoord.x * (uint32_t)texCoord.y;
    const int uSampleDim = 256, uSampleNum = 131072;
    sh.sampleOffset = U.spp * uSampleDim;
    if (sh.sampleOffset > uSampleNum * uSampleDim) sh.sampleOffset -= uSampleNum * uSampleDim;
}

inline void accumulateStats(Stats* st, const Shader& sh, uint64_t paths) {
    if (!st) return;
#pragma omp atomic
    st->rays += sh.rays;
#pragma omp atomic
    st->nodeVisits += sh.nodeVisits;
#pragma omp atomic
    st->triTests += sh.triTests;
#pragma omp atomic
    st->paths += paths;
}

// path_integ_naive.glsl:145-174 over the whole film (NaivePath.cpp:94-100).
// rowBegin/rowEnd restrict the pass to a band of rows (bounded CPU-baseline samples).
inline void pathPass(const Scene& S, const ZlRenderParams& U, float* filmPx, Stats* st,
                     int rowBegin = 0, int rowEnd = -1, int rowStride = 1) {
    Film film{filmPx, U.filmW, U.filmH};
    if (rowEnd < 0) rowEnd = U.filmH;
    if (rowStride < 1) rowStride = 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = rowBegin; y < rowEnd; y += rowStride) {
        Shader acc(S, U, U.sampler);
        for (int x = 0; x < U.filmW; x++) {
            Shader sh(S, U, U.sampler);
            vec2 scrCoord = vec2((float)x, (float)y) / vec2((float)U.filmW, (float)U.filmH);
            int sampleIdx = 0;
            seedPixel(sh, scrCoord);
            Ray ray = sh.thinLensCameraSampleRay(scrCoord, sh.sample4D(sampleIdx));
            vec3 result = pathIntegTrace(sh, ray, sampleIdx);
            if (!hasNan(result)) film.add(x, y, result);
            acc.rays += sh.rays; acc.nodeVisits += sh.nodeVisits; acc.triTests += sh.triTests;
        }
        accumulateStats(st, acc, (uint64_t)U.filmW);
    }
}

// ------------------------------------------------------------------------------------------
// light_path_integ.glsl
// ------------------------------------------------------------------------------------------
inline void accumulateFilmSplat(Shader& sh, Film& film, vec2 uv, vec3 res, float scale, Stats* st) {  // :34-43
    if (!sh.inFilmBound(uv)) return;
    int ix = (int)(uv.x * (float)film.W), iy = (int)(uv.y * (float)film.H);
    if (ix < 0 || iy < 0 || ix >= film.W || iy >= film.H) return;   // App. B #14: GL drops the OOB write
    film.atomicAdd(ix, iy, res * scale);
    if (st) {
#pragma omp atomic
        st->splats += 1;
    }
}

inline void lightIntegTrace(Shader& sh, int& s, Film& film, Stats* st) {                  // :45-146
    const Scene& S = sh.S; const ZlRenderParams& U = sh.U;
    int light = sh.lightSampleOne(sh.sample2D(s));
    float pdfSource = sh.lightPdfSampleOne(light);
    Ray ray; vec3 wo; vec3 throughput;
    {
        int triId = light + S.objPrimCount;
        vec3 pLit = sh.triangleSampleUniform(triId, sh.sample2D(s));
        CameraIiSample ciSamp = sh.thinLensCameraSampleIi(pLit, sh.sample2D(s));
        if (ciSamp.pdf > 0) {
            vec3 pCam = pLit + ciSamp.wi * ciSamp.dist;
            float pdfPos = 1.0f / sh.triangleArea(triId);
            if (sh.visible(pLit, pCam)) {
                vec3 Le = sh.lightLe(light, pLit, ciSamp.wi);
                vec3 contrib = Le * ciSamp.Ii / (ciSamp.pdf * pdfPos * pdfSource);
                if (!isBlack(contrib)) accumulateFilmSplat(sh, film, ciSamp.uv, contrib, 1.0f, st);
            }
        }
        LightLeSample leSamp = sh.lightSampleOneLe(light, sh.sample4D(s));
        vec3 nl = sh.triangleSurfaceInfo(triId, leSamp.ray.ori).ng;
        wo = -leSamp.ray.dir;
        ray = rayOffseted(leSamp.ray);
        throughput = leSamp.Le * absDot(nl, -wo) / (pdfSource * leSamp.pdfPos * leSamp.pdfDir);
    }
    for (int bounce = 1; bounce <= U.maxDepth; bounce++) {
        float dist;
        int id = sh.bvhHit(ray, dist);
        if (id == -1) break;
        if (id - S.objPrimCount >= 0) break;
        vec3 pos = rayPoint(ray, dist);
        SurfaceInfo surf = sh.triangleSurfaceInfo(id, pos);
        int matTexId = S.matTexIndices[id];
        int matId = matTexId & 0x0000ffff;
        int texId = matTexId >> 16;
        uint32_t matType = sh.loadMaterialType(matId);
        if (matType != Dielectric && matType != ThinDielectric) {
            if (dot(surf.ns, wo) < 0) { surf.ns = -surf.ns; surf.ng = -surf.ng; }
        }
        BSDFParam matParam = sh.loadMaterial(matType, matId, texId, surf.uv);
        {
            CameraIiSample ciSamp = sh.thinLensCameraSampleIi(pos, sh.sample2D(s));
            if (ciSamp.pdf > 0) {
                vec3 pCam = pos + ciSamp.wi * ciSamp.dist;
                if (sh.visible(pos, pCam)) {
                    vec3 bsdf = sh.materialBSDF(matType, matParam, wo, ciSamp.wi, surf.ns, Importance);
                    float cosWi = satDot(surf.ng, ciSamp.wi) * std::fabs(dot(surf.ns, wo) / dot(surf.ng, wo));
                    vec3 res = ciSamp.Ii * bsdf * throughput * cosWi / ciSamp.pdf;
                    if (!hasNan(res) && !std::isnan(ciSamp.pdf) && ciSamp.pdf > 1e-8f && !isBlack(res))
                        accumulateFilmSplat(sh, film, ciSamp.uv, res, 1.0f, st);
                }
            }
        }
        BSDFSample samp = sh.materialSample(matType, matParam, surf.ns, wo, Importance, sh.sample3D(s));
        vec3 wi = samp.wi;
        float bsdfPdf = samp.pdf;
        vec3 bsdf = samp.bsdf;
        uint32_t flag = samp.flag;
        bool deltaBsdf = (flag == SpecRefl || flag == SpecTrans);
        if (bsdfPdf < 1e-8f || std::isnan(bsdfPdf)) break;
        if (U.russianRoulette) {
            float continueProb = gmin(maxComponent(bsdf / bsdfPdf), 1.0f);
            if (sh.sample1D(s) >= continueProb) break;
            throughput /= continueProb;
        }
        float cosWi = deltaBsdf ? 1.0f : std::fabs(dot(surf.ng, wi) * dot(surf.ns, wo) / dot(surf.ng, wo));
        throughput *= bsdf * cosWi / bsdfPdf;
        ray = rayOffseted(pos, wi);
        wo = -wi;
    }
}

// light_path_integ.glsl:148-159 over blocksOnePass*1536 invocations (LightPath.cpp:104).
// idBegin/idEnd restrict the pass to a sub-range of invocation ids.
inline void lightPass(const Scene& S, const ZlRenderParams& U, float* filmPx, Stats* st,
                      long idBegin = 0, long idEnd = -1) {
    Film film{filmPx, U.filmW, U.filmH};
    long total = (long)ZL_LIGHT_GROUP_SIZE * U.blocksOnePass;
    if (idEnd < 0) idEnd = total;
    if (S.numLightTriangles <= 0) return;
#pragma omp parallel for schedule(dynamic, 256)
    for (long id = idBegin; id < idEnd; id++) {
        Shader sh(S, U, 0);     // uSampler forced to 0 (LightPath.cpp:48-49)
        int sampleIdx = 0;
        sh.setRngSeed((uint32_t)U.spp * ((uint32_t)ZL_LIGHT_GROUP_SIZE * (uint32_t)U.blocksOnePass) + (uint32_t)id + (uint32_t)U.freeCounter);
        lightIntegTrace(sh, sampleIdx, film, st);
        accumulateStats(st, sh, 1);
    }
}

// ------------------------------------------------------------------------------------------
// triple_path_pass_pt.glsl
// ------------------------------------------------------------------------------------------
inline float remap(float p) { return p < 1e-8f ? 1.0f : p * p; }                          // :44-47
inline float weightS0(float s1s0, float t1s0) { return 1.0f / (1.0f + s1s0 + t1s0); }     // :49-52
inline float weightS1(float s1s0, float t1s1) { return s1s0 / (1.0f + s1s0 + s1s0 * t1s1); }  // :54-57

inline vec3 traceCameraPath(Shader& sh, Ray ray, int& s) {                                // :59-195
    const Scene& S = sh.S; const ZlRenderParams& U = sh.U;
    float primDist;
    int id = sh.bvhHit(ray, primDist);
    vec3 pos = rayPoint(ray, primDist);
    if (id == -1) return sh.envLe(ray.dir);
    else if (id - S.objPrimCount >= 0) return sh.lightLe(id - S.objPrimCount, pos, -ray.dir);

    SurfaceInfo surf = sh.triangleSurfaceInfo(id, pos);
    vec3 wo = -ray.dir;
    vec3 prevNorm = sh.uCamF();
    vec3 result = vec3(0.0f);
    vec3 throughput = vec3(1.0f);

    CameraPdf camPdf = sh.thinLensCameraPdfIe(ray);
    float primaryPdf = remap(camPdf.pdfPos) / remap(camPdf.pdfDir * absDot(surf.ns, ray.dir) / square(primDist));
    float t1s0 = primaryPdf;
    float t1s1 = primaryPdf;

    for (int bounce = 1; bounce <= U.maxDepth; bounce++) {
        if (bounce > 1) surf = sh.triangleSurfaceInfo(id, pos);
        int matTexId = S.matTexIndices[id];
        int matId = matTexId & 0x0000ffff;
        int texId = matTexId >> 16;
        uint32_t matType = sh.loadMaterialType(matId);
        if (matType != Dielectric && matType != ThinDielectric) {
            if (dot(surf.ns, wo) < 0) { surf.ns = -surf.ns; surf.ng = -surf.ng; }
        }
        BSDFParam matParam = sh.loadMaterial(matType, matId, texId, surf.uv);
        {
            int light = sh.lightSampleOne(sh.sample2D(s));
            int triId = light + S.objPrimCount;
            float pdfSource = sh.lightPdfSampleOne(light);
            vec3 pLit = sh.triangleSampleUniform(triId, sh.sample2D(s));
            vec3 wi = normalize(pLit - pos);
            vec3 Le = sh.lightLe(light, pLit, -wi);
            if (!isBlack(Le) && sh.visible(pos, pLit)) {
                vec3 nLit = sh.triangleSurfaceInfo(triId, pLit).ng;
                float pA = pdfSource / sh.triangleArea(triId);
                float dist2 = distSquare(pos, pLit);
                float pS = pA * dist2 / absDot(nLit, wi);
                vec4 bsdfAndPdf = sh.materialBSDFAndPdf(matType, matParam, wo, wi, surf.ns, Radiance);
                float pdfRev = sh.materialPdf(matType, matParam, wi, wo, surf.ns, Importance);
                float pdfPLit = remap(pA);
                float coefToSurf = remap(0.5f * PiInv * absDot(surf.ns, wi));
                float coefToLight = remap(bsdfAndPdf.w * satDot(nLit, -wi));
                float coefToPrev = (bounce == 1) ? 1.0f : remap(pdfRev * absDot(prevNorm, wo));
                float coefDist = remap(dist2);
                float weight = weightS1(pdfPLit * coefDist / coefToLight, t1s1 * coefToSurf * coefToPrev / coefDist);
                result += Le * bsdfAndPdf.xyz() * throughput * absDot(surf.ns, wi) / pS * weight;
            }
        }
        BSDFSample samp = sh.materialSample(matType, matParam, surf.ns, wo, Radiance, sh.sample3D(s));
        vec3 wi = samp.wi;
        float bsdfPdf = samp.pdf;
        vec3 bsdf = samp.bsdf;
        uint32_t flag = samp.flag;
        bool deltaBsdf = (flag == SpecRefl || flag == SpecTrans);
        if (bsdfPdf < 1e-8f) break;
        throughput *= bsdf / bsdfPdf * (deltaBsdf ? 1.0f : absDot(surf.ns, wi));

        Ray nextRay = rayOffseted(pos, wi);
        float dist;
        int nextId = sh.bvhHit(nextRay, dist);
        int lightId = nextId - S.objPrimCount;
        vec3 nextPos = rayPoint(nextRay, dist);
        float pdfDirToNext = sh.materialPdf(matType, matParam, wo, wi, surf.ns, Radiance);
        float pdfDirToPrev = sh.materialPdf(matType, matParam, wi, wo, surf.ns, Importance);

        if (nextId == -1) break;
        else if (lightId >= 0) {
            vec3 nLit = sh.triangleSurfaceInfo(nextId, nextPos).ng;
            LightPdf pdfLit = sh.lightPdfLe(lightId, makeRay(nextPos, -wi));
            float pdfPLit = remap(pdfLit.pdfPos * sh.lightPdfSampleOne(lightId));
            float coefToLight = remap(pdfDirToNext * satDot(nLit, -wi));
            float coefToSurf = remap(pdfLit.pdfDir * absDot(surf.ns, wi));
            float coefToPrev = (bounce == 1) ? 1.0f : remap(pdfDirToPrev * absDot(prevNorm, wo));
            float coefDist = remap(dist * dist);
            float weight = std::isnan(t1s0) ? 0.0f : weightS0(pdfPLit * coefDist / coefToLight,
                                                             t1s0 * coefToSurf * pdfPLit * coefToPrev / coefToLight);
            result += sh.lightLe(lightId, nextPos, -wi) * throughput * weight;
            break;
        }
        if (U.russianRoulette) {
            float continueProb = gmin(maxComponent(bsdf / bsdfPdf), 0.95f);
            if (sh.sample1D(s) >= continueProb) break;
            throughput /= continueProb;
        }
        float coef = ((bounce == 1) ? 1.0f : remap(pdfDirToPrev * absDot(prevNorm, wo))) /
                     remap(pdfDirToNext * absDot(sh.triangleNormalShad(nextId, nextPos), wi));
        t1s0 *= coef;
        t1s1 *= coef;
        prevNorm = surf.ns;
        pos = nextPos;
        wo = -wi;
        id = nextId;
    }
    return result;
}

// triple_path_pass_pt.glsl:197-224 (TriplePath.cpp:117-121)
inline void triplePtPass(const Scene& S, const ZlRenderParams& U, float* filmPx, Stats* st,
                         int rowBegin = 0, int rowEnd = -1, int rowStride = 1) {
    Film film{filmPx, U.filmW, U.filmH};
    if (rowEnd < 0) rowEnd = U.filmH;
    if (rowStride < 1) rowStride = 1;
    if (S.numLightTriangles <= 0) return;   // the kernel samples area lights unconditionally
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = rowBegin; y < rowEnd; y += rowStride) {
        Shader acc(S, U, U.sampler);
        for (int x = 0; x < U.filmW; x++) {
            Shader sh(S, U, U.sampler);
            vec2 scrCoord = vec2((float)x, (float)y) / vec2((float)U.filmW, (float)U.filmH);
            int sampleIdx = 0;
            seedPixel(sh, scrCoord);
            Ray ray = sh.thinLensCameraSampleRay(scrCoord, sh.sample4D(sampleIdx));
            vec3 result = traceCameraPath(sh, ray, sampleIdx);
            if (!hasNan(result)) film.add(x, y, result);
            acc.rays += sh.rays; acc.nodeVisits += sh.nodeVisits; acc.triTests += sh.triTests;
        }
        accumulateStats(st, acc, (uint64_t)U.filmW);
    }
}

// ------------------------------------------------------------------------------------------
// triple_path_pass_lpt.glsl
// ------------------------------------------------------------------------------------------
inline float weightT1(float s0t1, float s1t1) { return 1.0f / (s0t1 + s1t1 + 1.0f); }     // :53-56

inline void traceLightPath(Shader& sh, int& s, Film& film, Stats* st) {                   // :58-182
    const Scene& S = sh.S; const ZlRenderParams& U = sh.U;
    int light = sh.lightSampleOne(sh.sample2D(s));
    float pdfSource = sh.lightPdfSampleOne(light);
    Ray ray; vec3 wo; vec3 throughput;
    vec3 prevNorm; float prevPdfDir;
    float s0t1, s1t1;
    {
        int triId = light + S.objPrimCount;
        vec3 pLit = sh.triangleSampleUniform(triId, sh.sample2D(s));   // drawn and unused (App. B #19)
        (void)pLit;
        LightLeSample leSamp = sh.lightSampleOneLe(light, sh.sample4D(s));
        vec3 nl = sh.triangleSurfaceInfo(triId, leSamp.ray.ori).ng;
        wo = -leSamp.ray.dir;
        ray = rayOffseted(leSamp.ray);
        throughput = leSamp.Le * absDot(nl, -wo) / (pdfSource * leSamp.pdfPos * leSamp.pdfDir);
        prevNorm = nl;
        prevPdfDir = leSamp.pdfDir;
        s0t1 = 1.0f / remap(leSamp.pdfPos * pdfSource);
        s1t1 = 1.0f;
    }
    for (int bounce = 1; bounce <= U.maxDepth; bounce++) {
        float dist;
        int id = sh.bvhHit(ray, dist);
        if (id == -1) break;
        if (id - S.objPrimCount >= 0) break;
        vec3 pos = rayPoint(ray, dist);
        SurfaceInfo surf = sh.triangleSurfaceInfo(id, pos);
        int matTexId = S.matTexIndices[id];
        int matId = matTexId & 0x0000ffff;
        int texId = matTexId >> 16;
        uint32_t matType = sh.loadMaterialType(matId);
        if (matType != Dielectric && matType != ThinDielectric) {
            if (dot(surf.ns, wo) < 0) { surf.ns = -surf.ns; surf.ng = -surf.ng; }
        }
        float coefToPos = remap(prevPdfDir * absDot(surf.ns, wo));
        s0t1 /= coefToPos;
        s1t1 /= coefToPos / (bounce == 1 ? remap(dist * dist) : 1.0f);
        BSDFParam matParam = sh.loadMaterial(matType, matId, texId, surf.uv);
        {
            CameraIiSample ciSamp = sh.thinLensCameraSampleIi(pos, sh.sample2D(s));
            if (ciSamp.pdf > 0) {
                vec3 pCam = pos + ciSamp.wi * ciSamp.dist;
                if (sh.visible(pos, pCam)) {
                    float cosWi = satDot(surf.ng, ciSamp.wi) * std::fabs(dot(surf.ns, wo) / dot(surf.ng, wo));
                    vec3 bsdf = sh.materialBSDF(matType, matParam, wo, ciSamp.wi, surf.ns, Importance);
                    vec3 contrib = ciSamp.Ii * bsdf * throughput * cosWi / ciSamp.pdf;
                    float coefToSurf = remap(sh.thinLensCameraPdfIe(makeRay(pCam, -ciSamp.wi)).pdfDir * satDot(surf.ns, ciSamp.wi));
                    float coefToPrev = remap(sh.materialPdf(matType, matParam, ciSamp.wi, wo, surf.ns, Radiance) *
                                             absDot(prevNorm, wo));
                    float coefDist = remap(ciSamp.dist * ciSamp.dist);
                    float coef0 = coefToSurf * coefToPrev / coefDist;
                    float coef1 = ((bounce == 1) ? 1.0f : coefToPrev) * coefToSurf / coefDist;
                    float weight = weightT1(s0t1 * coef0, s1t1 * coef1);
                    vec3 res = contrib * weight;
                    if (!hasNan(res) && !std::isnan(ciSamp.pdf) && ciSamp.pdf > 1e-8f && !isBlack(res))
                        accumulateFilmSplat(sh, film, ciSamp.uv, res, U.scale, st);
                }
            }
        }
        BSDFSample samp = sh.materialSample(matType, matParam, surf.ns, wo, Importance, sh.sample3D(s));
        vec3 wi = samp.wi;
        float bsdfPdf = samp.pdf;
        vec3 bsdf = samp.bsdf;
        uint32_t flag = samp.flag;
        bool deltaBsdf = (flag == SpecRefl || flag == SpecTrans);
        if (bsdfPdf < 1e-8f || std::isnan(bsdfPdf)) break;
        if (U.russianRoulette) {
            float continueProb = gmin(maxComponent(bsdf / bsdfPdf), 1.0f);
            if (sh.sample1D(s) >= continueProb) break;
            throughput /= continueProb;
        }
        float coefToPrev = remap(sh.materialPdf(matType, matParam, wi, wo, surf.ns, Radiance) * absDot(prevNorm, wo));
        s0t1 *= coefToPrev;
        s1t1 *= (bounce == 1) ? 1.0f : coefToPrev;
        prevPdfDir = sh.materialPdf(matType, matParam, wo, wi, surf.ns, Importance);
        prevNorm = surf.ns;
        float cosWi = deltaBsdf ? 1.0f : std::fabs(dot(surf.ng, wi) * dot(surf.ns, wo) / dot(surf.ng, wo));
        throughput *= bsdf * cosWi / bsdfPdf;
        ray = rayOffseted(pos, wi);
        wo = -wi;
    }
}

// triple_path_pass_lpt.glsl:184-193 (TriplePath.cpp:123-127)
inline void tripleLptPass(const Scene& S, const ZlRenderParams& U, float* filmPx, Stats* st,
                          long idBegin = 0, long idEnd = -1) {
    Film film{filmPx, U.filmW, U.filmH};
    long total = (long)ZL_LIGHT_GROUP_SIZE * U.blocksOnePass;
    if (idEnd < 0) idEnd = total;
    if (S.numLightTriangles <= 0) return;
#pragma omp parallel for schedule(dynamic, 256)
    for (long id = idBegin; id < idEnd; id++) {
        Shader sh(S, U, 0);     // uSampler forced to 0 (TriplePath.cpp:72)
        int sampleIdx = 0;
        sh.setRngSeed((uint32_t)U.spp * ((uint32_t)ZL_LIGHT_GROUP_SIZE * (uint32_t)U.blocksOnePass * (uint32_t)U.loopsPerPass) +
                      (uint32_t)id + (uint32_t)U.freeCounter);
        for (int i = 0; i < U.loopsPerPass; i++) traceLightPath(sh, sampleIdx, film, st);
        accumulateStats(st, sh, (uint64_t)U.loopsPerPass);
    }
}

}  // namespace zo
