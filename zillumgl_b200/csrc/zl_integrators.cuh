// zl_integrators.cuh — the four integrator kernels as device functions:
//   pathIntegTrace   (path_integ_naive.glsl:35-143)   MIS path tracer
//   lightIntegTrace  (light_path_integ.glsl:45-146)   adjoint light tracer with camera splatting
//   traceCameraPath  (triple_path_pass_pt.glsl:59-195) triple tracer, s=0 / s=1 strategies
//   traceLightPath   (triple_path_pass_lpt.glsl:58-182) triple tracer, t=1 strategy
// Film: W*H float4, row 0 = bottom; rgb running sums.  Splats use one
// red.global.add.v4.f32 (sm_90+) instead of three scalar float atomics on a 3WxH r32f image.
#pragma once
#include "zl_shading.cuh"

namespace zl {

// light_path_integ.glsl:34-43 / triple_path_pass_lpt.glsl:36-46
ZL_DEV void accumulateFilm(const DScene& S, float4* __restrict__ film, const ZlRenderParams& U, float2 uv, float3 res) {
    if (!inFilmBound(uv)) return;
    int ix = (int)(uv.x * (float)U.filmW), iy = (int)(uv.y * (float)U.filmH);
    if (ix < 0 || iy < 0 || ix >= U.filmW || iy >= U.filmH) return;   // uv == 1.0: GL drops the OOB write (App. B #14)
    countEvent(S, 4);
    float4* p = film + (size_t)iy * U.filmW + ix;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(res.x), "f"(res.y), "f"(res.z), "f"(0.0f) : "memory");
}

// path_integ_naive.glsl:35-143
ZL_DEV float3 pathIntegTrace(const DScene& S, const ZlRenderParams& U, Ray ray, SamplerState& st) {
    float primDist;
    int id = bvhHit(S, ray, primDist);
    float3 pos = rayPoint(ray, primDist);
    if (id == -1) return envLe(S, U, ray.dir);
    else if (id - S.objPrimCount >= 0) return lightLe(S, id - S.objPrimCount, pos, -ray.dir);

    float3 wo = -ray.dir;
    float3 result = f3(0.0f);
    float3 throughput = f3(1.0f);

    for (int bounce = 1; bounce <= U.maxDepth; bounce++) {
        ShadingPoint sp = loadShadingPoint(S, id, triangleSurfaceInfo(S, id, pos), wo);
        const float3 ns = sp.surf.ns;

        if (U.sampleLight) {
            float ud = sample1D(st);
            float4 us = sample4D(st);
#ifndef ZL_INSTRUMENT
            LightLiSample samp = sampleLightAndEnv(S, U, pos, ud, us);
            if (samp.pdf > 0.0f) {
                float4 bsdfAndPdf = materialBSDFAndPdf(sp.matType, sp.mat, wo, samp.wi, ns, Radiance);
                float weight = biHeuristic(samp.pdf, bsdfAndPdf.w);
                result += f3(bsdfAndPdf) * throughput * satDot(ns, samp.wi) * samp.coef * weight;
            }
#else
            // Counting build: the same sample with the shadow ray deferred, so that the rays the wavefront variant does NOT trace —
            // a sample rejected after the test (pdf < 1e-8) or a contribution of exactly zero, wfShadeKernel — are counted apart
            // (counters 6..8: rays, entries, triangle tests).  The reference casts them all; both totals are reported.
            DeferredVis vis; vis.pending = false; vis.dist = 0.0f; vis.ray = makeRay(pos, f3(0.0f));
            LightLiSample samp = sampleLightAndEnv(S, U, pos, ud, us, vis);
            float3 contrib = f3(0.0f);
            bool traced = false;
            if (samp.pdf > 0.0f) {
                float4 bsdfAndPdf = materialBSDFAndPdf(sp.matType, sp.mat, wo, samp.wi, ns, Radiance);
                float weight = biHeuristic(samp.pdf, bsdfAndPdf.w);
                contrib = f3(bsdfAndPdf) * throughput * satDot(ns, samp.wi) * samp.coef * weight;
                traced = !(contrib.x == 0.0f && contrib.y == 0.0f && contrib.z == 0.0f);
            }
            if (vis.pending) {
                TraceCounters c{0, 0};
                const bool occluded = bvhTestCall<true>(S.nodes, S.triPos, S.bvhSize, vis.ray.ori, vis.ray.dir, vis.dist, &c) != 0;
                countRay(S, c);
                if (!traced && S.counters) {
                    atomicAdd(S.counters + 6, 1ull);
                    atomicAdd(S.counters + 7, (unsigned long long)c.nodes);
                    atomicAdd(S.counters + 8, (unsigned long long)c.tris);
                }
                if (!occluded) result += contrib;
            }
#endif
        }

        BSDFSample samp = materialSample(sp.matType, sp.mat, ns, wo, Radiance, sample3D(st), st);
        float3 wi = samp.wi;
        float bsdfPdf = samp.pdf;
        float3 bsdf = samp.bsdf;
        bool deltaBsdf = (samp.flag == SpecRefl || samp.flag == SpecTrans);
        if (bsdfPdf < 1e-8f) break;
        throughput *= bsdf / bsdfPdf * (deltaBsdf ? 1.0f : absDot(ns, wi));

        ray = rayOffseted(pos, wi);
        float dist;
        int nextId = bvhHit(S, ray, dist);
        int lightId = nextId - S.objPrimCount;
        float3 nextPos = rayPoint(ray, dist);

        if (nextId == -1) {
            float3 radiance = envLe(S, U, wi);
            float weight = 1.0f;
            if (U.sampleLight && !deltaBsdf) {
                float envPdf = envPdfLi(S, U, wi) * pdfSelectEnv(S, U);
                weight = (envPdf <= 0.0f) ? 0.0f : biHeuristic(bsdfPdf, envPdf);
            }
            result += radiance * throughput * weight;
            break;
        } else if (lightId >= 0) {
            float3 radiance = lightLe(S, lightId, nextPos, -wi);
            float weight = 1.0f;
            if (U.sampleLight && !deltaBsdf) {
                float lightPdf = lightPdfLi(S, lightId, pos, nextPos) * pdfSelectLight(S, U, lightId);
                weight = (lightPdf <= 0.0f) ? 0.0f : biHeuristic(bsdfPdf, lightPdf);
            }
            result += radiance * throughput * weight;
            break;
        }
        if (U.russianRoulette) {
            float continueProb = gmin(maxComponent(bsdf / bsdfPdf), 0.95f);
            if (sample1D(st) >= continueProb) break;
            throughput /= continueProb;
        }
        id = nextId;
        pos = nextPos;
        wo = -wi;
    }
    return result;
}

// light_path_integ.glsl:45-146
ZL_DEV void lightIntegTrace(const DScene& S, const ZlRenderParams& U, SamplerState& st, float4* __restrict__ film) {
    int light = lightSampleOne(S, sample2D(st));
    float pdfSource = lightPdfSampleOne(S, light);
    Ray ray; float3 wo; float3 throughput;
    {
        int triId = light + S.objPrimCount;
        float3 pLit = triangleSampleUniform(S, triId, sample2D(st));
        CameraIiSample ciSamp = thinLensCameraSampleIi(U, pLit, sample2D(st));
        if (ciSamp.pdf > 0) {
            float3 pCam = pLit + ciSamp.wi * ciSamp.dist;
            float pdfPos = 1.0f / triangleAreaId(S, triId);
            if (visible(S, pLit, pCam)) {
                float3 Le = lightLe(S, light, pLit, ciSamp.wi);
                float3 contrib = Le * ciSamp.Ii / (ciSamp.pdf * pdfPos * pdfSource);
                if (!isBlack(contrib)) accumulateFilm(S, film, U, ciSamp.uv, contrib);
            }
        }
        LightLeSample leSamp = lightSampleOneLe(S, light, sample4D(st));
        float3 nl = triangleNg(S, triId, leSamp.ray.ori);
        wo = -leSamp.ray.dir;
        ray = rayOffseted(leSamp.ray);
        throughput = leSamp.Le * absDot(nl, -wo) / (pdfSource * leSamp.pdfPos * leSamp.pdfDir);
    }
    for (int bounce = 1; bounce <= U.maxDepth; bounce++) {
        float dist;
        int id = bvhHit(S, ray, dist);
        if (id == -1) break;
        if (id - S.objPrimCount >= 0) break;
        float3 pos = rayPoint(ray, dist);
        ShadingPoint sp = loadShadingPoint(S, id, triangleSurfaceInfo(S, id, pos), wo);
        const float3 ns = sp.surf.ns, ng = sp.surf.ng;
        {
            CameraIiSample ciSamp = thinLensCameraSampleIi(U, pos, sample2D(st));
            if (ciSamp.pdf > 0) {
                float3 pCam = pos + ciSamp.wi * ciSamp.dist;
                if (visible(S, pos, pCam)) {
                    float3 bsdf = materialBSDF(sp.matType, sp.mat, wo, ciSamp.wi, ns, Importance);
                    float cosWi = satDot(ng, ciSamp.wi) * fabsf(dot(ns, wo) / dot(ng, wo));
                    float3 res = ciSamp.Ii * bsdf * throughput * cosWi / ciSamp.pdf;
                    if (!hasNan(res) && !isnan(ciSamp.pdf) && ciSamp.pdf > 1e-8f && !isBlack(res))
                        accumulateFilm(S, film, U, ciSamp.uv, res);
                }
            }
        }
        BSDFSample samp = materialSample(sp.matType, sp.mat, ns, wo, Importance, sample3D(st), st);
        float3 wi = samp.wi;
        float bsdfPdf = samp.pdf;
        float3 bsdf = samp.bsdf;
        bool deltaBsdf = (samp.flag == SpecRefl || samp.flag == SpecTrans);
        if (bsdfPdf < 1e-8f || isnan(bsdfPdf)) break;
        if (U.russianRoulette) {
            float continueProb = gmin(maxComponent(bsdf / bsdfPdf), 1.0f);
            if (sample1D(st) >= continueProb) break;
            throughput /= continueProb;
        }
        float cosWi = deltaBsdf ? 1.0f : fabsf(dot(ng, wi) * dot(ns, wo) / dot(ng, wo));
        throughput *= bsdf * cosWi / bsdfPdf;
        ray = rayOffseted(pos, wi);
        wo = -wi;
    }
}

// triple_path_pass_pt.glsl:44-57 / triple_path_pass_lpt.glsl:48-56
ZL_DEV float remap(float p) { return p < 1e-8f ? 1.0f : p * p; }
ZL_DEV float weightS0(float s1s0, float t1s0) { return 1.0f / (1.0f + s1s0 + t1s0); }
ZL_DEV float weightS1(float s1s0, float t1s1) { return s1s0 / (1.0f + s1s0 + s1s0 * t1s1); }
ZL_DEV float weightT1(float s0t1, float s1t1) { return 1.0f / (s0t1 + s1t1 + 1.0f); }

// triple_path_pass_pt.glsl:59-195
ZL_DEV float3 traceCameraPath(const DScene& S, const ZlRenderParams& U, Ray ray, SamplerState& st) {
    float primDist;
    int id = bvhHit(S, ray, primDist);
    float3 pos = rayPoint(ray, primDist);
    if (id == -1) return envLe(S, U, ray.dir);
    else if (id - S.objPrimCount >= 0) return lightLe(S, id - S.objPrimCount, pos, -ray.dir);

    SurfaceInfo surf0 = triangleSurfaceInfo(S, id, pos);
    float3 wo = -ray.dir;
    float3 prevNorm = camF(U);
    float3 result = f3(0.0f);
    float3 throughput = f3(1.0f);

    CameraPdf camPdf = thinLensCameraPdfIe(U, ray);
    float primaryPdf = remap(camPdf.pdfPos) / remap(camPdf.pdfDir * absDot(surf0.ns, ray.dir) / square(primDist));
    float t1s0 = primaryPdf;
    float t1s1 = primaryPdf;

    for (int bounce = 1; bounce <= U.maxDepth; bounce++) {
        ShadingPoint sp = loadShadingPoint(S, id, (bounce > 1) ? triangleSurfaceInfo(S, id, pos) : surf0, wo);
        const float3 ns = sp.surf.ns;
        {
            int light = lightSampleOne(S, sample2D(st));
            int triId = light + S.objPrimCount;
            float pdfSource = lightPdfSampleOne(S, light);
            float3 pLit = triangleSampleUniform(S, triId, sample2D(st));
            float3 wi = normalize(pLit - pos);
            float3 Le = lightLe(S, light, pLit, -wi);
            if (!isBlack(Le) && visible(S, pos, pLit)) {
                float3 nLit = triangleNg(S, triId, pLit);
                float pA = pdfSource / triangleAreaId(S, triId);
                float dist2 = distSquare(pos, pLit);
                float pS = pA * dist2 / absDot(nLit, wi);
                float4 bsdfAndPdf = materialBSDFAndPdf(sp.matType, sp.mat, wo, wi, ns, Radiance);
                float pdfRev = materialPdf(sp.matType, sp.mat, wi, wo, ns, Importance);
                float pdfPLit = remap(pA);
                float coefToSurf = remap(0.5f * PiInv * absDot(ns, wi));
                float coefToLight = remap(bsdfAndPdf.w * satDot(nLit, -wi));
                float coefToPrev = (bounce == 1) ? 1.0f : remap(pdfRev * absDot(prevNorm, wo));
                float coefDist = remap(dist2);
                float weight = weightS1(pdfPLit * coefDist / coefToLight, t1s1 * coefToSurf * coefToPrev / coefDist);
                result += Le * f3(bsdfAndPdf) * throughput * absDot(ns, wi) / pS * weight;
            }
        }
        BSDFSample samp = materialSample(sp.matType, sp.mat, ns, wo, Radiance, sample3D(st), st);
        float3 wi = samp.wi;
        float bsdfPdf = samp.pdf;
        float3 bsdf = samp.bsdf;
        bool deltaBsdf = (samp.flag == SpecRefl || samp.flag == SpecTrans);
        if (bsdfPdf < 1e-8f) break;
        throughput *= bsdf / bsdfPdf * (deltaBsdf ? 1.0f : absDot(ns, wi));

        Ray nextRay = rayOffseted(pos, wi);
        float dist;
        int nextId = bvhHit(S, nextRay, dist);
        int lightId = nextId - S.objPrimCount;
        float3 nextPos = rayPoint(nextRay, dist);
        float pdfDirToNext = materialPdf(sp.matType, sp.mat, wo, wi, ns, Radiance);
        float pdfDirToPrev = materialPdf(sp.matType, sp.mat, wi, wo, ns, Importance);

        if (nextId == -1) break;
        else if (lightId >= 0) {
            float3 nLit = triangleNg(S, nextId, nextPos);
            LightPdf pdfLit = lightPdfLe(S, lightId, makeRay(nextPos, -wi));
            float pdfPLit = remap(pdfLit.pdfPos * lightPdfSampleOne(S, lightId));
            float coefToLight = remap(pdfDirToNext * satDot(nLit, -wi));
            float coefToSurf = remap(pdfLit.pdfDir * absDot(ns, wi));
            float coefToPrev = (bounce == 1) ? 1.0f : remap(pdfDirToPrev * absDot(prevNorm, wo));
            float coefDist = remap(dist * dist);
            float weight = isnan(t1s0) ? 0.0f : weightS0(pdfPLit * coefDist / coefToLight,
                                                         t1s0 * coefToSurf * pdfPLit * coefToPrev / coefToLight);
            result += lightLe(S, lightId, nextPos, -wi) * throughput * weight;
            break;
        }
        if (U.russianRoulette) {
            float continueProb = gmin(maxComponent(bsdf / bsdfPdf), 0.95f);
            if (sample1D(st) >= continueProb) break;
            throughput /= continueProb;
        }
        float coef = ((bounce == 1) ? 1.0f : remap(pdfDirToPrev * absDot(prevNorm, wo))) /
                     remap(pdfDirToNext * absDot(triangleNormalShad(S, nextId, nextPos), wi));
        t1s0 *= coef;
        t1s1 *= coef;
        prevNorm = ns;
        pos = nextPos;
        wo = -wi;
        id = nextId;
    }
    return result;
}

// triple_path_pass_lpt.glsl:58-182
ZL_DEV void traceLightPath(const DScene& S, const ZlRenderParams& U, SamplerState& st, float4* __restrict__ film) {
    int light = lightSampleOne(S, sample2D(st));
    float pdfSource = lightPdfSampleOne(S, light);
    Ray ray; float3 wo; float3 throughput;
    float3 prevNorm; float prevPdfDir;
    float s0t1, s1t1;
    {
        int triId = light + S.objPrimCount;
        (void)sample2D(st);   // `pLit` is drawn and never used (triple_path_pass_lpt.glsl:75, App. B #19)
        LightLeSample leSamp = lightSampleOneLe(S, light, sample4D(st));
        float3 nl = triangleNg(S, triId, leSamp.ray.ori);
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
        int id = bvhHit(S, ray, dist);
        if (id == -1) break;
        if (id - S.objPrimCount >= 0) break;
        float3 pos = rayPoint(ray, dist);
        ShadingPoint sp = loadShadingPoint(S, id, triangleSurfaceInfo(S, id, pos), wo);
        const float3 ns = sp.surf.ns, ng = sp.surf.ng;

        float coefToPos = remap(prevPdfDir * absDot(ns, wo));
        s0t1 /= coefToPos;
        s1t1 /= coefToPos / (bounce == 1 ? remap(dist * dist) : 1.0f);
        {
            CameraIiSample ciSamp = thinLensCameraSampleIi(U, pos, sample2D(st));
            if (ciSamp.pdf > 0) {
                float3 pCam = pos + ciSamp.wi * ciSamp.dist;
                if (visible(S, pos, pCam)) {
                    float cosWi = satDot(ng, ciSamp.wi) * fabsf(dot(ns, wo) / dot(ng, wo));
                    float3 bsdf = materialBSDF(sp.matType, sp.mat, wo, ciSamp.wi, ns, Importance);
                    float3 contrib = ciSamp.Ii * bsdf * throughput * cosWi / ciSamp.pdf;
                    float coefToSurf = remap(thinLensCameraPdfIe(U, makeRay(pCam, -ciSamp.wi)).pdfDir * satDot(ns, ciSamp.wi));
                    float coefToPrev = remap(materialPdf(sp.matType, sp.mat, ciSamp.wi, wo, ns, Radiance) * absDot(prevNorm, wo));
                    float coefDist = remap(ciSamp.dist * ciSamp.dist);
                    float coef0 = coefToSurf * coefToPrev / coefDist;
                    float coef1 = ((bounce == 1) ? 1.0f : coefToPrev) * coefToSurf / coefDist;
                    float weight = weightT1(s0t1 * coef0, s1t1 * coef1);
                    float3 res = contrib * weight;
                    if (!hasNan(res) && !isnan(ciSamp.pdf) && ciSamp.pdf > 1e-8f && !isBlack(res))
                        accumulateFilm(S, film, U, ciSamp.uv, res * U.scale);
                }
            }
        }
        BSDFSample samp = materialSample(sp.matType, sp.mat, ns, wo, Importance, sample3D(st), st);
        float3 wi = samp.wi;
        float bsdfPdf = samp.pdf;
        float3 bsdf = samp.bsdf;
        bool deltaBsdf = (samp.flag == SpecRefl || samp.flag == SpecTrans);
        if (bsdfPdf < 1e-8f || isnan(bsdfPdf)) break;
        if (U.russianRoulette) {
            float continueProb = gmin(maxComponent(bsdf / bsdfPdf), 1.0f);
            if (sample1D(st) >= continueProb) break;
            throughput /= continueProb;
        }
        float coefToPrev = remap(materialPdf(sp.matType, sp.mat, wi, wo, ns, Radiance) * absDot(prevNorm, wo));
        s0t1 *= coefToPrev;
        s1t1 *= (bounce == 1) ? 1.0f : coefToPrev;
        prevPdfDir = materialPdf(sp.matType, sp.mat, wo, wi, ns, Importance);
        prevNorm = ns;
        float cosWi = deltaBsdf ? 1.0f : fabsf(dot(ng, wi) * dot(ns, wo) / dot(ng, wo));
        throughput *= bsdf * cosWi / bsdfPdf;
        ray = rayOffseted(pos, wi);
        wo = -wi;
    }
}

}  // namespace zl
