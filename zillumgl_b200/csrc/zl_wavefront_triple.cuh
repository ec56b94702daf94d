// zl_wavefront_triple.cuh — wavefront stages of the triple tracer: the camera pass with the s=0 / s=1
// strategies (triple_path_pass_pt.glsl:59-195) and the light pass with the t=1 strategy
// (triple_path_pass_lpt.glsl:58-182).  Included at the end of zl_wavefront.cuh; shares WfState, the queues,
// wfGenerateKernel (camera rays), wfResolveKernel (primary miss / emitter), wfLightGenerateKernel's slot
// convention, the sort kernels and wfTraceSimpleKernel (MODE 0 with visible()'s 1e-5 offset for the PT
// pass, MODE 1 for the LPT pass).  The running MIS ratios live in WfState::aux / nrm; the distance returned by
// the closest-hit traversal (needed by remap(dist*dist)) in WfState::tdist.
#pragma once

namespace zl {

// ---------------------------------------------------------------------------------------------
// PT pass: loop body of traceCameraPath, one material type per kernel
// ---------------------------------------------------------------------------------------------
template <uint32_t TYPE>
__global__ void __launch_bounds__(128, (TYPE == 0u ? ZL_WF_STAGE_MINB : ZL_WF_SHADE_MINB_OTHER)) wfTripleShadeKernel(const DScene S, const ZlRenderParams Uin, const WfState W, float4* __restrict__ film, const int b) {
    const ZlRenderParams U = wfPassParams(Uin, W);       // graph replays read the pass index from device memory
    __shared__ uint32_t row[256];
    stageSobolRow(S, U, row);
    __syncthreads();
    int* const cnt = W.cnt + kWfCntStride * b;
    const int n = cnt[kCntIn + TYPE];
    const int* __restrict__ qin = W.qIn[TYPE];
    const int endedBase = wfEndedBase(W, b);
    const int stride = gridDim.x * blockDim.x;
    for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {
        const int i = i0 + (threadIdx.x & 31);
        const bool valid = i < n;
        const int slot = valid ? qin[i] : 0;
        bool toS = false, toE = false, toT = false;
        if (valid) {
            const float4 h = W.hit[b & 1][slot];
            const float3 pos = f3(h);
            const int id = __float_as_int(h.w);
            const float3 dIn = f3(W.dir[slot]);            // direction the path arrived with (= wi of the previous vertex)
            const float3 wo = -dIn;
            float3 throughput = f3(W.thr[slot]);
            const float4 r4 = W.res[slot];
            float3 result = f3(r4);
            const uint4 sm = W.smp[slot];
            SamplerState st = makeSampler(S, U, row, U.sampler);
            st.randSeed = sm.x; st.sampleSeed = sm.y; st.s = (int)sm.z;
            bool alive = true;
            if (b > 1) {
                const float4 c = W.shc[slot];                                  // s=1 connection of bounce b-1, visibility now known
                if (__float_as_int(c.w) != 0) result += f3(c);
                if (U.russianRoulette) {                                       // triple_path_pass_pt.glsl:176-182
                    const float continueProb = r4.w;
                    if (sample1D(st) >= continueProb) {      // ended by roulette: film write left to the resolve kernel (see wfShadeKernel)
                        alive = false; toT = true;
                        W.res[slot] = make_float4(result.x, result.y, result.z, 1.0f);
                        W.shc[slot] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0));
                        W.thr[slot] = make_float4(throughput.x, throughput.y, throughput.z, __int_as_float(2));
                    } else throughput /= continueProb;
                }
            }
            if (alive) {
                SurfaceInfo surf = triangleSurfaceInfo(S, id, pos);
                const float3 nsRaw = surf.ns;                                  // = triangleNormalShad(id, pos) (same expression, intersection.glsl:149-173 / 188-224)
                float t1s0, t1s1;
                float3 prevNorm;
                if (b == 1) {                                                  // :68-77
                    const Ray ray = makeRay(f3(W.hit[0][slot]), dIn);
                    const float primDist = W.tdist[slot];
                    CameraPdf camPdf = thinLensCameraPdfIe(U, ray);
                    const float primaryPdf = remap(camPdf.pdfPos) / remap(camPdf.pdfDir * absDot(nsRaw, ray.dir) / square(primDist));
                    t1s0 = primaryPdf; t1s1 = primaryPdf;
                    prevNorm = camF(U);
                } else {                                                       // end of iteration b-1 (:183-190)
                    const float4 a = W.aux[slot];
                    const float coef = a.z / remap(a.w * absDot(nsRaw, dIn));
                    t1s0 = a.x * coef; t1s1 = a.y * coef;
                    prevNorm = f3(W.nrm[slot]);
                }
                const int matTexId = __ldg(&S.matTex[id]);
                const int matId = matTexId & 0x0000ffff, texId = matTexId >> 16;
                if (TYPE != Dielectric && TYPE != ThinDielectric) {
                    if (dot(surf.ns, wo) < 0) { surf.ns = -surf.ns; surf.ng = -surf.ng; }
                }
                const BSDFParam mat = loadMaterial(S, TYPE, matId, texId, surf.uv);
                const float3 ns = surf.ns;
                float4 shOut = make_float4(0.0f, 0.0f, 0.0f, 0.0f), shcOut = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0));
                {   // s = 1: connect to a sampled point on an area light (:84-112)
                    const int light = lightSampleOne(S, sample2D(st));
                    const int triId = light + S.objPrimCount;
                    const float pdfSource = lightPdfSampleOne(S, light);
                    const float3 pLit = triangleSampleUniform(S, triId, sample2D(st));
                    const float3 wi = normalize(pLit - pos);
                    const float3 Le = lightLe(S, light, pLit, -wi);
                    if (!isBlack(Le)) {
                        const float3 nLit = triangleNg(S, triId, pLit);
                        const float pA = pdfSource / triangleAreaId(S, triId);
                        const float dist2 = distSquare(pos, pLit);
                        const float pS = pA * dist2 / absDot(nLit, wi);
                        const float4 bsdfAndPdf = materialBSDFAndPdfT<TYPE>(mat, wo, wi, ns, Radiance);
                        const float pdfRev = materialPdfT<TYPE>(mat, wi, wo, ns, Importance);
                        const float pdfPLit = remap(pA);
                        const float coefToSurf = remap(0.5f * PiInv * absDot(ns, wi));
                        const float coefToLight = remap(bsdfAndPdf.w * satDot(nLit, -wi));
                        const float coefToPrev = (b == 1) ? 1.0f : remap(pdfRev * absDot(prevNorm, wo));
                        const float coefDist = remap(dist2);
                        const float weight = weightS1(pdfPLit * coefDist / coefToLight, t1s1 * coefToSurf * coefToPrev / coefDist);
                        const float3 contrib = Le * f3(bsdfAndPdf) * throughput * absDot(ns, wi) / pS * weight;
                        const VisRay v = visibleRay(pos, pLit);                // origin = pos + dir * 1e-5: the trace kernel's shadowEps
                        shOut = make_float4(v.dir.x, v.dir.y, v.dir.z, v.dist);
                        shcOut = make_float4(contrib.x, contrib.y, contrib.z, __int_as_float(1));
                        toS = !(contrib.x == 0.0f && contrib.y == 0.0f && contrib.z == 0.0f);      // a zero contribution needs no visibility test (see wfShadeKernel)
                    }
                }
                BSDFSample samp = materialSampleT<TYPE>(mat, ns, wo, Radiance, sample3D(st), st);
                const float3 wi = samp.wi;
                const float bsdfPdf = samp.pdf;
                const bool deltaBsdf = (samp.flag == SpecRefl || samp.flag == SpecTrans);
                int flags = deltaBsdf ? 1 : 0;
                float rrProb = 1.0f;
                if (bsdfPdf < 1e-8f) { flags |= 2; toT = true; }
                else {
                    throughput *= samp.bsdf / bsdfPdf * (deltaBsdf ? 1.0f : absDot(ns, wi));
                    rrProb = gmin(maxComponent(samp.bsdf / bsdfPdf), 0.95f);
                    const float pdfDirToNext = materialPdfT<TYPE>(mat, wo, wi, ns, Radiance);
                    const float pdfDirToPrev = materialPdfT<TYPE>(mat, wi, wo, ns, Importance);
                    const float coefToPrev = (b == 1) ? 1.0f : remap(pdfDirToPrev * absDot(prevNorm, wo));
                    W.aux[slot] = make_float4(t1s0, t1s1, coefToPrev, pdfDirToNext);
                    W.nrm[slot] = make_float4(ns.x, ns.y, ns.z, 0.0f);
                    toE = true;
                }
                W.dir[slot] = make_float4(wi.x, wi.y, wi.z, bsdfPdf);
                W.thr[slot] = make_float4(throughput.x, throughput.y, throughput.z, __int_as_float(flags));
                W.res[slot] = make_float4(result.x, result.y, result.z, rrProb);
                W.smp[slot] = make_uint4(st.randSeed, st.sampleSeed, (uint32_t)st.s, 0u);
                if (toS) W.sh[slot] = shOut;
                W.shc[slot] = shcOut;
            }
        }
        wfAppendRays(S, W, cnt, b, slot, toS, toE, false);
        wfAppend(W.qT + endedBase, cnt + kCntT, toT, slot);
    }
}

// camera paths that end at bounce b >= 1 (triple_path_pass_pt.glsl:148-175): s=1 result, s=0 weight of an emitter hit
__global__ void __launch_bounds__(128, ZL_WF_STAGE_MINB) wfTripleResolveKernel(const DScene S, const ZlRenderParams Uin, const WfState W, float4* __restrict__ film, const int b) {
    const ZlRenderParams U = wfPassParams(Uin, W);       // graph replays read the pass index from device memory
    const int n = W.cnt[kWfCntStride * b + kCntT];
    const int* __restrict__ qT = W.qT + wfEndedBase(W, b);
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int slot = qT[i];
        float3 result = f3(W.res[slot]);
        const float4 c = W.shc[slot];
        if (__float_as_int(c.w) != 0) result += f3(c);
        const float4 t4 = W.thr[slot];
        const int flags = __float_as_int(t4.w);
        if (!(flags & 2)) {
            const float4 hn = W.hit[(b + 1) & 1][slot];
            const int nextId = __float_as_int(hn.w);
            const int lightId = nextId - S.objPrimCount;
            if (nextId != -1 && lightId >= 0) {
                const float3 wi = f3(W.dir[slot]), throughput = f3(t4), nextPos = f3(hn);
                const float4 a = W.aux[slot];
                const float3 ns = f3(W.nrm[slot]);
                const float dist = W.tdist[slot];
                const float t1s0 = a.x, coefToPrev = a.z, pdfDirToNext = a.w;
                const float3 nLit = triangleNg(S, nextId, nextPos);
                const LightPdf pdfLit = lightPdfLe(S, lightId, makeRay(nextPos, -wi));
                const float pdfPLit = remap(pdfLit.pdfPos * lightPdfSampleOne(S, lightId));
                const float coefToLight = remap(pdfDirToNext * satDot(nLit, -wi));
                const float coefToSurf = remap(pdfLit.pdfDir * absDot(ns, wi));
                const float coefDist = remap(dist * dist);
                const float weight = isnan(t1s0) ? 0.0f : weightS0(pdfPLit * coefDist / coefToLight, t1s0 * coefToSurf * pdfPLit * coefToPrev / coefToLight);
                result += lightLe(S, lightId, nextPos, -wi) * throughput * weight;
            }
        }
        wfFilmAdd(W, U, film, slot, result);
    }
}

// ---------------------------------------------------------------------------------------------
// LPT pass
// ---------------------------------------------------------------------------------------------
// first part of traceLightPath (triple_path_pass_lpt.glsl:58-92); seed and `resume` as in wfLightGenerateKernel
__global__ void __launch_bounds__(128, ZL_WF_STAGE_MINB) wfTripleLightGenerateKernel(const DScene S, const ZlRenderParams Uin, const WfState W, const long long total,
                                                                  const uint32_t seedMul, const int resume) {
    const ZlRenderParams U = wfPassParams(Uin, W);       // graph replays read the pass index from device memory
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = id < total;
    const int slot = (int)id;
    if (valid) {
        SamplerState st = makeSampler(S, U, nullptr, 0);                      // TriplePath.cpp:72
        st.randSeed = resume ? W.smp[slot].x : (uint32_t)U.spp * seedMul + (uint32_t)id + (uint32_t)U.freeCounter;
        const int light = lightSampleOne(S, sample2D(st));
        const float pdfSource = lightPdfSampleOne(S, light);
        const int triId = light + S.objPrimCount;
        (void)sample2D(st);                                                   // `pLit` is drawn and never used (App. B #19)
        LightLeSample leSamp = lightSampleOneLe(S, light, sample4D(st));
        const float3 nl = triangleNg(S, triId, leSamp.ray.ori);
        const float3 wo = -leSamp.ray.dir;
        const Ray ray = rayOffseted(leSamp.ray);
        const float3 throughput = leSamp.Le * absDot(nl, -wo) / (pdfSource * leSamp.pdfPos * leSamp.pdfDir);
        W.hit[0][slot] = make_float4(ray.ori.x, ray.ori.y, ray.ori.z, __int_as_float(-1));
        W.dir[slot] = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, 0.0f);
        W.thr[slot] = make_float4(throughput.x, throughput.y, throughput.z, 0.0f);
        W.smp[slot] = make_uint4(st.randSeed, 0u, 0u, 0u);
        W.aux[slot] = make_float4(1.0f / remap(leSamp.pdfPos * pdfSource), 1.0f, leSamp.pdfDir, 0.0f);   // {s0t1, s1t1, prevPdfDir}
        W.nrm[slot] = make_float4(nl.x, nl.y, nl.z, 0.0f);
    }
    wfAppend(W.qE, W.cnt + kCntE, valid && U.maxDepth >= 1, slot);
}

// loop body of traceLightPath after the bvhHit (triple_path_pass_lpt.glsl:99-180)
template <uint32_t TYPE>
__global__ void __launch_bounds__(128, (TYPE == 0u ? ZL_WF_STAGE_MINB : ZL_WF_SHADE_MINB_OTHER)) wfTripleLightShadeKernel(const DScene S, const ZlRenderParams Uin, const WfState W, const int b) {
    const ZlRenderParams U = wfPassParams(Uin, W);       // graph replays read the pass index from device memory
    int* const cnt = W.cnt + kWfCntStride * b;
    const int n = cnt[kCntIn + TYPE];
    const int* __restrict__ qin = W.qIn[TYPE];
    const int stride = gridDim.x * blockDim.x;
    for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {
        const int i = i0 + (threadIdx.x & 31);
        const bool valid = i < n;
        const int slot = valid ? qin[i] : 0;
        bool toS = false, toE = false;
        if (valid) {
            const float4 h = W.hit[b & 1][slot];
            const float3 pos = f3(h);
            const int id = __float_as_int(h.w);
            const float3 wo = -f3(W.dir[slot]);
            float3 throughput = f3(W.thr[slot]);
            SamplerState st = makeSampler(S, U, nullptr, 0);
            st.randSeed = W.smp[slot].x;
            const float4 a = W.aux[slot];
            float s0t1 = a.x, s1t1 = a.y;
            const float prevPdfDir = a.z;
            const float3 prevNorm = f3(W.nrm[slot]);
            const float dist = W.tdist[slot];
            SurfaceInfo surf = triangleSurfaceInfo(S, id, pos);
            const int matTexId = __ldg(&S.matTex[id]);
            const int matId = matTexId & 0x0000ffff, texId = matTexId >> 16;
            if (TYPE != Dielectric && TYPE != ThinDielectric) {
                if (dot(surf.ns, wo) < 0) { surf.ns = -surf.ns; surf.ng = -surf.ng; }
            }
            const BSDFParam mat = loadMaterial(S, TYPE, matId, texId, surf.uv);
            const float3 ns = surf.ns, ng = surf.ng;
            const float coefToPos = remap(prevPdfDir * absDot(ns, wo));
            s0t1 /= coefToPos;
            s1t1 /= coefToPos / (b == 1 ? remap(dist * dist) : 1.0f);
            {
                CameraIiSample ciSamp = thinLensCameraSampleIi(U, pos, sample2D(st));
                if (ciSamp.pdf > 0) {
                    const float3 pCam = pos + ciSamp.wi * ciSamp.dist;
                    const float cosWi = satDot(ng, ciSamp.wi) * fabsf(dot(ns, wo) / dot(ng, wo));
                    const float3 bsdf = materialBSDFT<TYPE>(mat, wo, ciSamp.wi, ns, Importance);
                    const float3 contrib = ciSamp.Ii * bsdf * throughput * cosWi / ciSamp.pdf;
                    const float coefToSurf = remap(thinLensCameraPdfIe(U, makeRay(pCam, -ciSamp.wi)).pdfDir * satDot(ns, ciSamp.wi));
                    const float coefToPrev = remap(materialPdfT<TYPE>(mat, ciSamp.wi, wo, ns, Radiance) * absDot(prevNorm, wo));
                    const float coefDist = remap(ciSamp.dist * ciSamp.dist);
                    const float coef0 = coefToSurf * coefToPrev / coefDist;
                    const float coef1 = ((b == 1) ? 1.0f : coefToPrev) * coefToSurf / coefDist;
                    const float weight = weightT1(s0t1 * coef0, s1t1 * coef1);
                    const float3 res = contrib * weight;
                    if (!hasNan(res) && !isnan(ciSamp.pdf) && ciSamp.pdf > 1e-8f && !isBlack(res) && inFilmBound(ciSamp.uv)) {
                        wfStoreSplat(W, slot, visibleRay(pos, pCam), ciSamp.uv, res * U.scale);
                        toS = true;
                    }
                }
            }
            BSDFSample samp = materialSampleT<TYPE>(mat, ns, wo, Importance, sample3D(st), st);
            const float3 wi = samp.wi;
            const float bsdfPdf = samp.pdf;
            const bool deltaBsdf = (samp.flag == SpecRefl || samp.flag == SpecTrans);
            bool alive = !(bsdfPdf < 1e-8f || isnan(bsdfPdf));
            if (alive && U.russianRoulette) {
                const float continueProb = gmin(maxComponent(samp.bsdf / bsdfPdf), 1.0f);
                if (sample1D(st) >= continueProb) alive = false;
                else throughput /= continueProb;
            }
            if (alive && b < U.maxDepth) {
                const float coefToPrev = remap(materialPdfT<TYPE>(mat, wi, wo, ns, Radiance) * absDot(prevNorm, wo));
                s0t1 *= coefToPrev;
                s1t1 *= (b == 1) ? 1.0f : coefToPrev;
                const float nextPdfDir = materialPdfT<TYPE>(mat, wo, wi, ns, Importance);
                const float cosWi = deltaBsdf ? 1.0f : fabsf(dot(ng, wi) * dot(ns, wo) / dot(ng, wo));
                throughput *= samp.bsdf * cosWi / bsdfPdf;
                W.dir[slot] = make_float4(wi.x, wi.y, wi.z, bsdfPdf);
                W.thr[slot] = make_float4(throughput.x, throughput.y, throughput.z, 0.0f);
                W.aux[slot] = make_float4(s0t1, s1t1, nextPdfDir, 0.0f);
                W.nrm[slot] = make_float4(ns.x, ns.y, ns.z, 0.0f);
                toE = true;
            }
            W.smp[slot] = make_uint4(st.randSeed, 0u, 0u, 0u);
        }
        wfAppendRays(S, W, cnt, b, slot, toS, toE, true);
    }
}

}  // namespace zl
