// zl_wavefront_light.cuh — wavefront stages of the adjoint light tracer (light_path_integ.glsl:45-146).
// Included at the end of zl_wavefront.cuh (same namespace, same WfState, same queues, same trace and
// sort kernels).  Slot = light-path invocation id.  Stage order of one pass:
//   wfLightGenerateKernel        light pick, point on the light, camera connection of the light point
//                                (deferred visible() ray + splat), emission ray                 -> S(0), E(0)
//   per b = 0..maxDepth:  sort, wfTraceSimpleKernel<MODE 1>(b)  (unoccluded connections splat; hits -> qIn[type])
//                         wfLightShadeKernel<TYPE>(b+1)          surface, camera connection, BSDF sample, RR -> S, E
// The megakernel spends 76 % of its stall samples waiting for instructions and runs with 9.8 of 32
// lanes (profiles/r1_ncu_lightPassKernel_megakernel.csv); here each shade kernel holds one BSDF.
// Splat order differs from the megakernel's (atomics), values per path do not.
#pragma once

namespace zl {

// visible(x, y) of intersection.glsl:429-434 as a ray to be traced later
struct VisRay { float3 ori, dir; float dist; };
ZL_DEV VisRay visibleRay(float3 x, float3 y) {
    VisRay v;
    v.dist = distance(x, y) - 2e-5f;
    v.dir = normalize(y - x);
    v.ori = x + v.dir * 1e-5f;
    return v;
}
ZL_DEV void wfStoreSplat(const WfState& W, int slot, const VisRay& v, float2 uv, float3 contrib) {
    W.sho[slot] = make_float4(v.ori.x, v.ori.y, v.ori.z, v.dist);
    W.sh[slot] = make_float4(v.dir.x, v.dir.y, v.dir.z, uv.x);
    W.shc[slot] = make_float4(contrib.x, contrib.y, contrib.z, uv.y);
}

// first part of lightIntegTrace (light_path_integ.glsl:45-78).  `seedMul` = invocations per pass (uSpp stride
// of the seed, LightPath.cpp / light_path_integ.glsl:153); `resume` != 0 continues the RNG stream left in smp
// by the previous loop of the same invocation (triple LPT runs uLoopsPerPass paths per invocation).
__global__ void __launch_bounds__(128, ZL_WF_STAGE_MINB) wfLightGenerateKernel(const DScene S, const ZlRenderParams Uin, const WfState W, const long long total,
                                                            const uint32_t seedMul, const int resume) {
    const ZlRenderParams U = wfPassParams(Uin, W);       // graph replays read the pass index from device memory
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = id < total;
    const int slot = (int)id;
    bool toS = false, toE = false;
    if (valid) {
        SamplerState st = makeSampler(S, U, nullptr, 0);                      // uSampler forced to 0 (LightPath.cpp:48-49)
        st.randSeed = resume ? W.smp[slot].x : (uint32_t)U.spp * seedMul + (uint32_t)id + (uint32_t)U.freeCounter;
        const int light = lightSampleOne(S, sample2D(st));
        const float pdfSource = lightPdfSampleOne(S, light);
        const int triId = light + S.objPrimCount;
        const float3 pLit = triangleSampleUniform(S, triId, sample2D(st));
        CameraIiSample ciSamp = thinLensCameraSampleIi(U, pLit, sample2D(st));
        if (ciSamp.pdf > 0) {
            const float3 pCam = pLit + ciSamp.wi * ciSamp.dist;
            const float pdfPos = 1.0f / triangleAreaId(S, triId);
            const float3 Le = lightLe(S, light, pLit, ciSamp.wi);
            const float3 contrib = Le * ciSamp.Ii / (ciSamp.pdf * pdfPos * pdfSource);
            if (!isBlack(contrib) && inFilmBound(ciSamp.uv)) { wfStoreSplat(W, slot, visibleRay(pLit, pCam), ciSamp.uv, contrib); toS = true; }
        }
        LightLeSample leSamp = lightSampleOneLe(S, light, sample4D(st));
        const float3 nl = triangleNg(S, triId, leSamp.ray.ori);
        const float3 wo = -leSamp.ray.dir;
        const Ray ray = rayOffseted(leSamp.ray);                             // offset twice (App. B #4)
        const float3 throughput = leSamp.Le * absDot(nl, -wo) / (pdfSource * leSamp.pdfPos * leSamp.pdfDir);
        W.hit[0][slot] = make_float4(ray.ori.x, ray.ori.y, ray.ori.z, __int_as_float(-1));
        W.dir[slot] = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, 0.0f);
        W.thr[slot] = make_float4(throughput.x, throughput.y, throughput.z, 0.0f);
        W.smp[slot] = make_uint4(st.randSeed, 0u, 0u, 0u);
        toE = U.maxDepth >= 1;
    }
    wfAppend(W.qS, W.cnt + kCntS, toS, slot);
    wfAppend(W.qE, W.cnt + kCntE, toE, slot);
}

// loop body of lightIntegTrace after the bvhHit (light_path_integ.glsl:84-144), one material type per kernel
template <uint32_t TYPE>
__global__ void __launch_bounds__(128, (TYPE == 0u ? ZL_WF_STAGE_MINB : ZL_WF_SHADE_MINB_OTHER)) wfLightShadeKernel(const DScene S, const ZlRenderParams Uin, const WfState W, const int b) {
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
            SurfaceInfo surf = triangleSurfaceInfo(S, id, pos);
            const int matTexId = __ldg(&S.matTex[id]);
            const int matId = matTexId & 0x0000ffff, texId = matTexId >> 16;
            if (TYPE != Dielectric && TYPE != ThinDielectric) {
                if (dot(surf.ns, wo) < 0) { surf.ns = -surf.ns; surf.ng = -surf.ng; }
            }
            const BSDFParam mat = loadMaterial(S, TYPE, matId, texId, surf.uv);
            const float3 ns = surf.ns, ng = surf.ng;
            {
                CameraIiSample ciSamp = thinLensCameraSampleIi(U, pos, sample2D(st));
                if (ciSamp.pdf > 0) {
                    const float3 pCam = pos + ciSamp.wi * ciSamp.dist;
                    const float3 bsdf = materialBSDFT<TYPE>(mat, wo, ciSamp.wi, ns, Importance);
                    const float cosWi = satDot(ng, ciSamp.wi) * fabsf(dot(ns, wo) / dot(ng, wo));
                    const float3 res = ciSamp.Ii * bsdf * throughput * cosWi / ciSamp.pdf;
                    if (!hasNan(res) && !isnan(ciSamp.pdf) && ciSamp.pdf > 1e-8f && !isBlack(res) && inFilmBound(ciSamp.uv)) {
                        wfStoreSplat(W, slot, visibleRay(pos, pCam), ciSamp.uv, res);
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
                const float cosWi = deltaBsdf ? 1.0f : fabsf(dot(ng, wi) * dot(ns, wo) / dot(ng, wo));
                throughput *= samp.bsdf * cosWi / bsdfPdf;
                W.dir[slot] = make_float4(wi.x, wi.y, wi.z, bsdfPdf);
                W.thr[slot] = make_float4(throughput.x, throughput.y, throughput.z, 0.0f);
                toE = true;
            }
            W.smp[slot] = make_uint4(st.randSeed, 0u, 0u, 0u);
        }
        wfAppendRays(S, W, cnt, b, slot, toS, toE, true);
    }
}

}  // namespace zl
