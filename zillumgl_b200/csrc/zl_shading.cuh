// zl_shading.cuh — device shading library: sampler, camera, microfacet + BSDFs, material
// loader, area-light and environment-map sampling.  Restates random.glsl, camera.glsl,
// microfacet.glsl, material.glsl, material_loader.glsl and light.glsl (citations inline);
// expression order is the GLSL's so the CPU oracle and this code agree up to libm.
// Deviations (mirrored by the oracle): alias-table indices are clamped (rand() can be 1.0,
// App. B #13); light triangles read zero texcoords (App. B #24).
#pragma once
#include "zl_traverse.cuh"

namespace zl {

// material.glsl:4-21
enum : uint32_t { Diffuse = 1u << 0, GlosRefl = 1u << 1, GlosTrans = 1u << 2, SpecRefl = 1u << 3, SpecTrans = 1u << 4, Invalid = 1u << 16 };
enum : uint32_t { Lambertian = 0, PrincipledBRDF = 1, MetalWorkflow = 2, Dielectric = 3, ThinDielectric = 4 };
enum : uint32_t { Radiance = 0, Importance = 1 };

struct BSDFParam {                       // material.glsl:23-39
    float3 baseColor; float subsurface;
    float metallic, roughness, specular, specularTint;
    float sheen, sheenTint, clearcoat, clearcoatGloss;
    float ior;
};
struct BSDFSample { float3 wi; float pdf; float3 bsdf; float eta; uint32_t flag; };   // material.glsl:41-48
struct SurfaceInfo { float3 ns, ng; float2 uv; };
struct CameraPdf { float pdfPos, pdfDir; };
struct CameraIiSample { float3 wi; float3 Ii; float dist; float2 uv; float pdf; };
struct LightPdf { float pdfPos, pdfDir; };
struct LightLiSample { float3 wi; float3 coef; float pdf; };
struct LightLeSample { Ray ray; float3 Le; float pdfPos; float pdfDir; };

ZL_DEV BSDFSample makeBSDFSample(float3 wi, float pdf, float3 bsdf, float eta, uint32_t flag) {
    BSDFSample s; s.wi = wi; s.pdf = pdf; s.bsdf = bsdf; s.eta = eta; s.flag = flag; return s;
}
ZL_DEV BSDFSample InvalidBSDFSample() { return makeBSDFSample(f3(0.0f), 0.0f, f3(0.0f), 0.0f, Invalid); }

// -------------------------------------------------------------------------------------------
// Per-invocation sampler state (random.glsl:3,58-59).  The Sobol row of this pass
// (sampleOffset = uSpp * 256, identical for every pixel) is staged in shared memory by the
// kernel; `row` points at it.  Dimensions >= 256 spill into the next sample's row exactly
// like the reference's flat table (App. B #17) through the slow path.
// -------------------------------------------------------------------------------------------
struct SamplerState {
    uint32_t randSeed;
    uint32_t sampleSeed;
    int sampleOffset;            // uSpp * 256 (wrapped), flat index of dimension 0
    int uSampler;                // 0 = hash RNG, 1 = Sobol
    int s;                       // the `Sampler` dimension counter
    const uint32_t* row;         // 256 Sobol words of this pass (shared memory)
    const uint32_t* matrices;    // generator matrices (global) for the spill path
};

ZL_DEV uint32_t sobolSample(const uint32_t* __restrict__ matrices, uint32_t index, int dim) {   // Sampler.cpp:19-28
    uint32_t r = 0;
    for (int i = dim * 32; index != 0; index >>= 1, i++)
        if (index & 1u) r ^= __ldg(matrices + i);
    return r;
}
ZL_DEV float rand(SamplerState& st) {                                                     // random.glsl:15-19
    st.randSeed = hash(st.randSeed);
    return (float)st.randSeed * (1.0f / 4294967296.0f);
}
ZL_DEV float sample1D(SamplerState& st) {                                                 // random.glsl:61-70
    if (st.uSampler == 0) return rand(st);
    uint32_t r;
    if (st.s < 256) r = st.row[st.s];
    else { int flat = st.sampleOffset + st.s; r = sobolSample(st.matrices, (uint32_t)(flat / 256), flat % 256); }
    r ^= st.sampleSeed;
    st.sampleSeed = hash(st.sampleSeed);
    st.s++;
    return (float)r / 4294967296.0f;
}
ZL_DEV float2 sample2D(SamplerState& st) { float a = sample1D(st); float b = sample1D(st); return f2(a, b); }
ZL_DEV float3 sample3D(SamplerState& st) { float a = sample1D(st); float b = sample1D(st); float c = sample1D(st); return f3(a, b, c); }
ZL_DEV float4 sample4D(SamplerState& st) { float a = sample1D(st); float b = sample1D(st); float c = sample1D(st); float d = sample1D(st); return make_float4(a, b, c, d); }

// Per-pixel seeds of the camera-path kernels (path_integ_naive.glsl:156-164)
ZL_DEV void seedPixel(SamplerState& st, const DScene& S, const ZlRenderParams& U, float2 scrCoord) {
    float2 texSize = f2((float)U.filmW, (float)U.filmH);
    float2 noiseCoord = sampleNoise(S, scrCoord);
    noiseCoord = sampleNoise(S, noiseCoord);
    float2 texCoord = texSize * noiseCoord;
    st.randSeed = ((uint32_t)texCoord.x * (uint32_t)U.freeCounter) + (uint32_t)texCoord.y;
    st.sampleSeed = (uint32_t)texCoord.x * (uint32_t)texCoord.y;
}
// sampleOffset of a pass (path_integ_naive.glsl:163-164)
__host__ __device__ inline int passSampleOffset(int spp) {
    const int uSampleDim = 256, uSampleNum = 131072;
    int off = spp * uSampleDim;
    if (off > uSampleNum * uSampleDim) off -= uSampleNum * uSampleDim;
    return off;
}

// -------------------------------------------------------------------------------------------
// Geometry helpers (intersection.glsl:123-224) on the pre-gathered triangle records
// -------------------------------------------------------------------------------------------
struct TriVerts { float3 a, b, c; float tax, tbx, tcx; };
ZL_DEV TriVerts loadTriangle(const DScene& S, int id) {
    const float4* __restrict__ tp = S.triPos + 3 * (size_t)id;
    float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
    TriVerts t; t.a = f3(a); t.b = f3(b); t.c = f3(c); t.tax = a.w; t.tbx = b.w; t.tcx = c.w;
    return t;
}
ZL_DEV float3 triangleSampleUniform(const DScene& S, int id, float2 u) {                 // :123-134
    TriVerts t = loadTriangle(S, id);
    return sampleTriangleUniform(t.a, t.b, t.c, u);
}
ZL_DEV float triangleAreaId(const DScene& S, int id) {                                   // :136-147
    TriVerts t = loadTriangle(S, id);
    return triangleArea(t.a, t.b, t.c);
}
ZL_DEV float3 triangleNormalShad(const DScene& S, int id, float3 p) {                    // :149-173
    TriVerts t = loadTriangle(S, id);
    const float4* __restrict__ tn = S.triNrm + 3 * (size_t)id;
    float3 na = f3(__ldg(tn)), nb = f3(__ldg(tn + 1)), nc = f3(__ldg(tn + 2));
    float3 pa = t.a - p, pb = t.b - p, pc = t.c - p;
    float areaInv = 1.0f / length(cross(t.b - t.a, t.c - t.a));
    float la = length(cross(pb, pc)) * areaInv;
    float lb = length(cross(pc, pa)) * areaInv;
    float lc = 1.0f - la - lb;
    return normalize(na * la + nb * lb + nc * lc);
}
ZL_CALL SurfaceInfo triangleSurfaceInfoCall(const float4* __restrict__ triPos, const float4* __restrict__ triNrm, int id, float3 p) {   // :188-224
    const float4* __restrict__ tp = triPos + 3 * (size_t)id;
    float4 a4 = __ldg(tp), b4 = __ldg(tp + 1), c4 = __ldg(tp + 2);
    TriVerts t; t.a = f3(a4); t.b = f3(b4); t.c = f3(c4); t.tax = a4.w; t.tbx = b4.w; t.tcx = c4.w;
    const float4* __restrict__ tn = triNrm + 3 * (size_t)id;
    float4 n0 = __ldg(tn), n1 = __ldg(tn + 1), n2 = __ldg(tn + 2);
    float3 na = f3(n0), nb = f3(n1), nc = f3(n2);
    float2 ta = f2(t.tax, n0.w), tb = f2(t.tbx, n1.w), tc = f2(t.tcx, n2.w);
    float3 pa = t.a - p, pb = t.b - p, pc = t.c - p;
    float areaInv = 1.0f / length(cross(t.b - t.a, t.c - t.a));
    float la = length(cross(pb, pc)) * areaInv;
    float lb = length(cross(pc, pa)) * areaInv;
    float lc = 1.0f - la - lb;
    SurfaceInfo ret;
    ret.ns = normalize(na * la + nb * lb + nc * lc);
    ret.ng = normalize(cross(pa, pb));
    ret.uv = ta * la + tb * lb + tc * lc;
    if (dot(ret.ns, ret.ng) < 0) ret.ng = -ret.ng;
    return ret;
}
ZL_DEV SurfaceInfo triangleSurfaceInfo(const DScene& S, int id, float3 p) { return triangleSurfaceInfoCall(S.triPos, S.triNrm, id, p); }
// only the geometric normal (what lightLe / lightPdfLi / lightPdfLe use of triangleSurfaceInfo)
ZL_DEV float3 triangleNg(const DScene& S, int id, float3 p) { return triangleSurfaceInfo(S, id, p).ng; }

// -------------------------------------------------------------------------------------------
// camera.glsl
// -------------------------------------------------------------------------------------------
ZL_DEV float3 camF(const ZlRenderParams& U) { return f3(U.camera.F[0], U.camera.F[1], U.camera.F[2]); }
ZL_DEV float3 camR(const ZlRenderParams& U) { return f3(U.camera.R[0], U.camera.R[1], U.camera.R[2]); }
ZL_DEV float3 camU(const ZlRenderParams& U) { return f3(U.camera.U[0], U.camera.U[1], U.camera.U[2]); }
ZL_DEV float3 camPos(const ZlRenderParams& U) { return f3(U.camera.pos[0], U.camera.pos[1], U.camera.pos[2]); }
ZL_DEV bool inFilmBound(float2 uv) { return uv.x >= 0 && uv.x <= 1.0f && uv.y >= 0 && uv.y <= 1.0f; }   // :52-55
ZL_DEV bool thinLensCameraDelta(const ZlRenderParams& U) { return U.camera.lensRadius <= 1e-6f; }     // :57-60
ZL_DEV Ray thinLensCameraSampleRay(const ZlRenderParams& U, float2 uv, float4 u) {                    // :62-77
    float2 texelSize = f2(1.0f / (float)U.filmW, 1.0f / (float)U.filmH);
    float2 biasedCoord = uv + texelSize * f2(u.x, u.y);
    float2 ndc = biasedCoord * 2.0f - 1.0f;
    float2 disk = toConcentricDisk(f2(u.z, u.w)) * U.camera.lensRadius;
    float3 pLens = f3(disk.x, disk.y, 0.0f);
    float2 fp = ndc * f2(U.camera.asp, 1.0f) * U.camera.focalDist * U.camera.tanFOV;
    float3 pFocusPlane = f3(fp.x, fp.y, U.camera.focalDist);
    float3 dir = pFocusPlane - pLens;
    dir = normalize(camR(U) * dir.x + camU(U) * dir.y + camF(U) * dir.z);
    Ray ret;
    ret.ori = camPos(U) + camR(U) * pLens.x + camU(U) * pLens.y;
    ret.dir = dir;
    return ret;
}
ZL_DEV float2 thinLensCameraRasterPos(const ZlRenderParams& U, Ray ray) {                 // :79-91
    float cosTheta = dot(ray.dir, camF(U));
    float dFocus = U.camera.focalDist / cosTheta;
    const float* m = U.camera.matInv;
    Mat3 inv{f3(m[0], m[1], m[2]), f3(m[3], m[4], m[5]), f3(m[6], m[7], m[8])};
    float3 pFocus = inv * (rayPoint(ray, dFocus) - camPos(U));
    float aspect = (float)U.filmW / (float)U.filmH;
    float2 av = f2(aspect, 1.0f) * U.camera.tanFOV;
    pFocus = pFocus / (f3(av.x, av.y, 1.0f) * U.camera.focalDist);
    return (f2(pFocus.x, pFocus.y) + 1.0f) * 0.5f;
}
ZL_DEV float3 thinLensCameraIe(const ZlRenderParams& U, Ray ray) {                        // :93-107
    float cosTheta = dot(ray.dir, camF(U));
    if (cosTheta < 1e-6f) return f3(0.0f);
    float2 pRaster = thinLensCameraRasterPos(U, ray);
    if (!inFilmBound(pRaster)) return f3(0.0f);
    float tanFOVInv = 1.0f / U.camera.tanFOV;
    float cos2Theta = cosTheta * cosTheta;
    float lensArea = thinLensCameraDelta(U) ? 1.0f : Pi * U.camera.lensRadius * U.camera.lensRadius;
    return f3(0.25f) * square(tanFOVInv / cos2Theta) / (lensArea * U.camera.asp);
}
ZL_DEV CameraIiSample thinLensCameraSampleIi(const ZlRenderParams& U, float3 ref, float2 u) {   // :109-127
    CameraIiSample r;
    float2 disk = toConcentricDisk(u) * U.camera.lensRadius;
    float3 pLens = f3(disk.x, disk.y, 0.0f);
    float3 y = camPos(U) + camR(U) * pLens.x + camU(U) * pLens.y + camF(U) * pLens.z;
    float dist = distance(ref, y);
    float3 wi = normalize(y - ref);
    float cosTheta = satDot(camF(U), -wi);
    if (cosTheta < 1e-6f) { r.wi = f3(0.0f); r.Ii = f3(0.0f); r.dist = 0.0f; r.uv = f2(0.0f, 0.0f); r.pdf = 0.0f; return r; }
    Ray ray = makeRay(y, -wi);
    r.Ii = thinLensCameraIe(U, ray);
    r.uv = thinLensCameraRasterPos(U, ray);
    float lensArea = thinLensCameraDelta(U) ? 1.0f : Pi * U.camera.lensRadius * U.camera.lensRadius;
    r.pdf = dist * dist / (cosTheta * lensArea);
    r.wi = wi; r.dist = dist;
    return r;
}
ZL_DEV CameraPdf thinLensCameraPdfIe(const ZlRenderParams& U, Ray ray) {                  // :129-142
    CameraPdf z; z.pdfPos = 0.0f; z.pdfDir = 0.0f;
    float cosTheta = dot(camF(U), ray.dir);
    if (cosTheta < 1e-6f) return z;
    float2 pRaster = thinLensCameraRasterPos(U, ray);
    if (!inFilmBound(pRaster)) return z;
    z.pdfPos = thinLensCameraDelta(U) ? 1.0f : 1.0f / (Pi * U.camera.lensRadius * U.camera.lensRadius);
    z.pdfDir = 1.0f / (cosTheta * cosTheta * cosTheta);
    return z;
}

// -------------------------------------------------------------------------------------------
// microfacet.glsl
// -------------------------------------------------------------------------------------------
ZL_DEV float schlickW(float cosTheta) { return pow5(1.0f - cosTheta); }                   // :4-7
ZL_DEV float3 schlickF(float cosTheta, float3 F0) { return F0 + (f3(1.0f) - F0) * pow5(1.0f - cosTheta); }   // :9-12
ZL_DEV float3 schlickF(float cosTheta, float3 F0, float roughness) {                      // :14-17
    return F0 + (gmax(f3(1.0f - roughness), F0) - F0) * pow5(1.0f - cosTheta);
}
ZL_DEV float schlickG(float cosTheta, float alpha) {                                      // :19-23
    float k = alpha * 0.5f;
    return cosTheta / (cosTheta * (1.0f - k) + k);
}
ZL_DEV float smithG(float3 n, float3 wo, float3 wi, float alpha) {                        // :25-28
    return schlickG(absDot(n, wo), alpha) * schlickG(absDot(n, wi), alpha);
}
ZL_DEV float ggx(float cosTheta, float alpha) {                                           // :30-40
    if (cosTheta < 1e-6f) return 0.0f;
    float a2 = alpha * alpha;
    float denom = cosTheta * cosTheta * (a2 - 1.0f) + 1.0f;
    denom = denom * denom * Pi;
    return a2 / denom;
}
ZL_DEV float ggxD(float3 n, float3 m, float alpha) { return ggx(dot(n, m), alpha); }      // :42-45
ZL_DEV float ggxPdfWm(float3 n, float3 m, float3 wo, float alpha) { return ggx(dot(n, m), alpha); }   // :47-50
ZL_DEV float ggxPdfVisibleWm(float3 n, float3 m, float3 wo, float alpha) {                // :52-55
    return ggx(dot(n, m), alpha) * schlickG(dot(n, wo), alpha) * absDot(m, wo) / absDot(n, wo);
}
ZL_CALL float3 ggxSampleWm(float3 n, float3 wo, float alpha, float2 u) {                   // :57-64
    float2 xi = toConcentricDisk(u);
    float3 h = f3(xi.x, xi.y, sqrtf(gmax(0.0f, 1.0f - xi.x * xi.x - xi.y * xi.y)));
    h = normalize(h * f3(alpha, alpha, 1.0f));
    return normalToWorld(n, h);
}
ZL_CALL float3 ggxSampleVisibleWm(float3 n, float3 wo, float alpha, float2 u) {            // :74-92
    Mat3 tbn = tbnMatrix(n);
    Mat3 tbnInv = inverse(tbn);
    float3 vh = normalize((tbnInv * wo) * f3(alpha, alpha, 1.0f));
    float lensq = vh.x * vh.x + vh.y * vh.y;
    float3 t1 = lensq > 0.0f ? f3(-vh.y, vh.x, 0.0f) / sqrtf(lensq) : f3(1.0f, 0.0f, 0.0f);
    float3 t2 = cross(vh, t1);
    float2 xi = toConcentricDisk(u);
    float s = 0.5f * (1.0f + vh.z);
    xi.y = (1.0f - s) * sqrtf(1.0f - xi.x * xi.x) + s * xi.y;
    float3 h = t1 * xi.x + t2 * xi.y + vh * sqrtf(gmax(0.0f, 1.0f - xi.x * xi.x - xi.y * xi.y));
    h = normalize(f3(h.x * alpha, h.y * alpha, gmax(0.0f, h.z)));
    return normalToWorld(n, h);
}
ZL_DEV float gtr1(float cosTheta, float alpha) {                                          // :94-98
    float a2 = alpha * alpha;
    return (a2 - 1.0f) / (2.0f * Pi * zl_logf(alpha) * (1.0f + (a2 - 1.0f) * cosTheta * cosTheta));
}
ZL_DEV float gtr1D(float3 n, float3 m, float alpha) { return gtr1(satDot(n, m), alpha); } // :100-103
ZL_CALL float3 gtr1SampleWm(float3 n, float3 wo, float alpha, float2 u) {                  // :105-115
    float cosTheta = sqrtf(gmax(0.0f, (1.0f - zl_powf(alpha, 1.0f - u.x)) / (1.0f - alpha)));
    float sinTheta = sqrtf(gmax(0.0f, 1.0f - cosTheta * cosTheta));
    float phi = 2.0f * u.y * Pi;
    float sp, cp;
    zl_sincosf(phi, &sp, &cp);
    float3 m = normalize(f3(cp * sinTheta, sp * sinTheta, cosTheta));
    if (!sameHemisphere(n, wo, m)) m = -m;
    return normalize(normalToWorld(n, m));
}
ZL_DEV float gtr1PdfWm(float3 n, float3 m, float3 wo, float alpha) { return gtr1D(n, m, alpha) * absDot(n, m); }   // :117-120

// -------------------------------------------------------------------------------------------
// material.glsl
// -------------------------------------------------------------------------------------------
ZL_DEV bool approximateDelta(float roughness) { return roughness < 0.02f; }               // :63-66
ZL_DEV float3 lambertian(const BSDFParam& p) { return p.baseColor * PiInv; }              // :68-71
ZL_DEV float lambertianPdf(float3 wi, float3 n) { return satDot(wi, n) * PiInv; }         // :73-76
ZL_DEV BSDFSample lambertianSample(float3 n, const BSDFParam& p, float3 u) {              // :78-83
    float4 s = sampleCosineWeighted(n, f2(u.y, u.z));
    float3 wi = f3(s);
    float pdf = satDot(n, wi) * PiInv;
    return makeBSDFSample(wi, pdf, p.baseColor * PiInv, 1.0f, Diffuse);
}
ZL_CALL float3 metalWorkflow(float3 wo, float3 wi, float3 n, const BSDFParam& param) {     // :85-114
    float3 baseColor = param.baseColor;
    float metallic = param.metallic, roughness = param.roughness;
    float alpha = square(roughness);
    float3 h = normalize(wi + wo);
    if (!sameHemisphere(n, wo, wi)) return f3(0.0f);
    float cosWi = dot(n, wi), cosWo = dot(n, wo);
    float3 f0 = mix(f3(0.04f), baseColor, metallic);
    float3 f = schlickF(satDot(h, wo), f0, roughness);
    float d = ggxD(n, h, alpha);
    float g = smithG(n, wo, wi, alpha);
    float3 kd = f3(1.0f) - f;
    kd *= 1.0f - metallic;
    float denom = 4.0f * cosWo * cosWi;
    if (denom < 1e-7f) return f3(0.0f);
    return kd * baseColor * PiInv + f * d * g / denom;
}
ZL_CALL float metalWorkflowPdf(float3 wo, float3 wi, float3 n, const BSDFParam& param) {   // :116-124
    float alpha = square(param.roughness);
    float3 h = normalize(wo + wi);
    float pdfDiff = satDot(n, wi) * PiInv;
    float pdfSpec = ggxPdfVisibleWm(n, h, wo, alpha) / (4.0f * absDot(h, wo));
    float spec = 1.0f / (2.0f - param.metallic);
    return mix(pdfDiff, pdfSpec, spec);
}
ZL_CALL BSDFSample metalWorkflowSample(float3 n, float3 wo, const BSDFParam& param, float3 u) {   // :126-149
    float alpha = square(param.roughness);
    float spec = 1.0f / (2.0f - param.metallic);
    uint32_t type = u.x > spec ? Diffuse : GlosRefl;
    float3 wi;
    if (type == Diffuse) wi = f3(sampleCosineWeighted(n, f2(u.y, u.z)));
    else {
        float3 h = ggxSampleVisibleWm(n, wo, alpha, f2(u.y, u.z));
        wi = reflect(-wo, h);
    }
    float cosWi = dot(n, wi);
    if (cosWi < 0) return InvalidBSDFSample();
    float3 bsdf = metalWorkflow(wo, wi, n, param);
    float pdf = metalWorkflowPdf(wo, wi, n, param);
    return makeBSDFSample(wi, pdf, bsdf, 1.0f, type);
}
ZL_DEV bool refract(float3& wt, float3 wi, float3 n, float eta) {                         // :151-165
    float cosTi = dot(n, wi);
    if (cosTi < 0) eta = 1.0f / eta;
    float sin2Ti = gmax(0.0f, 1.0f - cosTi * cosTi);
    float sin2Tt = sin2Ti / (eta * eta);
    if (sin2Tt >= 1.0f) return false;
    float cosTt = sqrtf(1.0f - sin2Tt);
    if (cosTi < 0) cosTt = -cosTt;
    wt = normalize(-wi / eta + n * (cosTi / eta - cosTt));
    return true;
}
ZL_DEV float fresnelDielectric(float cosTi, float eta) {                                  // :167-186
    cosTi = gclamp(cosTi, -1.0f, 1.0f);
    if (cosTi < 0.0f) { eta = 1.0f / eta; cosTi = -cosTi; }
    float sinTi = sqrtf(1.0f - cosTi * cosTi);
    float sinTt = sinTi / eta;
    if (sinTt >= 1.0f) return 1.0f;
    float cosTt = sqrtf(1.0f - sinTt * sinTt);
    float rPa = (cosTi - eta * cosTt) / (cosTi + eta * cosTt);
    float rPe = (eta * cosTi - cosTt) / (eta * cosTi + cosTt);
    return (rPa * rPa + rPe * rPe) * 0.5f;
}
ZL_CALL float3 dielectric(float3 wo, float3 wi, float3 n, const BSDFParam& param, uint32_t mode) {   // :188-223
    float3 baseColor = param.baseColor;
    float roughness = param.roughness, ior = param.ior;
    if (approximateDelta(roughness)) return f3(0.0f);
    float3 h = normalize(wo + wi);
    float hCosWo = absDot(h, wo), hCosWi = absDot(h, wi);
    float alpha = roughness * roughness;
    if (sameHemisphere(n, wo, wi)) {
        float refl = fresnelDielectric(absDot(h, wi), ior);
        return (hCosWo * hCosWi < 1e-7f) ? f3(0.0f)
             : baseColor * ggxD(n, h, alpha) * smithG(n, wo, wi, alpha) / (4.0f * hCosWo * hCosWi) * refl;
    } else {
        float eta = dot(n, wi) > 0 ? ior : 1.0f / ior;
        float sqrtDenom = dot(h, wo) + eta * dot(h, wi);
        float denom = sqrtDenom * sqrtDenom;
        denom *= absDot(n, wi) * absDot(n, wo);
        float refl = fresnelDielectric(dot(h, wi), eta);
        float factor = (mode == Radiance) ? square(1.0f / eta) : 1.0f;
        return (denom < 1e-7f) ? f3(0.0f)
             : baseColor * fabsf(ggxD(n, h, alpha) * smithG(n, wo, wi, alpha) * hCosWo * hCosWi) / denom * (1.0f - refl) * factor;
    }
}
ZL_CALL float dielectricPdf(float3 wo, float3 wi, float3 n, const BSDFParam& param) {      // :225-252
    float roughness = param.roughness, ior = param.ior;
    if (approximateDelta(roughness)) return 0.0f;
    if (sameHemisphere(n, wo, wi)) {
        float3 h = normalize(wo + wi);
        if (dot(wo, h) < 0.0f) return 0.0f;
        float refl = fresnelDielectric(absDot(h, wi), ior);
        return ggxPdfWm(n, h, wo, roughness * roughness) / (4.0f * absDot(h, wo)) * refl;
    } else {
        float eta = dot(n, wo) > 0 ? ior : 1.0f / ior;
        float3 h = normalize(wo + wi * eta);
        if (sameHemisphere(h, wo, wi)) return 0.0f;
        float trans = 1.0f - fresnelDielectric(absDot(h, wo), eta);
        float dHdWi = absDot(h, wi) / square(dot(h, wo) + eta * dot(h, wi));
        return ggxPdfWm(n, h, wo, roughness * roughness) * dHdWi * trans;
    }
}
ZL_CALL BSDFSample dielectricSample(float3 n, float3 wo, const BSDFParam& param, uint32_t mode, float3 u) {   // :254-338
    float3 baseColor = param.baseColor;
    float roughness = param.roughness, ior = param.ior;
    if (approximateDelta(roughness)) {
        float refl = fresnelDielectric(dot(n, wo), ior);
        if (u.x < refl) return makeBSDFSample(reflect(-wo, n), 1.0f, baseColor, 1.0f, SpecRefl);
        float3 wi;
        if (!refract(wi, wo, n, ior)) return InvalidBSDFSample();
        if (dot(n, wo) < 0) ior = 1.0f / ior;
        float factor = (mode == Radiance) ? square(1.0f / ior) : 1.0f;
        return makeBSDFSample(wi, 1.0f, baseColor * factor, ior, SpecTrans);
    }
    float alpha = roughness * roughness;
    float3 h = ggxSampleWm(n, wo, alpha, f2(u.y, u.z));
    if (dot(n, h) < 0.0f) h = -h;
    float refl = fresnelDielectric(dot(h, wo), ior);
    if (u.x < refl) {
        float3 wi = -reflect(wo, h);
        if (!sameHemisphere(n, wo, wi)) return InvalidBSDFSample();
        float p = ggxPdfWm(n, h, wo, alpha) / (4.0f * absDot(h, wo));
        float hCosWo = absDot(h, wo), hCosWi = absDot(h, wi);
        float3 r = (hCosWo * hCosWi < 1e-7f) ? f3(0.0f)
                 : baseColor * ggxD(n, h, alpha) * smithG(n, wo, wi, alpha) / (4.0f * hCosWo * hCosWi);
        if (isnan(p)) p = 0.0f;
        return makeBSDFSample(wi, p, r, 1.0f, GlosRefl);
    }
    float3 wi;
    if (!refract(wi, wo, h, ior)) return InvalidBSDFSample();
    if (sameHemisphere(n, wo, wi)) return InvalidBSDFSample();
    if (absDot(n, wi) < 1e-10f) return InvalidBSDFSample();
    float hCosWo = absDot(h, wo), hCosWi = absDot(h, wi);
    if (dot(h, wo) < 0) ior = 1.0f / ior;
    float sqrtDenom = dot(h, wo) + ior * dot(h, wi);
    float denom = sqrtDenom * sqrtDenom;
    float dHdWi = hCosWi / denom;
    float factor = (mode == Radiance) ? square(1.0f / ior) : 1.0f;
    denom *= absDot(n, wi) * absDot(n, wo);
    float3 t = (denom < 1e-7f) ? f3(0.0f)
             : baseColor * fabsf(ggxD(n, h, alpha) * smithG(n, wo, wi, alpha) * hCosWo * hCosWi) / denom;
    float p = ggxPdfWm(n, h, wo, alpha) * dHdWi;
    if (isnan(p)) p = 0.0f;
    return makeBSDFSample(wi, p, t * factor, ior, GlosTrans);
}
ZL_DEV BSDFSample thinDielectricSample(float3 n, float3 wo, const BSDFParam& param, float3 u) {   // :340-355
    if (dot(n, wo) < 0) n = -n;
    float refl = fresnelDielectric(dot(n, wo), param.ior);
    float trans = 1.0f - refl;
    if (refl < 1.0f) {
        refl += trans * trans * refl / (1.0f - refl * refl);
        trans = 1.0f - refl;
    }
    return (u.x < refl) ? makeBSDFSample(reflect(-wo, n), 1.0f, param.baseColor, 1.0f, SpecRefl)
                        : makeBSDFSample(-wo, 1.0f, param.baseColor, 1.0f, SpecTrans);
}
ZL_DEV float3 principledMetal(float3 wo, float3 wi, float3 n, float3 fm0, float alpha) {  // :357-373
    float cosWo = satDot(n, wo), cosWi = satDot(n, wi);
    float3 h = normalize(wo + wi);
    if (cosWo < 1e-10f || cosWi < 1e-10f) return f3(0.0f);
    float3 fm = schlickF(absDot(h, wo), fm0);
    float dm = ggxD(n, h, alpha);
    float gm = smithG(n, wo, wi, alpha);
    float denom = 4.0f * cosWi * cosWo;
    if (denom < 1e-7f) return f3(0.0f);
    return fm * dm * gm / denom;
}
ZL_DEV float principledMetalPdf(float3 wo, float3 wi, float3 n, float alpha) {            // :375-379
    float3 h = normalize(wo + wi);
    return ggxPdfVisibleWm(n, h, wo, alpha) / (4.0f * absDot(h, wo));
}
ZL_DEV float3 principledClearcoat(float3 wo, float3 wi, float3 n, float3 baseColor, float alpha) {   // :393-409
    float cosWo = satDot(n, wo), cosWi = satDot(n, wi);
    float3 h = normalize(wo + wi);
    if (cosWo < 1e-6f || cosWi < 1e-6f) return f3(0.0f);
    float3 fc = schlickF(absDot(h, wo), baseColor);
    float dc = gtr1D(n, h, alpha);
    float gc = smithG(n, wo, wi, 0.25f);
    float denom = 4.0f * cosWi * cosWo;
    if (denom < 1e-7f) return f3(0.0f);
    return fc * dc * gc / denom;
}
ZL_DEV float principledClearcoatPdf(float3 wo, float3 wi, float3 n, float alpha) {        // :411-415
    float3 h = normalize(wo + wi);
    return gtr1PdfWm(n, h, wo, alpha) / (4.0f * absDot(h, wo));
}
ZL_DEV float3 principledDiffuse(float3 wo, float3 wi, float3 n, float3 baseColor, float subsurface, float roughness) {   // :429-452
    float cosWo = satDot(n, wo), cosWi = satDot(n, wi);
    if (cosWo < 1e-10f || cosWi < 1e-10f) return f3(0.0f);
    float3 h = normalize(wo + wi);
    float hCosWi = dot(h, wi);
    float hCosWi2 = hCosWi * hCosWi;
    float fi = schlickW(cosWi), fo = schlickW(cosWo);
    float3 fd90 = f3(0.5f + 2.0f * roughness * hCosWi2);
    float3 fd = mix(f3(1.0f), fd90, fi) * mix(f3(1.0f), fd90, fo);
    float3 baseDiffuse = baseColor * fd * PiInv;
    float3 fss90 = f3(roughness * hCosWi2);
    float3 fss = mix(f3(1.0f), fss90, fi) * mix(f3(1.0f), fss90, fo);
    float3 ss = baseColor * PiInv * 1.25f * (fss * (1.0f / (cosWi + cosWo) - 0.5f) + f3(0.5f));
    return mix(baseDiffuse, ss, subsurface);
}
ZL_DEV float3 principledFm0(const BSDFParam& param) {
    float lum = luminance(param.baseColor);
    float3 tintColor = lum > 0 ? param.baseColor / lum : f3(1.0f);
    return mix(0.08f * param.specular * mix(f3(1.0f), tintColor, param.specularTint), param.baseColor, param.metallic);
}
ZL_CALL float3 principledBRDF(float3 wo, float3 wi, float3 n, const BSDFParam& param) {    // :462-490
    float3 res = f3(0.0f);
    float3 baseColor = param.baseColor;
    float alpha = square(param.roughness);
    float clearcoatAlpha = mix(0.1f, 0.001f, param.clearcoatGloss);
    float lum = luminance(baseColor);
    float3 tintColor = lum > 0 ? baseColor / lum : f3(1.0f);
    float3 fm0 = mix(0.08f * param.specular * mix(f3(1.0f), tintColor, param.specularTint), baseColor, param.metallic);
    float hCosWi = dot(normalize(wo + wi), wi);
    res += principledDiffuse(wo, wi, n, baseColor, param.subsurface, param.roughness) * (1.0f - param.metallic);
    res += principledMetal(wo, wi, n, fm0, alpha);
    res += principledClearcoat(wo, wi, n, baseColor, clearcoatAlpha) * param.clearcoat * 0.25f;
    res += mix(f3(1.0f), tintColor, param.sheenTint) * schlickW(hCosWi) * param.sheen * (dot(n, wi) < 0.0f ? 0.0f : 1.0f);
    return res;
}
ZL_CALL float principledBRDFPdf(float3 wo, float3 wi, float3 n, const BSDFParam& param) {  // :492-514
    float pdf = 0.0f;
    float alpha = square(param.roughness);
    float clearcoatAlpha = mix(0.1f, 0.001f, param.clearcoatGloss);
    float spec = 1.0f / (2.0f - param.metallic);
    float cosinePdf = absDot(n, wi) * PiInv;
    pdf += cosinePdf * (1.0f - spec);
    pdf += principledMetalPdf(wo, wi, n, alpha) * spec;
    pdf += principledClearcoatPdf(wo, wi, n, clearcoatAlpha) * 0.25f * param.clearcoat;
    return pdf / (1.0f + 0.25f * param.clearcoat);
}
// material.glsl:516-555.  Only the sampled direction of the chosen lobe is used; the lobe is
// picked with the hash RNG even in Sobol mode, clearcoat is sampled with the base alpha and the
// returned flag is always Diffuse (App. B #7).
ZL_CALL BSDFSample principledBRDFSample(float3 n, float3 wo, const BSDFParam& param, float3 u, SamplerState& st) {
    float alpha = square(param.roughness);
    float spec = 1.0f / (2.0f - param.metallic);
    float cdf0 = 1.0f - spec, cdf1 = 1.0f, cdf2 = 1.0f + param.clearcoat * 0.25f;
    float3 wi;
    float s = rand(st) * cdf2;
    if (s <= cdf0) wi = f3(sampleCosineWeighted(n, f2(u.y, u.z)));
    else if (s <= cdf1) {
        float3 h = ggxSampleVisibleWm(n, wo, alpha, f2(u.y, u.z));
        wi = reflect(-wo, h);
        if (dot(n, wi) < 0.0f) wi = f3(0.0f);          // InvalidBSDFSample.wi (material.glsl:386-387)
    } else {
        float3 h = gtr1SampleWm(n, wo, alpha, f2(u.y, u.z));
        wi = reflect(-wo, h);
        if (dot(n, wi) < 0.0f) wi = f3(0.0f);          // material.glsl:422-423
    }
    float3 bsdf = principledBRDF(wo, wi, n, param);
    float pdf = principledBRDFPdf(wo, wi, n, param);
    return makeBSDFSample(wi, pdf, bsdf, 1.0f, Diffuse);
}

// -------------------------------------------------------------------------------------------
// material_loader.glsl
// -------------------------------------------------------------------------------------------
ZL_DEV uint32_t loadMaterialType(const DScene& S, int matId) {                            // :3-6
    return (uint32_t)__float_as_int(__ldg(&S.materials[matId * 4 + 3]).y);
}
ZL_DEV BSDFParam loadMaterial(const DScene& S, uint32_t matType, int matId, int texId, float2 uv) {   // :8-97
    BSDFParam ret;
    ret.subsurface = 0; ret.metallic = 0; ret.roughness = 0; ret.specular = 0; ret.specularTint = 0;
    ret.sheen = 0; ret.sheenTint = 0; ret.clearcoat = 0; ret.clearcoatGloss = 0; ret.ior = 0;
    float4 baseRou = __ldg(&S.materials[matId * 4 + 0]);
    if (texId == -1) ret.baseColor = f3(baseRou);
    else {
        float2 uvScale = __ldg(&S.texScale[texId]);
        ret.baseColor = sampleAlbedo(S, f2(fract(uv.x), fract(uv.y)) * uvScale, texId);
    }
    switch (matType) {
    case PrincipledBRDF: {
        float4 a = __ldg(&S.materials[matId * 4 + 1]), b = __ldg(&S.materials[matId * 4 + 2]);
        ret.roughness = mix(0.0134f, 1.0f, baseRou.w);
        ret.subsurface = a.x; ret.metallic = a.y; ret.specular = a.z; ret.specularTint = a.w;
        ret.sheen = b.x; ret.sheenTint = b.y; ret.clearcoat = b.z; ret.clearcoatGloss = b.w;
        break; }
    case MetalWorkflow:
        ret.roughness = mix(0.0134f, 1.0f, baseRou.w);
        ret.metallic = __ldg(&S.materials[matId * 4 + 1]).y;
        break;
    case Dielectric:
        ret.roughness = baseRou.w;
        ret.ior = __ldg(&S.materials[matId * 4 + 3]).x;
        break;
    case ThinDielectric:
        ret.ior = __ldg(&S.materials[matId * 4 + 3]).x;
        break;
    default: break;
    }
    return ret;
}
ZL_CALL float3 materialBSDF(uint32_t matType, const BSDFParam& p, float3 wo, float3 wi, float3 n, uint32_t mode) {   // :99-115
    switch (matType) {
    case PrincipledBRDF: return principledBRDF(wo, wi, n, p);
    case MetalWorkflow: return metalWorkflow(wo, wi, n, p);
    case Dielectric: return dielectric(wo, wi, n, p, mode);
    case ThinDielectric: return f3(0.0f);
    default: return lambertian(p);
    }
}
ZL_CALL float materialPdf(uint32_t matType, const BSDFParam& p, float3 wo, float3 wi, float3 n, uint32_t mode) {    // :135-151
    switch (matType) {
    case PrincipledBRDF: return principledBRDFPdf(wo, wi, n, p);
    case MetalWorkflow: return metalWorkflowPdf(wo, wi, n, p);
    case Dielectric: return dielectricPdf(wo, wi, n, p);
    case ThinDielectric: return 0.0f;
    default: return lambertianPdf(wi, n);
    }
}
ZL_DEV float4 materialBSDFAndPdf(uint32_t matType, const BSDFParam& p, float3 wo, float3 wi, float3 n, uint32_t mode) {   // :117-133
    if (matType == ThinDielectric) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float3 b = materialBSDF(matType, p, wo, wi, n, mode);
    return make_float4(b.x, b.y, b.z, materialPdf(matType, p, wo, wi, n, mode));
}
ZL_CALL BSDFSample materialSample(uint32_t matType, const BSDFParam& p, float3 n, float3 wo, uint32_t mode, float3 u, SamplerState& st) {   // :153-169
    switch (matType) {
    case PrincipledBRDF: return principledBRDFSample(n, wo, p, u, st);
    case MetalWorkflow: return metalWorkflowSample(n, wo, p, u);
    case Dielectric: return dielectricSample(n, wo, p, mode, u);
    case ThinDielectric: return thinDielectricSample(n, wo, p, u);
    default: return lambertianSample(n, p, u);
    }
}

// Compile-time material type (the wavefront variant shades one material type per kernel, so only
// that type's BSDF code is reachable: small instruction footprint, no 5-way divergence).  TYPE 0
// stands for Lambertian AND every unknown type value, like the `default:` labels above.
template <uint32_t TYPE>
ZL_DEV float4 materialBSDFAndPdfT(const BSDFParam& p, float3 wo, float3 wi, float3 n, uint32_t mode) {
    if (TYPE == ThinDielectric) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float3 b; float pdf;
    if (TYPE == PrincipledBRDF) { b = principledBRDF(wo, wi, n, p); pdf = principledBRDFPdf(wo, wi, n, p); }
    else if (TYPE == MetalWorkflow) { b = metalWorkflow(wo, wi, n, p); pdf = metalWorkflowPdf(wo, wi, n, p); }
    else if (TYPE == Dielectric) { b = dielectric(wo, wi, n, p, mode); pdf = dielectricPdf(wo, wi, n, p); }
    else { b = lambertian(p); pdf = lambertianPdf(wi, n); }
    return make_float4(b.x, b.y, b.z, pdf);
}
template <uint32_t TYPE>
ZL_DEV float3 materialBSDFT(const BSDFParam& p, float3 wo, float3 wi, float3 n, uint32_t mode) {
    if (TYPE == ThinDielectric) return f3(0.0f);
    if (TYPE == PrincipledBRDF) return principledBRDF(wo, wi, n, p);
    if (TYPE == MetalWorkflow) return metalWorkflow(wo, wi, n, p);
    if (TYPE == Dielectric) return dielectric(wo, wi, n, p, mode);
    return lambertian(p);
}
template <uint32_t TYPE>
ZL_DEV float materialPdfT(const BSDFParam& p, float3 wo, float3 wi, float3 n, uint32_t mode) {
    if (TYPE == ThinDielectric) return 0.0f;
    if (TYPE == PrincipledBRDF) return principledBRDFPdf(wo, wi, n, p);
    if (TYPE == MetalWorkflow) return metalWorkflowPdf(wo, wi, n, p);
    if (TYPE == Dielectric) return dielectricPdf(wo, wi, n, p);
    return lambertianPdf(wi, n);
}
template <uint32_t TYPE>
ZL_DEV BSDFSample materialSampleT(const BSDFParam& p, float3 n, float3 wo, uint32_t mode, float3 u, SamplerState& st) {
    if (TYPE == PrincipledBRDF) return principledBRDFSample(n, wo, p, u, st);
    if (TYPE == MetalWorkflow) return metalWorkflowSample(n, wo, p, u);
    if (TYPE == Dielectric) return dielectricSample(n, wo, p, mode, u);
    if (TYPE == ThinDielectric) return thinDielectricSample(n, wo, p, u);
    return lambertianSample(n, p, u);
}
// queue bin of a material type value: 1..4 as they are, everything else shades as Lambertian
ZL_DEV int materialBin(uint32_t matType) { return (matType >= 1u && matType <= 4u) ? (int)matType : 0; }

// -------------------------------------------------------------------------------------------
// light.glsl
// -------------------------------------------------------------------------------------------
ZL_DEV int lightSampleOne(const DScene& S, float2 u, float4& powProbOut) {                // :68-72
    int cx = (int)((float)S.numLightTriangles * u.x);
    if (cx > S.numLightTriangles - 1) cx = S.numLightTriangles - 1;
    float4 pp = __ldg(&S.lightPowProb[cx]);
    return (u.y < pp.w) ? cx : __ldg(&S.lightAlias[cx]);
}
ZL_DEV int lightSampleOne(const DScene& S, float2 u) { float4 t; return lightSampleOne(S, u, t); }
ZL_DEV float3 lightPower(const DScene& S, int id) { return f3(__ldg(&S.lightPowProb[id])); }
ZL_DEV float lightPdfSampleOne(const DScene& S, int id) { return luminance(lightPower(S, id)) / S.lightSum; }   // :74-77
ZL_DEV float3 lightLe(const DScene& S, int id, float3 x, float3 wo) {                     // :79-86
    int triId = id + S.objPrimCount;
    float3 ng = triangleNg(S, triId, x);
    if (dot(wo, ng) <= 0.0f) return f3(0.0f);
    return lightPower(S, id) / triangleAreaId(S, triId) * 0.5f * PiInv;
}
ZL_DEV float lightPdfLi(const DScene& S, int id, float3 x, float3 y) {                    // :88-98
    int triId = id + S.objPrimCount;
    float3 norm = triangleNg(S, triId, y);
    float3 yx = normalize(x - y);
    float cosTheta = absDot(norm, yx);
    if (cosTheta < 1e-8f) return -1.0f;
    return distSquare(x, y) / (triangleAreaId(S, triId) * cosTheta);
}
ZL_DEV LightPdf lightPdfLe(const DScene& S, int id, Ray ray) {                            // :100-109
    LightPdf ret;
    int triId = id + S.objPrimCount;
    float3 norm = triangleNg(S, triId, ray.ori);
    ret.pdfPos = 1.0f / triangleAreaId(S, triId);
    ret.pdfDir = (dot(norm, ray.dir) <= 0) ? 0.0f : 0.5f * PiInv;
    return ret;
}
ZL_DEV LightLeSample lightSampleOneLe(const DScene& S, int id, float4 u) {                // :111-120
    int triId = id + S.objPrimCount;
    float3 ori = triangleSampleUniform(S, triId, f2(u.x, u.y));
    float3 norm = triangleNg(S, triId, ori);
    float4 samp = sampleCosineWeighted(norm, f2(u.z, u.w));
    LightLeSample r;
    r.ray = rayOffseted(makeRay(ori, f3(samp)));
    r.Le = lightLe(S, id, ori, f3(samp));
    r.pdfPos = 1.0f / triangleAreaId(S, triId);
    r.pdfDir = samp.w;
    return r;
}
ZL_DEV LightLiSample invalidLiSample() { LightLiSample r; r.wi = f3(0.0f); r.coef = f3(0.0f); r.pdf = 0.0f; return r; }
// Visibility policy of the NEE samplers.  The megakernel traces the shadow ray where the GLSL does
// (ImmediateVis); the wavefront variant records it (DeferredVis), computes the sample as if visible and
// drops the contribution later if the queued shadow ray is occluded.  No random number is drawn after the
// test in light.glsl:122-155 / :207-219, so both orders give the same sample stream and the same values.
struct ImmediateVis {
    ZL_DEV bool occluded(const DScene& S, Ray ray, float dist) { return bvhTest(S, ray, dist); }
};
struct DeferredVis {
    Ray ray; float dist; bool pending;
    ZL_DEV bool occluded(const DScene&, Ray r, float d) { ray = r; dist = d; pending = true; return false; }
};
template <class Vis>
ZL_DEV LightLiSample lightSampleLi(const DScene& S, int id, float3 x, float2 u, Vis& vis) {         // :122-155
    int triId = id + S.objPrimCount;
    TriVerts t = loadTriangle(S, triId);
    float3 y = sampleTriangleUniform(t.a, t.b, t.c, u);
    float3 wi = normalize(y - x);
    float3 norm = triangleNg(S, triId, y);
    float cosTheta = dot(norm, -wi);
    if (cosTheta < 1e-6f) return invalidLiSample();
    Ray lightRay = rayOffseted(x, wi);
    float dist = distance(x, y);
    float pdf = dist * dist / (triangleArea(t.a, t.b, t.c) * cosTheta);
    float testDist = dist - 1e-4f - 1e-6f;
    if (vis.occluded(S, lightRay, testDist) || pdf < 1e-8f) return invalidLiSample();
    float3 weight = lightLe(S, id, y, -wi);
    float pdfSample = luminance(lightPower(S, id)) / S.lightSum;
    pdf *= pdfSample;
    LightLiSample r; r.wi = wi; r.coef = weight / pdf; r.pdf = pdf;
    return r;
}
ZL_CALL float3 envLeCall(const ushort4* __restrict__ env, int envW, int envH, float envRotation, float3 wi) {   // :163-167
    DScene S;
    S.env = env; S.envW = envW; S.envH = envH;
    wi = rotateZ(wi, -envRotation);
    return sampleEnv(S, sphereToPlane(wi));
}
ZL_DEV float3 envLe(const DScene& S, const ZlRenderParams& U, float3 wi) { return envLeCall(S.env, S.envW, S.envH, U.envRotation, wi); }
ZL_DEV float envGetPortion(const DScene& S, const ZlRenderParams& U, float3 wi) { return luminance(envLe(S, U, wi)) / S.envSum; }   // :169-172
ZL_DEV float envPdfLi(const DScene& S, const ZlRenderParams& U, float3 wi) {              // :174-179
    if (S.envSum == 0.0f) return 0.0f;
    return envGetPortion(S, U, wi) * (float)S.envW * (float)S.envH * 0.5f * square(PiInv);
}
ZL_DEV float4 envSampleWi(const DScene& S, const ZlRenderParams& U, float4 u) {           // :181-205
    int w = S.envW, h = S.envH;
    int rx = (int)((float)h * u.x);
    if (rx > h - 1) rx = h - 1;
    int2 re = __ldg(&S.envAlias[(size_t)rx * (w + 1) + w]);
    int row = (u.y < __int_as_float(re.y)) ? rx : re.x;
    int cx = (int)((float)w * u.z);
    if (cx > w - 1) cx = w - 1;
    int2 ce = __ldg(&S.envAlias[(size_t)row * (w + 1) + cx]);
    int col = (u.w < __int_as_float(ce.y)) ? cx : ce.x;
    float2 uv = f2((float)col + 0.5f, (float)row + 0.5f) / f2((float)w, (float)h);
    float3 wi = planeToSphere(uv);
    wi = rotateZ(wi, U.envRotation);
    float pdf = envGetPortion(S, U, wi) * (float)w * (float)h * 0.5f * square(PiInv);
    return make_float4(wi.x, wi.y, wi.z, pdf);
}
template <class Vis>
ZL_DEV LightLiSample envSampleLi(const DScene& S, const ZlRenderParams& U, float3 x, float4 u, Vis& vis) {   // :207-219
    float4 sp = envSampleWi(S, U, u);
    float3 wi = f3(sp);
    float pdf = sp.w;
    Ray ray = rayOffseted(x, wi);
    if (vis.occluded(S, ray, 1e8f) || pdf == 0.0f) return invalidLiSample();
    LightLiSample r; r.wi = wi; r.coef = envLe(S, U, wi) / pdf; r.pdf = pdf;
    return r;
}
template <class Vis>
ZL_DEV LightLiSample sampleLightAndEnv(const DScene& S, const ZlRenderParams& U, float3 x, float ud, float4 us, Vis& vis) {   // :221-235
    float pdfSampleLight = 0.0f;
    if (S.numLightTriangles > 0)
        pdfSampleLight = U.lightEnvUniformSample ? U.lightPortion : S.lightSum / (S.lightSum + S.envSum);
    bool sampleLight = ud < pdfSampleLight;
    float pdfSelect = sampleLight ? pdfSampleLight : 1.0f - pdfSampleLight;
    LightLiSample samp;
    if (sampleLight) {
        int id = lightSampleOne(S, f2(us.x, us.y));                                       // lightSampleOneLi :157-161
        samp = lightSampleLi(S, id, x, f2(us.z, us.w), vis);
    } else samp = envSampleLi(S, U, x, us, vis);
    samp.coef /= pdfSelect;
    samp.pdf *= pdfSelect;
    return samp;
}
ZL_DEV LightLiSample sampleLightAndEnv(const DScene& S, const ZlRenderParams& U, float3 x, float ud, float4 us) {
    ImmediateVis vis;
    return sampleLightAndEnv(S, U, x, ud, us, vis);
}
ZL_DEV float pdfSelectLight(const DScene& S, const ZlRenderParams& U, int id) {           // :237-242
    float fstPdf = luminance(lightPower(S, id)) / S.lightSum;
    float sndPdf = U.lightEnvUniformSample ? U.lightPortion : S.lightSum / (S.lightSum + S.envSum);
    return fstPdf * sndPdf;
}
ZL_DEV float pdfSelectEnv(const DScene& S, const ZlRenderParams& U) {                     // :244-247
    return U.lightEnvUniformSample ? (1.0f - U.lightPortion) : S.envSum / (S.lightSum + S.envSum);
}

// surface + material of a hit (the block every integrator repeats, path_integ_naive.glsl:54-67)
struct ShadingPoint { SurfaceInfo surf; uint32_t matType; BSDFParam mat; };
ZL_DEV ShadingPoint loadShadingPoint(const DScene& S, int id, const SurfaceInfo& surfIn, float3 wo) {
    ShadingPoint sp;
    countEvent(S, 3);
    sp.surf = surfIn;
    int matTexId = __ldg(&S.matTex[id]);
    int matId = matTexId & 0x0000ffff;
    int texId = matTexId >> 16;
    sp.matType = loadMaterialType(S, matId);
    if (sp.matType != Dielectric && sp.matType != ThinDielectric) {
        if (dot(sp.surf.ns, wo) < 0) { sp.surf.ns = -sp.surf.ns; sp.surf.ng = -sp.surf.ng; }
    }
    sp.mat = loadMaterial(S, sp.matType, matId, texId, sp.surf.uv);
    return sp;
}

}  // namespace zl
