// ORACLE — test infrastructure only (see zo_vec.h header).  Pinned to oracle/_ref by tests/test_ref_parity.py.
// zo_shade.h — function-for-function restatement of the reference shader libraries:
//   src/shader/random.glsl, math.glsl, intersection.glsl, camera.glsl, microfacet.glsl,
//   material.glsl, material_loader.glsl, light.glsl.
// One `Shader` object = one GLSL invocation (it owns the per-invocation globals randSeed,
// sampleOffset, sampleSeed of random.glsl:3,58-59).  Function names equal the GLSL names.
// Deliberate deviations (both mirrored by the CUDA path):
//   * alias-table indices are clamped to n-1 (rand() can return exactly 1.0),
//   * splats with uv == 1.0 are dropped by a bounds check (GL drops the OOB image write),
//   * out-of-range texelFetch returns 0.
#pragma once
#include "zo_scene.h"
#include "../include/zl_libm.h"   // sin/cos/atan2/log/pow pinned for CUDA, oracle and oracle/_ref alike

namespace zo {

static const float Pi = 3.14159265358979323846f;   // math.glsl:4
static const float PiInv = 1.0f / Pi;              // math.glsl:5

struct Ray { vec3 ori, dir; };
struct SurfaceInfo { vec3 ns, ng; vec2 uv; };
struct HitInfo { bool hit; float dist; };
struct CameraPdf { float pdfPos, pdfDir; };
struct CameraIiSample { vec3 wi; vec3 Ii; float dist; vec2 uv; float pdf; };
struct LightPdf { float pdfPos, pdfDir; };
struct LightLiSample { vec3 wi; vec3 coef; float pdf; };
struct LightLeSample { Ray ray; vec3 Le; float pdfPos; float pdfDir; };

// material.glsl:4-21
enum : uint32_t { Diffuse = 1u << 0, GlosRefl = 1u << 1, GlosTrans = 1u << 2, SpecRefl = 1u << 3,
                  SpecTrans = 1u << 4, Invalid = 1u << 16 };
enum : uint32_t { Lambertian = 0, PrincipledBRDF = 1, MetalWorkflow = 2, Dielectric = 3, ThinDielectric = 4 };
enum : uint32_t { Radiance = 0, Importance = 1 };

struct BSDFParam {   // material.glsl:23-39
    vec3 baseColor; float subsurface = 0;
    float metallic = 0, roughness = 0, specular = 0, specularTint = 0;
    float sheen = 0, sheenTint = 0, clearcoat = 0, clearcoatGloss = 0;
    float ior = 0;
};
struct BSDFSample { vec3 wi; float pdf; vec3 bsdf; float eta; uint32_t flag; };   // material.glsl:41-48

inline BSDFSample makeBSDFSample(vec3 wi, float pdf, vec3 bsdf, float eta, uint32_t flag) {
    BSDFSample s; s.wi = wi; s.pdf = pdf; s.bsdf = bsdf; s.eta = eta; s.flag = flag; return s;
}
inline BSDFSample InvalidBSDFSample() { return makeBSDFSample(vec3(0.0f), 0.0f, vec3(0.0f), 0.0f, Invalid); }
inline Ray makeRay(vec3 o, vec3 d) { Ray r; r.ori = o; r.dir = d; return r; }
inline vec3 rayPoint(Ray r, float t) { return r.ori + r.dir * t; }                       // intersection.glsl:26-29
inline Ray rayOffseted(vec3 ori, vec3 dir) { return makeRay(ori + dir * 1e-4f, dir); }    // :31-37
inline Ray rayOffseted(Ray r) { return rayOffseted(r.ori, r.dir); }                       // :39-42

// ---- random.glsl:5-13 ----
inline uint32_t hash(uint32_t seed) {
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}

// ---- math.glsl ----
inline float square(float x) { return x * x; }                                            // :8-11
inline float biHeuristic(float f, float g) { return f * f / (f * f + g * g); }            // :20-23
inline vec2 toConcentricDisk(vec2 v) {                                                    // :25-41
    if (v.x == 0.0f && v.y == 0.0f) return vec2(0.0f, 0.0f);
    v = v * 2.0f - 1.0f;
    float phi, r;
    if (v.x * v.x > v.y * v.y) { r = v.x; phi = Pi * v.y / v.x * 0.25f; }
    else { r = v.y; phi = Pi * 0.5f - Pi * v.x / v.y * 0.25f; }
    return vec2(r * zl_cosf(phi), r * zl_sinf(phi));
}
inline float satDot(vec3 a, vec3 b) { return gmax(dot(a, b), 0.0f); }                     // :43-46
inline float absDot(vec3 a, vec3 b) { return std::fabs(dot(a, b)); }                      // :48-51
inline float distSquare(vec3 x, vec3 y) { return dot(x - y, x - y); }                     // :53-56
inline vec2 sphereToPlane(vec3 uv) {                                                      // :58-64
    float theta = zl_atan2f(uv.y, uv.x);
    if (theta < 0.0f) theta += Pi * 2.0f;
    float phi = zl_atan2f(length(vec2(uv.x, uv.y)), uv.z);
    return vec2(theta * PiInv * 0.5f, phi * PiInv);
}
inline vec3 planeToSphere(vec2 uv) {                                                      // :66-71
    float theta = uv.x * Pi * 2.0f;
    float phi = uv.y * Pi;
    return vec3(zl_cosf(theta) * zl_sinf(phi), zl_sinf(theta) * zl_sinf(phi), zl_cosf(phi));
}
inline vec3 getTangent(vec3 n) { return (std::fabs(n.z) > 0.999f) ? vec3(0, 1, 0) : vec3(0, 0, 1); }  // :73-76
inline mat3 tbnMatrix(vec3 n) {                                                           // :78-84
    vec3 t = getTangent(n);
    vec3 b = normalize(cross(n, t));
    t = cross(b, n);
    return mat3(t, b, n);
}
inline vec3 normalToWorld(vec3 n, vec3 v) { return normalize(tbnMatrix(n) * v); }         // :86-89
inline vec4 sampleCosineWeighted(vec3 n, vec2 u) {                                        // :99-105
    vec2 uv = toConcentricDisk(u);
    float z = std::sqrt(1.0f - dot(uv, uv));
    vec3 v = normalToWorld(n, vec3(uv, z));
    return vec4(v, PiInv * z);
}
inline bool sameHemisphere(vec3 n, vec3 a, vec3 b) { return dot(n, a) * dot(n, b) > 0; }  // :107-110
inline int maxExtent(vec3 v) {                                                            // :112-118
    if (v.x > v.y) return v.x > v.z ? 0 : 2;
    else return v.y > v.z ? 1 : 2;
}
inline float maxComponent(vec3 v) { return gmax(v.x, gmax(v.y, v.z)); }                   // :120-123
inline int cubemapFace(vec3 dir) {                                                        // :125-131
    int maxDim = maxExtent(gabs(dir));
    if (maxDim == 0) return dir.x > 0 ? 0 : 1;
    if (maxDim == 1) return dir.y > 0 ? 2 : 3;
    return dir.z > 0 ? 4 : 5;
}
inline vec3 sampleTriangleUniform(vec3 va, vec3 vb, vec3 vc, vec2 uv) {                   // :139-145
    float r = std::sqrt(uv.y);
    float u = 1.0f - r;
    float v = uv.x * r;
    return va * (1.0f - u - v) + vb * u + vc * v;
}
inline float triangleArea(vec3 va, vec3 vb, vec3 vc) { return 0.5f * length(cross(vc - va, vb - va)); }  // :147-150
inline vec3 rotateZ(vec3 v, float angle) {                                                // :180-185
    float cost = zl_cosf(angle), sint = zl_sinf(angle);
    return vec3(v.x * cost - v.y * sint, v.x * sint + v.y * cost, v.z);
}
inline float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }                      // :187-191
inline float luminance(vec3 c) { return dot(c, vec3(0.299f, 0.587f, 0.114f)); }           // :193-196
inline bool isBlack(vec3 c) { return luminance(c) < 1e-5f; }                              // :198-201
inline bool hasNan(vec3 c) { return std::isnan(c.x) || std::isnan(c.y) || std::isnan(c.z); }  // :203-206

// ---- intersection.glsl:63-109 ----
inline HitInfo intersectTriangle(vec3 a, vec3 b, vec3 c, Ray ray) {
    HitInfo ret; ret.hit = false; ret.dist = 0.0f;
    const float eps = 1e-6f;
    vec3 ab = b - a, ac = c - a;
    vec3 o = ray.ori, d = ray.dir;
    vec3 p = cross(d, ac);
    float det = dot(ab, p);
    if (std::fabs(det) < eps) return ret;
    vec3 ao = o - a;
    if (det < 0) { ao = -ao; det = -det; }
    float u = dot(ao, p);
    if (u < 0.0f || u > det) return ret;
    vec3 q = cross(ao, ab);
    float v = dot(d, q);
    if (v < 0.0f || u + v > det) return ret;
    float t = dot(ac, q) / det;
    ret.hit = (t > 0.0f);
    ret.dist = t;
    return ret;
}

// ---- microfacet.glsl ----
inline float schlickW(float cosTheta) { return pow5(1.0f - cosTheta); }                   // :4-7
inline vec3 schlickF(float cosTheta, vec3 F0) { return F0 + (vec3(1.0f) - F0) * pow5(1.0f - cosTheta); }  // :9-12
inline vec3 schlickF(float cosTheta, vec3 F0, float roughness) {                          // :14-17
    return F0 + (gmax(vec3(1.0f - roughness), F0) - F0) * pow5(1.0f - cosTheta);
}
inline float schlickG(float cosTheta, float alpha) {                                      // :19-23
    float k = alpha * 0.5f;
    return cosTheta / (cosTheta * (1.0f - k) + k);
}
inline float smithG(vec3 n, vec3 wo, vec3 wi, float alpha) {                              // :25-28
    return schlickG(absDot(n, wo), alpha) * schlickG(absDot(n, wi), alpha);
}
inline float ggx(float cosTheta, float alpha) {                                           // :30-40
    if (cosTheta < 1e-6f) return 0.0f;
    float a2 = alpha * alpha;
    float nom = a2;
    float denom = cosTheta * cosTheta * (a2 - 1.0f) + 1.0f;
    denom = denom * denom * Pi;
    return nom / denom;
}
inline float ggxD(vec3 n, vec3 m, float alpha) { return ggx(dot(n, m), alpha); }          // :42-45
inline float ggxPdfWm(vec3 n, vec3 m, vec3 wo, float alpha) { return ggx(dot(n, m), alpha); }  // :47-50
inline float ggxPdfVisibleWm(vec3 n, vec3 m, vec3 wo, float alpha) {                      // :52-55
    return ggx(dot(n, m), alpha) * schlickG(dot(n, wo), alpha) * absDot(m, wo) / absDot(n, wo);
}
inline vec3 ggxSampleWm(vec3 n, vec3 wo, float alpha, vec2 u) {                           // :57-64
    vec2 xi = toConcentricDisk(u);
    vec3 h = vec3(xi.x, xi.y, std::sqrt(gmax(0.0f, 1.0f - xi.x * xi.x - xi.y * xi.y)));
    h = normalize(h * vec3(alpha, alpha, 1.0f));
    return normalToWorld(n, h);
}
inline vec3 ggxSampleVisibleWm(vec3 n, vec3 wo, float alpha, vec2 u) {                    // :74-92
    mat3 tbn = tbnMatrix(n);
    mat3 tbnInv = inverse(tbn);
    vec3 vh = normalize((tbnInv * wo) * vec3(alpha, alpha, 1.0f));
    float lensq = vh.x * vh.x + vh.y * vh.y;
    vec3 t1 = lensq > 0.0f ? vec3(-vh.y, vh.x, 0.0f) / std::sqrt(lensq) : vec3(1.0f, 0.0f, 0.0f);
    vec3 t2 = cross(vh, t1);
    vec2 xi = toConcentricDisk(u);
    float s = 0.5f * (1.0f + vh.z);
    xi.y = (1.0f - s) * std::sqrt(1.0f - xi.x * xi.x) + s * xi.y;
    vec3 h = t1 * xi.x + t2 * xi.y + vh * std::sqrt(gmax(0.0f, 1.0f - xi.x * xi.x - xi.y * xi.y));
    h = normalize(vec3(h.x * alpha, h.y * alpha, gmax(0.0f, h.z)));
    return normalToWorld(n, h);
}
inline float gtr1(float cosTheta, float alpha) {                                          // :94-98
    float a2 = alpha * alpha;
    return (a2 - 1.0f) / (2.0f * Pi * zl_logf(alpha) * (1.0f + (a2 - 1.0f) * cosTheta * cosTheta));
}
inline float gtr1D(vec3 n, vec3 m, float alpha) { return gtr1(satDot(n, m), alpha); }     // :100-103
inline vec3 gtr1SampleWm(vec3 n, vec3 wo, float alpha, vec2 u) {                          // :105-115
    float cosTheta = std::sqrt(gmax(0.0f, (1.0f - zl_powf(alpha, 1.0f - u.x)) / (1.0f - alpha)));
    float sinTheta = std::sqrt(gmax(0.0f, 1.0f - cosTheta * cosTheta));
    float phi = 2.0f * u.y * Pi;
    vec3 m = normalize(vec3(zl_cosf(phi) * sinTheta, zl_sinf(phi) * sinTheta, cosTheta));
    if (!sameHemisphere(n, wo, m)) m = -m;
    return normalize(normalToWorld(n, m));
}
inline float gtr1PdfWm(vec3 n, vec3 m, vec3 wo, float alpha) { return gtr1D(n, m, alpha) * absDot(n, m); }  // :117-120

// ---- material.glsl (functions that need no invocation state) ----
inline bool approximateDelta(float roughness) { return roughness < 0.02f; }               // :63-66
inline vec3 lambertian(vec3, vec3, vec3, const BSDFParam& p, uint32_t) { return p.baseColor * PiInv; }   // :68-71
inline float lambertianPdf(vec3, vec3 wi, vec3 n, const BSDFParam&, uint32_t) { return satDot(wi, n) * PiInv; }  // :73-76
inline BSDFSample lambertianSample(vec3 n, vec3, const BSDFParam& p, uint32_t, vec3 u) {  // :78-83
    vec3 wi = sampleCosineWeighted(n, vec2(u.y, u.z)).xyz();
    float pdf = satDot(n, wi) * PiInv;
    return makeBSDFSample(wi, pdf, p.baseColor * PiInv, 1.0f, Diffuse);
}
inline vec3 metalWorkflow(vec3 wo, vec3 wi, vec3 n, const BSDFParam& param, uint32_t) {   // :85-114
    vec3 baseColor = param.baseColor;
    float metallic = param.metallic, roughness = param.roughness;
    float alpha = square(roughness);
    vec3 h = normalize(wi + wo);
    if (!sameHemisphere(n, wo, wi)) return vec3(0.0f);
    float cosWi = dot(n, wi), cosWo = dot(n, wo);
    vec3 f0 = mix(vec3(0.04f), baseColor, metallic);
    vec3 f = schlickF(satDot(h, wo), f0, roughness);
    float d = ggxD(n, h, alpha);
    float g = smithG(n, wo, wi, alpha);
    vec3 ks = f;
    vec3 kd = vec3(1.0f) - ks;
    kd *= 1.0f - metallic;
    float denom = 4.0f * cosWo * cosWi;
    if (denom < 1e-7f) return vec3(0.0f);
    return kd * baseColor * PiInv + f * d * g / denom;
}
inline float metalWorkflowPdf(vec3 wo, vec3 wi, vec3 n, const BSDFParam& param, uint32_t) {  // :116-124
    float alpha = square(param.roughness);
    vec3 h = normalize(wo + wi);
    float pdfDiff = satDot(n, wi) * PiInv;
    float pdfSpec = ggxPdfVisibleWm(n, h, wo, alpha) / (4.0f * absDot(h, wo));
    float spec = 1.0f / (2.0f - param.metallic);
    return mix(pdfDiff, pdfSpec, spec);
}
inline BSDFSample metalWorkflowSample(vec3 n, vec3 wo, const BSDFParam& param, uint32_t mode, vec3 u) {  // :126-149
    float roughness = param.roughness;
    float alpha = square(roughness);
    float spec = 1.0f / (2.0f - param.metallic);
    uint32_t type = u.x > spec ? Diffuse : GlosRefl;
    vec3 wi;
    if (type == Diffuse) wi = sampleCosineWeighted(n, vec2(u.y, u.z)).xyz();
    else {
        vec3 h = ggxSampleVisibleWm(n, wo, alpha, vec2(u.y, u.z));
        wi = reflect(-wo, h);
    }
    float cosWi = dot(n, wi);
    if (cosWi < 0) return InvalidBSDFSample();
    vec3 bsdf = metalWorkflow(wo, wi, n, param, mode);
    float pdf = metalWorkflowPdf(wo, wi, n, param, mode);
    return makeBSDFSample(wi, pdf, bsdf, 1.0f, type);
}
inline bool refract(vec3& wt, vec3 wi, vec3 n, float eta) {                               // :151-165 (user overload)
    float cosTi = dot(n, wi);
    if (cosTi < 0) eta = 1.0f / eta;
    float sin2Ti = gmax(0.0f, 1.0f - cosTi * cosTi);
    float sin2Tt = sin2Ti / (eta * eta);
    if (sin2Tt >= 1.0f) return false;
    float cosTt = std::sqrt(1.0f - sin2Tt);
    if (cosTi < 0) cosTt = -cosTt;
    wt = normalize(-wi / eta + n * (cosTi / eta - cosTt));
    return true;
}
inline float fresnelDielectric(float cosTi, float eta) {                                  // :167-186
    cosTi = gclamp(cosTi, -1.0f, 1.0f);
    if (cosTi < 0.0f) { eta = 1.0f / eta; cosTi = -cosTi; }
    float sinTi = std::sqrt(1.0f - cosTi * cosTi);
    float sinTt = sinTi / eta;
    if (sinTt >= 1.0f) return 1.0f;
    float cosTt = std::sqrt(1.0f - sinTt * sinTt);
    float rPa = (cosTi - eta * cosTt) / (cosTi + eta * cosTt);
    float rPe = (eta * cosTi - cosTt) / (eta * cosTi + cosTt);
    return (rPa * rPa + rPe * rPe) * 0.5f;
}
inline vec3 dielectric(vec3 wo, vec3 wi, vec3 n, const BSDFParam& param, uint32_t mode) {  // :188-223
    vec3 baseColor = param.baseColor;
    float roughness = param.roughness, ior = param.ior;
    if (approximateDelta(roughness)) return vec3(0.0f);
    vec3 h = normalize(wo + wi);
    float hCosWo = absDot(h, wo), hCosWi = absDot(h, wi);
    float alpha = roughness * roughness;
    if (sameHemisphere(n, wo, wi)) {
        float refl = fresnelDielectric(absDot(h, wi), ior);
        return (hCosWo * hCosWi < 1e-7f) ? vec3(0.0f)
             : baseColor * ggxD(n, h, alpha) * smithG(n, wo, wi, alpha) / (4.0f * hCosWo * hCosWi) * refl;
    } else {
        float eta = dot(n, wi) > 0 ? ior : 1.0f / ior;
        float sqrtDenom = dot(h, wo) + eta * dot(h, wi);
        float denom = sqrtDenom * sqrtDenom;
        denom *= absDot(n, wi) * absDot(n, wo);
        float refl = fresnelDielectric(dot(h, wi), eta);
        float factor = (mode == Radiance) ? square(1.0f / eta) : 1.0f;
        return (denom < 1e-7f) ? vec3(0.0f)
             : baseColor * std::fabs(ggxD(n, h, alpha) * smithG(n, wo, wi, alpha) * hCosWo * hCosWi) / denom * (1.0f - refl) * factor;
    }
}
inline float dielectricPdf(vec3 wo, vec3 wi, vec3 n, const BSDFParam& param, uint32_t) {  // :225-252
    float roughness = param.roughness, ior = param.ior;
    if (approximateDelta(roughness)) return 0.0f;
    if (sameHemisphere(n, wo, wi)) {
        vec3 h = normalize(wo + wi);
        if (dot(wo, h) < 0.0f) return 0.0f;
        float refl = fresnelDielectric(absDot(h, wi), ior);
        return ggxPdfWm(n, h, wo, roughness * roughness) / (4.0f * absDot(h, wo)) * refl;
    } else {
        float eta = dot(n, wo) > 0 ? ior : 1.0f / ior;
        vec3 h = normalize(wo + wi * eta);
        if (sameHemisphere(h, wo, wi)) return 0.0f;
        float trans = 1.0f - fresnelDielectric(absDot(h, wo), eta);
        float dHdWi = absDot(h, wi) / square(dot(h, wo) + eta * dot(h, wi));
        return ggxPdfWm(n, h, wo, roughness * roughness) * dHdWi * trans;
    }
}
inline BSDFSample dielectricSample(vec3 n, vec3 wo, const BSDFParam& param, uint32_t mode, vec3 u) {  // :254-338
    vec3 baseColor = param.baseColor;
    float roughness = param.roughness, ior = param.ior;
    if (approximateDelta(roughness)) {
        float refl = fresnelDielectric(dot(n, wo), ior);
        if (u.x < refl) {
            vec3 wi = reflect(-wo, n);
            return makeBSDFSample(wi, 1.0f, baseColor, 1.0f, SpecRefl);
        } else {
            vec3 wi;
            bool refr = refract(wi, wo, n, ior);
            if (!refr) return InvalidBSDFSample();
            if (dot(n, wo) < 0) ior = 1.0f / ior;
            float factor = (mode == Radiance) ? square(1.0f / ior) : 1.0f;
            return makeBSDFSample(wi, 1.0f, baseColor * factor, ior, SpecTrans);
        }
    } else {
        float alpha = roughness * roughness;
        vec3 h = ggxSampleWm(n, wo, alpha, vec2(u.y, u.z));
        if (dot(n, h) < 0.0f) h = -h;
        float refl = fresnelDielectric(dot(h, wo), ior);
        if (u.x < refl) {
            vec3 wi = -reflect(wo, h);
            if (!sameHemisphere(n, wo, wi)) return InvalidBSDFSample();
            float p = ggxPdfWm(n, h, wo, alpha) / (4.0f * absDot(h, wo));
            float hCosWo = absDot(h, wo), hCosWi = absDot(h, wi);
            vec3 r = (hCosWo * hCosWi < 1e-7f) ? vec3(0.0f)
                   : baseColor * ggxD(n, h, alpha) * smithG(n, wo, wi, alpha) / (4.0f * hCosWo * hCosWi);
            if (std::isnan(p)) p = 0.0f;
            return makeBSDFSample(wi, p, r, 1.0f, GlosRefl);
        } else {
            vec3 wi;
            bool refr = refract(wi, wo, h, ior);
            if (!refr) return InvalidBSDFSample();
            if (sameHemisphere(n, wo, wi)) return InvalidBSDFSample();
            if (absDot(n, wi) < 1e-10f) return InvalidBSDFSample();
            float hCosWo = absDot(h, wo), hCosWi = absDot(h, wi);
            if (dot(h, wo) < 0) ior = 1.0f / ior;
            float sqrtDenom = dot(h, wo) + ior * dot(h, wi);
            float denom = sqrtDenom * sqrtDenom;
            float dHdWi = hCosWi / denom;
            float factor = (mode == Radiance) ? square(1.0f / ior) : 1.0f;
            denom *= absDot(n, wi) * absDot(n, wo);
            vec3 t = (denom < 1e-7f) ? vec3(0.0f)
                   : baseColor * std::fabs(ggxD(n, h, alpha) * smithG(n, wo, wi, alpha) * hCosWo * hCosWi) / denom;
            float p = ggxPdfWm(n, h, wo, alpha) * dHdWi;
            if (std::isnan(p)) p = 0.0f;
            return makeBSDFSample(wi, p, t * factor, ior, GlosTrans);
        }
    }
}
inline BSDFSample thinDielectricSample(vec3 n, vec3 wo, const BSDFParam& param, uint32_t, vec3 u) {  // :340-355
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
inline vec3 principledMetal(vec3 wo, vec3 wi, vec3 n, vec3 fm0, float alpha) {            // :357-373
    float cosWo = satDot(n, wo), cosWi = satDot(n, wi);
    vec3 h = normalize(wo + wi);
    if (cosWo < 1e-10f || cosWi < 1e-10f) return vec3(0.0f);
    vec3 fm = schlickF(absDot(h, wo), fm0);
    float dm = ggxD(n, h, alpha);
    float gm = smithG(n, wo, wi, alpha);
    float denom = 4.0f * cosWi * cosWo;
    if (denom < 1e-7f) return vec3(0.0f);
    return fm * dm * gm / denom;
}
inline float principledMetalPdf(vec3 wo, vec3 wi, vec3 n, float alpha) {                  // :375-379
    vec3 h = normalize(wo + wi);
    return ggxPdfVisibleWm(n, h, wo, alpha) / (4.0f * absDot(h, wo));
}
inline BSDFSample principledMetalSample(vec3 n, vec3 wo, vec3 fm0, float alpha, vec3 u) {  // :381-391
    vec3 h = ggxSampleVisibleWm(n, wo, alpha, vec2(u.y, u.z));
    vec3 wi = reflect(-wo, h);
    if (dot(n, wi) < 0.0f) return InvalidBSDFSample();
    vec3 bsdf = principledMetal(wo, wi, n, fm0, alpha);
    float pdf = principledMetalPdf(wo, wi, n, alpha);
    return makeBSDFSample(wi, pdf, bsdf, 1.0f, GlosRefl);
}
inline vec3 principledClearcoat(vec3 wo, vec3 wi, vec3 n, vec3 baseColor, float alpha) {  // :393-409
    float cosWo = satDot(n, wo), cosWi = satDot(n, wi);
    vec3 h = normalize(wo + wi);
    if (cosWo < 1e-6f || cosWi < 1e-6f) return vec3(0.0f);
    vec3 fc = schlickF(absDot(h, wo), baseColor);
    float dc = gtr1D(n, h, alpha);
    float gc = smithG(n, wo, wi, 0.25f);
    float denom = 4.0f * cosWi * cosWo;
    if (denom < 1e-7f) return vec3(0.0f);
    return fc * dc * gc / denom;
}
inline float principledClearcoatPdf(vec3 wo, vec3 wi, vec3 n, float alpha) {              // :411-415
    vec3 h = normalize(wo + wi);
    return gtr1PdfWm(n, h, wo, alpha) / (4.0f * absDot(h, wo));
}
inline BSDFSample principledClearcoatSample(vec3 n, vec3 wo, vec3 baseColor, float alpha, vec3 u) {  // :417-427
    vec3 h = gtr1SampleWm(n, wo, alpha, vec2(u.y, u.z));
    vec3 wi = reflect(-wo, h);
    if (dot(n, wi) < 0.0f) return InvalidBSDFSample();
    vec3 bsdf = principledClearcoat(wo, wi, n, baseColor, alpha);
    float pdf = principledClearcoatPdf(wo, wi, n, alpha);
    return makeBSDFSample(wi, pdf, bsdf, 1.0f, GlosRefl);
}
inline vec3 principledDiffuse(vec3 wo, vec3 wi, vec3 n, vec3 baseColor, float subsurface, float roughness) {  // :429-452
    float cosWo = satDot(n, wo), cosWi = satDot(n, wi);
    if (cosWo < 1e-10f || cosWi < 1e-10f) return vec3(0.0f);
    vec3 h = normalize(wo + wi);
    float hCosWi = dot(h, wi);
    float hCosWi2 = hCosWi * hCosWi;
    float fi = schlickW(cosWi), fo = schlickW(cosWo);
    vec3 fd90 = vec3(0.5f + 2.0f * roughness * hCosWi2);
    vec3 fd = mix(vec3(1.0f), fd90, fi) * mix(vec3(1.0f), fd90, fo);
    vec3 baseDiffuse = baseColor * fd * PiInv;
    vec3 fss90 = vec3(roughness * hCosWi2);
    vec3 fss = mix(vec3(1.0f), fss90, fi) * mix(vec3(1.0f), fss90, fo);
    vec3 ss = baseColor * PiInv * 1.25f * (fss * (1.0f / (cosWi + cosWo) - 0.5f) + 0.5f);
    return mix(baseDiffuse, ss, subsurface);
}
inline BSDFSample principledDiffuseSample(vec3 n, vec3 wo, vec3 baseColor, float subsurface, float roughness, vec3 u) {  // :454-460
    vec4 samp = sampleCosineWeighted(n, vec2(u.y, u.z));
    vec3 wi = samp.xyz();
    vec3 bsdf = principledDiffuse(wo, wi, n, baseColor, subsurface, roughness);
    return makeBSDFSample(wi, samp.w, bsdf, 1.0f, Diffuse);
}
inline vec3 principledBRDF(vec3 wo, vec3 wi, vec3 n, const BSDFParam& param, uint32_t) {  // :462-490
    vec3 res = vec3(0.0f);
    vec3 baseColor = param.baseColor;
    float subsurface = param.subsurface, metallic = param.metallic, roughness = param.roughness;
    float specular = param.specular, specularTint = param.specularTint;
    float sheen = param.sheen, sheenTint = param.sheenTint;
    float clearcoat = param.clearcoat, clearcoatGloss = param.clearcoatGloss;
    float alpha = square(roughness);
    float clearcoatAlpha = mix(0.1f, 0.001f, clearcoatGloss);
    float lum = luminance(baseColor);
    vec3 tintColor = lum > 0 ? baseColor / lum : vec3(1.0f);
    vec3 fm0 = mix(0.08f * specular * mix(vec3(1.0f), tintColor, specularTint), baseColor, metallic);
    float hCosWi = dot(normalize(wo + wi), wi);
    res += principledDiffuse(wo, wi, n, baseColor, subsurface, roughness) * (1.0f - metallic);
    res += principledMetal(wo, wi, n, fm0, alpha);
    res += principledClearcoat(wo, wi, n, baseColor, clearcoatAlpha) * clearcoat * 0.25f;
    res += mix(vec3(1.0f), tintColor, sheenTint) * schlickW(hCosWi) * sheen * (dot(n, wi) < 0.0f ? 0.0f : 1.0f);
    return res;
}
inline float principledBRDFPdf(vec3 wo, vec3 wi, vec3 n, const BSDFParam& param, uint32_t) {  // :492-514
    float pdf = 0.0f;
    float metallic = param.metallic, roughness = param.roughness;
    float clearcoat = param.clearcoat, clearcoatGloss = param.clearcoatGloss;
    float alpha = square(roughness);
    float clearcoatAlpha = mix(0.1f, 0.001f, clearcoatGloss);
    float spec = 1.0f / (2.0f - metallic);
    float cosinePdf = absDot(n, wi) * PiInv;
    pdf += cosinePdf * (1.0f - spec);
    pdf += principledMetalPdf(wo, wi, n, alpha) * spec;
    pdf += principledClearcoatPdf(wo, wi, n, clearcoatAlpha) * 0.25f * clearcoat;
    return pdf / (1.0f + 0.25f * clearcoat);
}

// =====================================================================================
// One GLSL invocation.
// =====================================================================================
struct Shader {
    const Scene& S;
    const ZlRenderParams& U;
    // random.glsl:3,58-59
    uint32_t randSeed = 0;
    int sampleOffset = 0;
    uint32_t sampleSeed = 0;
    int uSampler;          // LightPath.cpp:49 / TriplePath.cpp:72 force 0 for the light kernels
    // bvhDebug-style counters (intersection.glsl:331-365), used for the roofline byte model
    uint64_t nodeVisits = 0, triTests = 0, rays = 0;
    // Test utility (NOT reference behaviour): also apply the product's conservative rejection for the
    // ignored slab of near-zero direction components, to prove on the CPU that it never changes a hit.
    bool cullIgnoredSlab = false;
    static bool outsideIgnoredSlab(float o, float d, float lo, float hi, float tMax) {
        float reach = std::fabs(d) * tMax + 2e-5f * tMax + 1e-5f * (std::fabs(o) + 1.0f);
        return (o - reach > hi) || (o + reach < lo);
    }

    Shader(const Scene& s, const ZlRenderParams& u, int samplerMode) : S(s), U(u), uSampler(samplerMode) {}

    // camera uniforms (camera.glsl:5-14)
    vec3 uCamF() const { return vec3(U.camera.F[0], U.camera.F[1], U.camera.F[2]); }
    vec3 uCamR() const { return vec3(U.camera.R[0], U.camera.R[1], U.camera.R[2]); }
    vec3 uCamU() const { return vec3(U.camera.U[0], U.camera.U[1], U.camera.U[2]); }
    vec3 uCamPos() const { return vec3(U.camera.pos[0], U.camera.pos[1], U.camera.pos[2]); }
    mat3 uCamMatInv() const {
        const float* m = U.camera.matInv;
        return mat3(vec3(m[0], m[1], m[2]), vec3(m[3], m[4], m[5]), vec3(m[6], m[7], m[8]));
    }

    // ---- random.glsl ----
    float rand() {                                                                        // :15-19
        randSeed = hash(randSeed);
        return (float)randSeed * (1.0f / 4294967296.0f);
    }
    void setRngSeed(uint32_t seed) { randSeed = seed; }                                   // :48-51
    // Sampler::sobolSample (Sampler.cpp:19-28) evaluated on the fly; the reference reads the
    // same value from the precomputed 131072x256 table (Sampler.cpp:48-64).
    uint32_t sobolTable(int flat) const {
        uint32_t index = (uint32_t)(flat / 256);
        int dim = flat % 256;
        uint32_t r = 0;
        for (int i = dim * 32; index != 0; index >>= 1, i++)
            if (index & 1u) r ^= S.sobolMatrices[i];
        return r;
    }
    float sample1D(int& s) {                                                              // :61-70
        if (uSampler == 0) return rand();
        uint32_t r = sobolTable(sampleOffset + s);
        r ^= sampleSeed;
        sampleSeed = hash(sampleSeed);
        s++;
        return (float)r / 4294967296.0f;
    }
    vec2 sample2D(int& s) { float a = sample1D(s); float b = sample1D(s); return vec2(a, b); }          // :72-75
    vec3 sample3D(int& s) { float a = sample1D(s); float b = sample1D(s); float c = sample1D(s); return vec3(a, b, c); }  // :77-80
    vec4 sample4D(int& s) { float a = sample1D(s); float b = sample1D(s); float c = sample1D(s); float d = sample1D(s); return vec4(a, b, c, d); }  // :82-85

    // ---- intersection.glsl ----
    HitInfo intersectTriangle(int id, Ray ray) const {                                    // :111-121
        int ia = S.fetchIndex(id * 3 + 0), ib = S.fetchIndex(id * 3 + 1), ic = S.fetchIndex(id * 3 + 2);
        return zo::intersectTriangle(S.fetchVertex(ia), S.fetchVertex(ib), S.fetchVertex(ic), ray);
    }
    vec3 triangleSampleUniform(int id, vec2 u) const {                                    // :123-134
        int ia = S.fetchIndex(id * 3 + 0), ib = S.fetchIndex(id * 3 + 1), ic = S.fetchIndex(id * 3 + 2);
        return sampleTriangleUniform(S.fetchVertex(ia), S.fetchVertex(ib), S.fetchVertex(ic), u);
    }
    float triangleArea(int id) const {                                                    // :136-147
        int ia = S.fetchIndex(id * 3 + 0), ib = S.fetchIndex(id * 3 + 1), ic = S.fetchIndex(id * 3 + 2);
        return zo::triangleArea(S.fetchVertex(ia), S.fetchVertex(ib), S.fetchVertex(ic));
    }
    vec3 triangleNormalShad(int id, vec3 p) const {                                       // :149-173
        int ia = S.fetchIndex(id * 3 + 0), ib = S.fetchIndex(id * 3 + 1), ic = S.fetchIndex(id * 3 + 2);
        vec3 a = S.fetchVertex(ia), b = S.fetchVertex(ib), c = S.fetchVertex(ic);
        vec3 na = S.fetchNormal(ia), nb = S.fetchNormal(ib), nc = S.fetchNormal(ic);
        vec3 pa = a - p, pb = b - p, pc = c - p;
        float areaInv = 1.0f / length(cross(b - a, c - a));
        float la = length(cross(pb, pc)) * areaInv;
        float lb = length(cross(pc, pa)) * areaInv;
        float lc = 1.0f - la - lb;
        return normalize(na * la + nb * lb + nc * lc);
    }
    SurfaceInfo triangleSurfaceInfo(int id, vec3 p) const {                               // :188-224
        SurfaceInfo ret;
        int ia = S.fetchIndex(id * 3 + 0), ib = S.fetchIndex(id * 3 + 1), ic = S.fetchIndex(id * 3 + 2);
        vec3 a = S.fetchVertex(ia), b = S.fetchVertex(ib), c = S.fetchVertex(ic);
        vec3 na = S.fetchNormal(ia), nb = S.fetchNormal(ib), nc = S.fetchNormal(ic);
        vec2 ta = S.fetchTexCoord(ia), tb = S.fetchTexCoord(ib), tc = S.fetchTexCoord(ic);
        vec3 pa = a - p, pb = b - p, pc = c - p;
        float areaInv = 1.0f / length(cross(b - a, c - a));
        float la = length(cross(pb, pc)) * areaInv;
        float lb = length(cross(pc, pa)) * areaInv;
        float lc = 1.0f - la - lb;
        ret.ns = normalize(na * la + nb * lb + nc * lc);
        ret.ng = normalize(cross(pa, pb));
        ret.uv = ta * la + tb * lb + tc * lc;
        if (dot(ret.ns, ret.ng) < 0) ret.ng = -ret.ng;
        return ret;
    }
    bool boxHit(int id, Ray ray, float& tMin) const {                                     // :226-329
        float tMax;
        vec3 pMin = S.fetchBound(id * 2 + 0), pMax = S.fetchBound(id * 2 + 1);
        const float eps = 1e-6f;
        vec3 o = ray.ori, d = ray.dir;
        if (std::fabs(d.x) > 1.0f - eps) {
            if (o.y > pMin.y && o.y < pMax.y && o.z > pMin.z && o.z < pMax.z) {
                float dxInv = 1.0f / d.x;
                float ta = (pMin.x - o.x) * dxInv, tb = (pMax.x - o.x) * dxInv;
                tMin = gmin(ta, tb); tMax = gmax(ta, tb);
                return tMax >= 0.0f && tMax >= tMin;
            } else return false;
        }
        if (std::fabs(d.y) > 1.0f - eps) {
            if (o.x > pMin.x && o.x < pMax.x && o.z > pMin.z && o.z < pMax.z) {
                float dyInv = 1.0f / d.y;
                float ta = (pMin.y - o.y) * dyInv, tb = (pMax.y - o.y) * dyInv;
                tMin = gmin(ta, tb); tMax = gmax(ta, tb);
                return tMax >= 0.0f && tMax >= tMin;
            } else return false;
        }
        if (std::fabs(d.z) > 1.0f - eps) {
            if (o.x > pMin.x && o.x < pMax.x && o.y > pMin.y && o.y < pMax.y) {
                float dzInv = 1.0f / d.z;
                float ta = (pMin.z - o.z) * dzInv, tb = (pMax.z - o.z) * dzInv;
                tMin = gmin(ta, tb); tMax = gmax(ta, tb);
                return tMax >= 0.0f && tMax >= tMin;
            } else return false;
        }
        vec3 dInv = 1.0f / d;
        vec3 vta = (pMin - o) * dInv, vtb = (pMax - o) * dInv;
        vec3 vtMin = gmin(vta, vtb), vtMax = gmax(vta, vtb);
        vec3 dt = vtMax - vtMin;
        float tyz = vtMax.z - vtMin.y, tzx = vtMax.x - vtMin.z, txy = vtMax.y - vtMin.x;
        if (std::fabs(d.x) < eps) {
            if (dt.y + dt.z > tyz) {
                tMin = gmax(vtMin.y, vtMin.z); tMax = gmin(vtMax.y, vtMax.z);
                if (cullIgnoredSlab && outsideIgnoredSlab(o.x, d.x, pMin.x, pMax.x, tMax)) return false;
                return tMax >= 0.0f && tMax >= tMin;
            }
        }
        if (std::fabs(d.y) < eps) {
            if (dt.z + dt.x > tzx) {
                tMin = gmax(vtMin.z, vtMin.x); tMax = gmin(vtMax.z, vtMax.x);
                if (cullIgnoredSlab && outsideIgnoredSlab(o.y, d.y, pMin.y, pMax.y, tMax)) return false;
                return tMax >= 0.0f && tMax >= tMin;
            }
        }
        if (std::fabs(d.z) < eps) {
            if (dt.x + dt.y > txy) {
                tMin = gmax(vtMin.x, vtMin.y); tMax = gmin(vtMax.x, vtMax.y);
                if (cullIgnoredSlab && outsideIgnoredSlab(o.z, d.z, pMin.z, pMax.z, tMax)) return false;
                return tMax >= 0.0f && tMax >= tMin;
            }
        }
        if (dt.y + dt.z > tyz && dt.z + dt.x > tzx && dt.x + dt.y > txy) {
            tMin = gmax(gmax(vtMin.x, vtMin.y), vtMin.z);
            tMax = gmin(gmin(vtMax.x, vtMax.y), vtMax.z);
            return tMax >= 0.0f && tMax >= tMin;
        }
        return false;
    }
    bool bvhTest(Ray ray, float dist) {                                                   // :367-393
        rays++;
        int uBvhSize = S.bvhSize;
        int tableOffset = cubemapFace(-ray.dir) * uBvhSize;
        int k = 0;
        while (k != uBvhSize) {
            const int32_t* e = &S.hitTable[3 * (size_t)(tableOffset + k)];
            int nodeIndex = e[0], primIndex = e[1];
            nodeVisits++;
            float boxDist = 0.0f;
            bool bHit = boxHit(nodeIndex, ray, boxDist);
            if (!bHit || (bHit && boxDist > dist)) { k = e[2]; continue; }
            if (primIndex >= 0) {
                triTests++;
                HitInfo hInfo = intersectTriangle(primIndex, ray);
                if (hInfo.hit && hInfo.dist < dist) return true;
            }
            k++;
        }
        return false;
    }
    int bvhHit(Ray ray, float& dist) {                                                    // :395-427
        rays++;
        dist = 1e8f;
        int closest = -1;
        int uBvhSize = S.bvhSize;
        int tableOffset = cubemapFace(-ray.dir) * uBvhSize;
        int k = 0;
        while (k != uBvhSize) {
            const int32_t* e = &S.hitTable[3 * (size_t)(tableOffset + k)];
            int nodeIndex = e[0], primIndex = e[1];
            nodeVisits++;
            float boxDist = 0.0f;
            bool bHit = boxHit(nodeIndex, ray, boxDist);
            if (!bHit || (bHit && boxDist > dist)) { k = e[2]; continue; }
            if (primIndex >= 0) {
                triTests++;
                HitInfo hInfo = intersectTriangle(primIndex, ray);
                if (hInfo.hit && hInfo.dist < dist) { dist = hInfo.dist; closest = primIndex; }
            }
            k++;
        }
        return closest;
    }
    bool visible(vec3 x, vec3 y) {                                                        // :429-434
        float dist = distance(x, y) - 2e-5f;
        vec3 wi = normalize(y - x);
        return !bvhTest(makeRay(x + wi * 1e-5f, wi), dist);
    }

    // ---- camera.glsl ----
    bool inFilmBound(vec2 uv) const { return uv.x >= 0 && uv.x <= 1.0f && uv.y >= 0 && uv.y <= 1.0f; }  // :52-55
    bool thinLensCameraDelta() const { return U.camera.lensRadius <= 1e-6f; }             // :57-60
    Ray thinLensCameraSampleRay(vec2 uv, vec4 u) const {                                  // :62-77
        vec2 texelSize = 1.0f / vec2((float)U.filmW, (float)U.filmH);
        vec2 biasedCoord = uv + texelSize * vec2(u.x, u.y);
        vec2 ndc = biasedCoord * 2.0f - 1.0f;
        vec3 pLens = vec3(toConcentricDisk(vec2(u.z, u.w)) * U.camera.lensRadius, 0.0f);
        vec3 pFocusPlane = vec3(ndc * vec2(U.camera.asp, 1.0f) * U.camera.focalDist * U.camera.tanFOV, U.camera.focalDist);
        vec3 dir = pFocusPlane - pLens;
        dir = normalize(uCamR() * dir.x + uCamU() * dir.y + uCamF() * dir.z);
        Ray ret;
        ret.ori = uCamPos() + uCamR() * pLens.x + uCamU() * pLens.y;
        ret.dir = dir;
        return ret;
    }
    vec2 thinLensCameraRasterPos(Ray ray) const {                                         // :79-91
        float cosTheta = dot(ray.dir, uCamF());
        float dFocus = U.camera.focalDist / cosTheta;
        vec3 pFocus = uCamMatInv() * (rayPoint(ray, dFocus) - uCamPos());
        vec2 filmSize = vec2((float)U.filmW, (float)U.filmH);
        float aspect = filmSize.x / filmSize.y;
        pFocus /= vec3(vec2(aspect, 1.0f) * U.camera.tanFOV, 1.0f) * U.camera.focalDist;
        vec2 ndc = vec2(pFocus.x, pFocus.y);
        return (ndc + 1.0f) * 0.5f;
    }
    vec3 thinLensCameraIe(Ray ray) const {                                                // :93-107
        float cosTheta = dot(ray.dir, uCamF());
        if (cosTheta < 1e-6f) return vec3(0.0f);
        vec2 pRaster = thinLensCameraRasterPos(ray);
        if (!inFilmBound(pRaster)) return vec3(0.0f);
        float tanFOVInv = 1.0f / U.camera.tanFOV;
        float cos2Theta = cosTheta * cosTheta;
        float lensArea = thinLensCameraDelta() ? 1.0f : Pi * U.camera.lensRadius * U.camera.lensRadius;
        return vec3(0.25f) * square(tanFOVInv / cos2Theta) / (lensArea * U.camera.asp);
    }
    CameraIiSample thinLensCameraSampleIi(vec3 ref, vec2 u) const {                       // :109-127
        CameraIiSample inv; inv.wi = vec3(0.0f); inv.Ii = vec3(0.0f); inv.dist = 0.0f; inv.uv = vec2(0.0f); inv.pdf = 0.0f;
        vec3 pLens = vec3(toConcentricDisk(u) * U.camera.lensRadius, 0.0f);
        vec3 y = uCamPos() + uCamR() * pLens.x + uCamU() * pLens.y + uCamF() * pLens.z;
        float dist = distance(ref, y);
        vec3 wi = normalize(y - ref);
        float cosTheta = satDot(uCamF(), -wi);
        if (cosTheta < 1e-6f) return inv;
        Ray ray = makeRay(y, -wi);
        vec3 Ie = thinLensCameraIe(ray);
        vec2 uv = thinLensCameraRasterPos(ray);
        float lensArea = thinLensCameraDelta() ? 1.0f : Pi * U.camera.lensRadius * U.camera.lensRadius;
        float pdf = dist * dist / (cosTheta * lensArea);
        CameraIiSample r; r.wi = wi; r.Ii = Ie; r.dist = dist; r.uv = uv; r.pdf = pdf;
        return r;
    }
    CameraPdf thinLensCameraPdfIe(Ray ray) const {                                        // :129-142
        CameraPdf z; z.pdfPos = 0.0f; z.pdfDir = 0.0f;
        float cosTheta = dot(uCamF(), ray.dir);
        if (cosTheta < 1e-6f) return z;
        vec2 pRaster = thinLensCameraRasterPos(ray);
        if (!inFilmBound(pRaster)) return z;
        CameraPdf r;
        r.pdfPos = thinLensCameraDelta() ? 1.0f : 1.0f / (Pi * U.camera.lensRadius * U.camera.lensRadius);
        r.pdfDir = 1.0f / (cosTheta * cosTheta * cosTheta);
        return r;
    }

    // ---- material.glsl:516-555 (needs rand()) ----
    BSDFSample principledBRDFSample(vec3 n, vec3 wo, const BSDFParam& param, uint32_t mode, vec3 u) {
        vec3 baseColor = param.baseColor;
        float subsurface = param.subsurface, metallic = param.metallic, roughness = param.roughness;
        float specular = param.specular, specularTint = param.specularTint;
        float clearcoat = param.clearcoat;
        float alpha = square(roughness);
        float spec = 1.0f / (2.0f - metallic);
        vec3 wi;
        float cdf[3];
        cdf[0] = 1.0f - spec;
        cdf[1] = 1.0f;
        cdf[2] = 1.0f + clearcoat * 0.25f;
        float s = rand() * cdf[2];
        if (s <= cdf[0]) wi = principledDiffuseSample(n, wo, baseColor, subsurface, roughness, u).wi;
        else if (s <= cdf[1]) {
            float lum = luminance(baseColor);
            vec3 tintColor = lum > 0 ? baseColor / lum : vec3(1.0f);
            vec3 fm0 = mix(0.08f * specular * mix(vec3(1.0f), tintColor, specularTint), baseColor, metallic);
            wi = principledMetalSample(n, wo, fm0, alpha, u).wi;
        } else wi = principledClearcoatSample(n, wo, baseColor, alpha, u).wi;
        vec3 bsdf = principledBRDF(wo, wi, n, param, mode);
        float pdf = principledBRDFPdf(wo, wi, n, param, mode);
        return makeBSDFSample(wi, pdf, bsdf, 1.0f, Diffuse);
    }

    // ---- material_loader.glsl ----
    uint32_t loadMaterialType(int matId) const { return (uint32_t)S.fetchMatTypeBits(matId * 4 + 3, 1); }  // :3-6
    vec3 loadTexturedBase(int texId, vec2 uv) const {                                     // :14-15
        vec2 uvScale = vec2(S.texUVScale[2 * texId], S.texUVScale[2 * texId + 1]);
        return S.sampleAlbedo(fract(uv) * uvScale, texId);
    }
    BSDFParam loadMaterial(uint32_t matType, int matId, int texId, vec2 uv) const {       // :8-97
        BSDFParam ret;
        vec4 baseRou = S.fetchMaterial(matId * 4 + 0);
        ret.baseColor = (texId == -1) ? baseRou.xyz() : loadTexturedBase(texId, uv);
        switch (matType) {
        case PrincipledBRDF: {
            vec4 a = S.fetchMaterial(matId * 4 + 1), b = S.fetchMaterial(matId * 4 + 2);
            ret.roughness = mix(0.0134f, 1.0f, baseRou.w);
            ret.subsurface = a.x; ret.metallic = a.y; ret.specular = a.z; ret.specularTint = a.w;
            ret.sheen = b.x; ret.sheenTint = b.y; ret.clearcoat = b.z; ret.clearcoatGloss = b.w;
            break; }
        case MetalWorkflow:
            ret.roughness = mix(0.0134f, 1.0f, baseRou.w);
            ret.metallic = S.fetchMaterial(matId * 4 + 1).y;
            break;
        case Dielectric:
            ret.roughness = baseRou.w;
            ret.ior = S.fetchMaterial(matId * 4 + 3).x;
            break;
        case ThinDielectric:
            ret.ior = S.fetchMaterial(matId * 4 + 3).x;
            break;
        default: break;   // Lambertian and unknown types: baseColor only
        }
        return ret;
    }
    vec3 materialBSDF(uint32_t matType, const BSDFParam& p, vec3 wo, vec3 wi, vec3 n, uint32_t mode) const {  // :99-115
        switch (matType) {
        case Lambertian: return lambertian(wo, wi, n, p, mode);
        case PrincipledBRDF: return principledBRDF(wo, wi, n, p, mode);
        case MetalWorkflow: return metalWorkflow(wo, wi, n, p, mode);
        case Dielectric: return dielectric(wo, wi, n, p, mode);
        case ThinDielectric: return vec3(0.0f);
        }
        return lambertian(wo, wi, n, p, mode);
    }
    float materialPdf(uint32_t matType, const BSDFParam& p, vec3 wo, vec3 wi, vec3 n, uint32_t mode) const {  // :135-151
        switch (matType) {
        case Lambertian: return lambertianPdf(wo, wi, n, p, mode);
        case PrincipledBRDF: return principledBRDFPdf(wo, wi, n, p, mode);
        case MetalWorkflow: return metalWorkflowPdf(wo, wi, n, p, mode);
        case Dielectric: return dielectricPdf(wo, wi, n, p, mode);
        case ThinDielectric: return 0.0f;
        }
        return lambertianPdf(wo, wi, n, p, mode);
    }
    vec4 materialBSDFAndPdf(uint32_t matType, const BSDFParam& p, vec3 wo, vec3 wi, vec3 n, uint32_t mode) const {  // :117-133
        if (matType == ThinDielectric) return vec4(0.0f);
        return vec4(materialBSDF(matType, p, wo, wi, n, mode), materialPdf(matType, p, wo, wi, n, mode));
    }
    BSDFSample materialSample(uint32_t matType, const BSDFParam& p, vec3 n, vec3 wo, uint32_t mode, vec3 u) {  // :153-169
        switch (matType) {
        case Lambertian: return lambertianSample(n, wo, p, mode, u);
        case PrincipledBRDF: return principledBRDFSample(n, wo, p, mode, u);
        case MetalWorkflow: return metalWorkflowSample(n, wo, p, mode, u);
        case Dielectric: return dielectricSample(n, wo, p, mode, u);
        case ThinDielectric: return thinDielectricSample(n, wo, p, mode, u);
        }
        return lambertianSample(n, wo, p, mode, u);
    }

    // ---- light.glsl ----
    int lightSampleOne(vec2 u) const {                                                    // :68-72
        int cx = (int)((float)S.numLightTriangles * u.x);
        if (cx > S.numLightTriangles - 1) cx = S.numLightTriangles - 1;   // App. B #13 fix
        return (u.y < S.lightProb[cx]) ? cx : S.lightAlias[cx];
    }
    float lightPdfSampleOne(int id) const { return luminance(S.fetchLightPower(id)) / S.lightSum; }  // :74-77
    vec3 lightLe(int id, vec3 x, vec3 wo) const {                                         // :79-86
        int triId = id + S.objPrimCount;
        SurfaceInfo sInfo = triangleSurfaceInfo(triId, x);
        if (dot(wo, sInfo.ng) <= 0.0f) return vec3(0.0f);
        return S.fetchLightPower(id) / triangleArea(triId) * 0.5f * PiInv;
    }
    float lightPdfLi(int id, vec3 x, vec3 y) const {                                      // :88-98
        int triId = id + S.objPrimCount;
        vec3 norm = triangleSurfaceInfo(triId, y).ng;
        vec3 yx = normalize(x - y);
        float cosTheta = absDot(norm, yx);
        if (cosTheta < 1e-8f) return -1.0f;
        return distSquare(x, y) / (triangleArea(triId) * cosTheta);
    }
    LightPdf lightPdfLe(int id, Ray ray) const {                                          // :100-109
        LightPdf ret;
        int triId = id + S.objPrimCount;
        vec3 norm = triangleSurfaceInfo(triId, ray.ori).ng;
        ret.pdfPos = 1.0f / triangleArea(triId);
        ret.pdfDir = (dot(norm, ray.dir) <= 0) ? 0.0f : 0.5f * PiInv;
        return ret;
    }
    LightLeSample lightSampleOneLe(int id, vec4 u) const {                                // :111-120
        int triId = id + S.objPrimCount;
        vec3 ori = triangleSampleUniform(triId, vec2(u.x, u.y));
        vec3 norm = triangleSurfaceInfo(triId, ori).ng;
        vec4 samp = sampleCosineWeighted(norm, vec2(u.z, u.w));
        LightLeSample r;
        r.ray = rayOffseted(makeRay(ori, samp.xyz()));
        r.Le = lightLe(id, ori, samp.xyz());
        r.pdfPos = 1.0f / triangleArea(triId);
        r.pdfDir = samp.w;
        return r;
    }
    LightLiSample lightSampleLi(int id, vec3 x, vec2 u) {                                 // :122-155
        LightLiSample inv; inv.wi = vec3(0.0f); inv.coef = vec3(0.0f); inv.pdf = 0.0f;
        int triId = id + S.objPrimCount;
        int ia = S.fetchIndex(triId * 3 + 0), ib = S.fetchIndex(triId * 3 + 1), ic = S.fetchIndex(triId * 3 + 2);
        vec3 a = S.fetchVertex(ia), b = S.fetchVertex(ib), c = S.fetchVertex(ic);
        vec3 y = sampleTriangleUniform(a, b, c, u);
        vec3 wi = normalize(y - x);
        vec3 norm = triangleSurfaceInfo(triId, y).ng;
        float cosTheta = dot(norm, -wi);
        if (cosTheta < 1e-6f) return inv;
        Ray lightRay = rayOffseted(x, wi);
        float dist = distance(x, y);
        float pdf = dist * dist / (zo::triangleArea(a, b, c) * cosTheta);
        float testDist = dist - 1e-4f - 1e-6f;
        if (bvhTest(lightRay, testDist) || pdf < 1e-8f) return inv;
        vec3 weight = lightLe(id, y, -wi);
        float pdfSample = luminance(S.fetchLightPower(id)) / S.lightSum;
        pdf *= pdfSample;
        LightLiSample r; r.wi = wi; r.coef = weight / pdf; r.pdf = pdf;
        return r;
    }
    LightLiSample lightSampleOneLi(vec3 x, vec4 u) {                                      // :157-161
        int id = lightSampleOne(vec2(u.x, u.y));
        return lightSampleLi(id, x, vec2(u.z, u.w));
    }
    vec3 envLe(vec3 wi) const {                                                           // :163-167
        wi = rotateZ(wi, -U.envRotation);
        return S.sampleEnv(sphereToPlane(wi));
    }
    float envGetPortion(vec3 wi) const { return luminance(envLe(wi)) / S.envSum; }        // :169-172
    float envPdfLi(vec3 wi) const {                                                       // :174-179
        if (S.envSum == 0.0f) return 0.0f;
        vec2 size = vec2((float)S.envW, (float)S.envH);
        return envGetPortion(wi) * size.x * size.y * 0.5f * square(PiInv);
    }
    vec4 envSampleWi(vec4 u) const {                                                      // :181-205
        int w = S.envW, h = S.envH;
        int rx = (int)((float)h * u.x);
        if (rx > h - 1) rx = h - 1;                                        // App. B #13 fix
        float ry = u.y;
        size_t rTex = (size_t)rx * (w + 1) + w;
        int row = (ry < S.envAliasProb[rTex]) ? rx : S.envAlias[rTex];
        int cx = (int)((float)w * u.z);
        if (cx > w - 1) cx = w - 1;                                        // App. B #13 fix
        float cy = u.w;
        size_t cTex = (size_t)row * (w + 1) + cx;
        int col = (cy < S.envAliasProb[cTex]) ? cx : S.envAlias[cTex];
        // float(row + 0.5): row + 0.5 is evaluated in float (light.glsl:199)
        vec2 uv = vec2((float)col + 0.5f, (float)row + 0.5f) / vec2((float)w, (float)h);
        vec3 wi = planeToSphere(uv);
        wi = rotateZ(wi, U.envRotation);
        float pdf = envGetPortion(wi) * (float)w * (float)h * 0.5f * square(PiInv);
        return vec4(wi, pdf);
    }
    LightLiSample envSampleLi(vec3 x, vec4 u) {                                           // :207-219
        LightLiSample inv; inv.wi = vec3(0.0f); inv.coef = vec3(0.0f); inv.pdf = 0.0f;
        vec4 sp = envSampleWi(u);
        vec3 wi = sp.xyz();
        float pdf = sp.w;
        Ray ray = rayOffseted(x, wi);
        float dist = 1e8f;
        if (bvhTest(ray, dist) || pdf == 0.0f) return inv;
        LightLiSample r; r.wi = wi; r.coef = envLe(wi) / pdf; r.pdf = pdf;
        return r;
    }
    LightLiSample sampleLightAndEnv(vec3 x, float ud, vec4 us) {                          // :221-235
        float pdfSampleLight = 0.0f;
        if (S.numLightTriangles > 0)
            pdfSampleLight = U.lightEnvUniformSample ? U.lightPortion : S.lightSum / (S.lightSum + S.envSum);
        bool sampleLight = ud < pdfSampleLight;
        float pdfSelect = sampleLight ? pdfSampleLight : 1.0f - pdfSampleLight;
        LightLiSample samp = sampleLight ? lightSampleOneLi(x, us) : envSampleLi(x, us);
        samp.coef /= pdfSelect;
        samp.pdf *= pdfSelect;
        return samp;
    }
    float pdfSelectLight(int id) const {                                                  // :237-242
        float fstPdf = luminance(S.fetchLightPower(id)) / S.lightSum;
        float sndPdf = U.lightEnvUniformSample ? U.lightPortion : S.lightSum / (S.lightSum + S.envSum);
        return fstPdf * sndPdf;
    }
    float pdfSelectEnv() const {                                                          // :244-247
        return U.lightEnvUniformSample ? (1.0f - U.lightPortion) : S.envSum / (S.lightSum + S.envSum);
    }
};

}  // namespace zo
