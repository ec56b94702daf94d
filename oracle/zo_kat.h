// ORACLE — test infrastructure only (see zo_vec.h header).  Pinned to oracle/_ref by tests/test_ref_parity.py.
// zo_kat.h — per-function known-answer evaluation; op table documented in
// include/zillum_cuda.h next to zl_debug_eval (the CUDA side implements the same table).
#pragma once
#include "zo_shade.h"

namespace zo {

inline int katEval(const Scene& S, const ZlRenderParams& U, int op, const float* in, int inStride,
                   float* out, int outStride, size_t n) {
    auto B = [](float f) { return (int)floatBits(f); };
    auto F = [](uint32_t u) { return bitsFloat(u); };
    auto V3 = [](const float* p) { return vec3(p[0], p[1], p[2]); };
    auto put3 = [](float* o, vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; };
    for (size_t i = 0; i < n; i++) {
        const float* a = in + i * inStride;
        float* o = out + i * outStride;
        Shader sh(S, U, U.sampler);
        switch (op) {
        case ZL_KAT_HASH: o[0] = F(hash((uint32_t)B(a[0]))); break;
        case ZL_KAT_SOBOL: o[0] = F(sh.sobolTable(B(a[0]) * 256 + B(a[1]))); break;
        case ZL_KAT_CUBEMAP_FACE: o[0] = F((uint32_t)cubemapFace(V3(a))); break;
        case ZL_KAT_BOXHIT: {
            // a[0] = threaded entry index k in the hit table of the ray's face
            Ray r = makeRay(V3(a + 1), V3(a + 4));
            int node = S.hitTable[3 * ((size_t)cubemapFace(-r.dir) * S.bvhSize + B(a[0]))];
            float t = 0.0f; bool h = sh.boxHit(node, r, t);
            o[0] = h ? 1.0f : 0.0f; o[1] = h ? t : 0.0f; break; }
        case ZL_KAT_TRIANGLE: {
            HitInfo h = sh.intersectTriangle(B(a[0]), makeRay(V3(a + 1), V3(a + 4)));
            o[0] = h.hit ? 1.0f : 0.0f; o[1] = h.hit ? h.dist : 0.0f; break; }
        case ZL_KAT_SURFACE: {
            SurfaceInfo s = sh.triangleSurfaceInfo(B(a[0]), V3(a + 1));
            put3(o, s.ns); put3(o + 3, s.ng); o[6] = s.uv.x; o[7] = s.uv.y; break; }
        case ZL_KAT_CAMERA_RAY: {
            Ray r = sh.thinLensCameraSampleRay(vec2(a[0], a[1]), vec4(a[2], a[3], a[4], a[5]));
            put3(o, r.ori); put3(o + 3, r.dir); break; }
        case ZL_KAT_CAMERA_II: {
            CameraIiSample c = sh.thinLensCameraSampleIi(V3(a), vec2(a[3], a[4]));
            put3(o, c.wi); put3(o + 3, c.Ii); o[6] = c.dist; o[7] = c.uv.x; o[8] = c.uv.y; o[9] = c.pdf; break; }
        case ZL_KAT_CAMERA_PDF: {
            CameraPdf c = sh.thinLensCameraPdfIe(makeRay(V3(a), V3(a + 3)));
            o[0] = c.pdfPos; o[1] = c.pdfDir; break; }
        case ZL_KAT_BSDF_EVAL: {
            int mat = B(a[0]), tex = B(a[1]);
            uint32_t type = sh.loadMaterialType(mat);
            BSDFParam p = sh.loadMaterial(type, mat, tex, vec2(a[2], a[3]));
            vec4 r = sh.materialBSDFAndPdf(type, p, V3(a + 4), V3(a + 7), V3(a + 10), (uint32_t)B(a[13]));
            o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; break; }
        case ZL_KAT_BSDF_SAMPLE: {
            int mat = B(a[0]), tex = B(a[1]);
            uint32_t type = sh.loadMaterialType(mat);
            BSDFParam p = sh.loadMaterial(type, mat, tex, vec2(a[2], a[3]));
            sh.randSeed = (uint32_t)B(a[14]);
            BSDFSample s = sh.materialSample(type, p, V3(a + 7), V3(a + 4), (uint32_t)B(a[10]), V3(a + 11));
            put3(o, s.wi); o[3] = s.pdf; put3(o + 4, s.bsdf); o[7] = s.eta; o[8] = F(s.flag); break; }
        case ZL_KAT_ENV_LE: { put3(o, sh.envLe(V3(a))); o[3] = sh.envPdfLi(V3(a)); break; }
        case ZL_KAT_ENV_SAMPLE: { vec4 r = sh.envSampleWi(vec4(a[0], a[1], a[2], a[3])); o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; break; }
        case ZL_KAT_LIGHT_LE: {
            int id = B(a[0]);
            put3(o, sh.lightLe(id, V3(a + 1), V3(a + 4))); o[3] = sh.lightPdfLi(id, V3(a + 1), V3(a + 7)); break; }
        case ZL_KAT_LIGHT_SAMPLE_LE: {
            LightLeSample l = sh.lightSampleOneLe(B(a[0]), vec4(a[1], a[2], a[3], a[4]));
            put3(o, l.ray.ori); put3(o + 3, l.ray.dir); put3(o + 6, l.Le); o[9] = l.pdfPos; o[10] = l.pdfDir; break; }
        case ZL_KAT_SAMPLE_LIGHT_ENV: {
            LightLiSample l = sh.sampleLightAndEnv(V3(a), a[3], vec4(a[4], a[5], a[6], a[7]));
            put3(o, l.wi); put3(o + 3, l.coef); o[6] = l.pdf; break; }
        case ZL_KAT_LIBM: {
            const float x = a[1], y = a[2];
            switch (B(a[0])) {
            case 0: o[0] = zl_sinf(x); break; case 1: o[0] = zl_cosf(x); break; case 2: o[0] = zl_atan2f(y, x); break;
            case 3: o[0] = zl_asinf(x); break; case 4: o[0] = zl_acosf(x); break; case 5: o[0] = zl_logf(x); break;
            case 6: o[0] = zl_powf(x, y); break; default: o[0] = zl_expf(x); break;
            }
            break; }
        default: return 1;
        }
    }
    return 0;
}

}  // namespace zo
