// ORACLE — test infrastructure only (see zo_vec.h header).  Pinned to oracle/_ref by tests/test_ref_parity.py.
// zo_api.cpp — extern "C" surface of the CPU oracle, loaded with ctypes by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "zo_host.h"
#include "zo_integrators.h"
#include "zo_kat.h"

using namespace zo;

extern "C" {

void* zo_scene_create(const ZlSceneDesc* desc) {
    Scene* s = new Scene(*desc);
    if (!desc->hitTable) {          // a scene flattened for device-side MTBVH threading carries no host table: the oracle builds its own
        PackedBVH b = buildBVH(desc->vertices, desc->indices, desc->numTriangles);
        s->hitTable = std::move(b.hitTable);
        if (!desc->bounds) s->bounds = std::move(b.bounds);      // a scene whose BVH is built on the device carries no tree at all
    }
    return s;
}
void zo_scene_destroy(void* s) { delete (Scene*)s; }

int zo_get_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void zo_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static void putStats(const Stats& st, uint64_t* out) {
    if (!out) return;
    out[0] = st.rays; out[1] = st.nodeVisits; out[2] = st.triTests; out[3] = st.paths; out[4] = st.splats;
}

// film: W*H*4 floats, accumulated into (caller clears).  stats: 5 x uint64 or NULL.
// rows rowBegin, rowBegin+rowStride, ... < rowEnd (rowEnd < 0: the whole film)
int zo_path_pass(void* scene, const ZlRenderParams* p, float* film, uint64_t* stats, int rowBegin, int rowEnd, int rowStride) {
    Stats st; pathPass(*(Scene*)scene, *p, film, &st, rowBegin, rowEnd, rowStride); putStats(st, stats); return 0;
}
int zo_light_pass(void* scene, const ZlRenderParams* p, float* film, uint64_t* stats, long idBegin, long idEnd) {
    Stats st; lightPass(*(Scene*)scene, *p, film, &st, idBegin, idEnd); putStats(st, stats); return 0;
}
int zo_triple_pt_pass(void* scene, const ZlRenderParams* p, float* film, uint64_t* stats, int rowBegin, int rowEnd, int rowStride) {
    Stats st; triplePtPass(*(Scene*)scene, *p, film, &st, rowBegin, rowEnd, rowStride); putStats(st, stats); return 0;
}
int zo_triple_lpt_pass(void* scene, const ZlRenderParams* p, float* film, uint64_t* stats, long idBegin, long idEnd) {
    Stats st; tripleLptPass(*(Scene*)scene, *p, film, &st, idBegin, idEnd); putStats(st, stats); return 0;
}

// bvhHit / bvhTest on an explicit ray set (same contract as zl_trace_rays).
// anyhit bit 1 (value 2): additionally apply the product's conservative ignored-slab rejection
// (test utility, see Shader::cullIgnoredSlab).
int zo_trace_rays(void* scene, const float* rays, size_t n, int anyhitFlags, const float* tMax,
                  int32_t* outIds, float* outT, int32_t* outSteps) {
    const Scene& S = *(Scene*)scene;
    ZlRenderParams dummy{};
    const int anyhit = anyhitFlags & 1;
#pragma omp parallel for schedule(dynamic, 1024)
    for (long i = 0; i < (long)n; i++) {
        Shader sh(S, dummy, 0);
        sh.cullIgnoredSlab = (anyhitFlags & 2) != 0;
        Ray r = makeRay(vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]),
                        vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
        if (anyhit) {
            bool occ = sh.bvhTest(r, tMax ? tMax[i] : 1e8f);
            outIds[i] = occ ? 1 : 0;
            if (outT) outT[i] = 0.0f;
        } else {
            float dist;
            int id = sh.bvhHit(r, dist);
            outIds[i] = id;
            if (outT) outT[i] = dist;
        }
        if (outSteps) { outSteps[2 * i] = (int32_t)sh.nodeVisits; outSteps[2 * i + 1] = (int32_t)sh.triTests; }
    }
    return 0;
}

// For each ray: the two smallest distinct-triangle hit distances found by brute force over
// all triangles (no BVH).  Used to classify "documented epsilon ties" in the ID parity test.
int zo_brute_force_two_nearest(void* scene, const float* rays, size_t n, int32_t* ids2, float* t2) {
    const Scene& S = *(Scene*)scene;
    ZlRenderParams dummy{};
#pragma omp parallel for schedule(dynamic, 16)
    for (long i = 0; i < (long)n; i++) {
        Shader sh(S, dummy, 0);
        Ray r = makeRay(vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]),
                        vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
        int b0 = -1, b1 = -1; float d0 = 1e8f, d1 = 1e8f;
        for (int t = 0; t < S.numTriangles; t++) {
            HitInfo h = sh.intersectTriangle(t, r);
            if (!h.hit) continue;
            if (h.dist < d0) { d1 = d0; b1 = b0; d0 = h.dist; b0 = t; }
            else if (h.dist < d1) { d1 = h.dist; b1 = t; }
        }
        ids2[2 * i] = b0; ids2[2 * i + 1] = b1; t2[2 * i] = d0; t2[2 * i + 1] = d1;
    }
    return 0;
}

// ---- host preparation ----
int zo_build_bvh(const float* vertices, const uint32_t* indices, int numTriangles,
                 float* boundsOut /*6*(2T-1)*/, int32_t* hitTableOut /*18*(2T-1)*/) {
    PackedBVH b = buildBVH(vertices, indices, numTriangles);
    std::memcpy(boundsOut, b.bounds.data(), b.bounds.size() * sizeof(float));
    std::memcpy(hitTableOut, b.hitTable.data(), b.hitTable.size() * sizeof(int32_t));
    return b.treeSize;
}
void zo_alias_table(const float* pdf, int n, int32_t* alias, float* prob) { buildAliasTable(pdf, n, alias, prob); }
float zo_env_tables(const float* rgb, int w, int h, int32_t* alias, float* pdf) { return buildEnvTables(rgb, w, h, alias, pdf); }
uint32_t zo_sobol_sample(const uint32_t* matrices, uint32_t index, int dim) { return sobolSample(matrices, index, dim); }
uint32_t zo_hash(uint32_t x) { return hash(x); }
void zo_camera_update(const float* pos, const float* angleDeg, float fovDeg, float aspect, float lensRadius,
                      float focalDist, ZlCamera* out) { cameraUpdate(pos, angleDeg, fovDeg, aspect, lensRadius, focalDist, out); }
float zo_light_table(const float* vertices, const uint32_t* indices, int numMeshes, const int* first,
                     const int* count, const float* power, float* lightPowerOut, float* pdfOut) {
    return buildLightTable(vertices, indices, numMeshes, first, count, power, lightPowerOut, pdfOut);
}
float zo_round_to_half(float f) { return roundToHalf(f); }

// ---- per-function known-answer evaluation (same op table as zl_debug_eval) ----
int zo_debug_eval(void* scene, const ZlRenderParams* p, int op, const float* in, int inStride,
                  float* out, int outStride, size_t n) {
    return katEval(*(Scene*)scene, *p, op, in, inStride, out, outStride, n);
}

// post_proc.glsl:12-59 (display stage): out rgba = gamma(toneMap(clamp(film.rgb * scale, 0, 1e30))), a = 1;
// out8 = the GL_UNSIGNED_BYTE read-back of Texture2D::readFromDevice (Texture.cpp:96-102): clamp to [0,1], x255, round.
void zo_post_proc(const float* film, size_t n, float scale, int toneMapper, float* outRgba, unsigned char* outRgb8) {
    auto calc = [](float x) {                                                          // :22-26, per component
        const float A = 0.22f, B = 0.3f, C = 0.1f, D = 0.2f, E = 0.01f, F = 0.3f;
        return (x * (x * A + B * C) + D * E) / (x * (x * A + B) + D * F) - E / F;
    };
    const float g = 1.0f / 2.2f;
    for (size_t i = 0; i < n; i++)
        for (int c = 0; c < 3; c++) {
            float x = film[4 * i + c] * scale;
            x = fminf(fmaxf(x, 0.0f), 1e30f);
            float m = x;
            if (toneMapper == 1) m = calc(x * 1.6f) / calc(11.2f);                     // filmic, :28-32
            else if (toneMapper == 2) m = (x * (x * 2.51f + 0.03f)) / (x * (x * 2.43f + 0.59f) + 0.14f);   // ACES, :34-37
            m = zl_powf(m, g);
            if (outRgba) outRgba[4 * i + c] = m;
            if (outRgb8) outRgb8[3 * i + c] = (unsigned char)rintf(fminf(fmaxf(m, 0.0f), 1.0f) * 255.0f);
        }
    if (outRgba) for (size_t i = 0; i < n; i++) outRgba[4 * i + 3] = 1.0f;
}

}  // extern "C"
