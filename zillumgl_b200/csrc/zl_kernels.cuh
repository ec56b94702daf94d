// zl_kernels.cuh — __global__ entry points.
//   pathPassKernel       <- path_integ_naive.glsl main (:145-174), NaivePath.cpp:94-100
//   triplePtPassKernel   <- triple_path_pass_pt.glsl main (:197-224), TriplePath.cpp:117-121
//   lightPassKernel      <- light_path_integ.glsl main (:148-159), LightPath.cpp:104
//   tripleLptPassKernel  <- triple_path_pass_lpt.glsl main (:184-193), TriplePath.cpp:123-127
//   traceRaysKernel      <- bvhHit / bvhTest / bvhDebug on an explicit ray buffer
// Launch geometry is chosen for B200, not copied from the 48x32 / 1536-wide GL work groups:
// 128-thread CTAs (4 warps) so that many CTAs fit per SM whatever the register count, and
// for the per-pixel kernels each warp owns an 8x4 pixel tile so primary rays of a warp pick
// the same MTBVH face and walk neighbouring nodes.
#pragma once
#include "zl_integrators.cuh"

#ifdef ZL_INSTRUMENT
#define ZL_KERNEL(name) name##Counted
#else
#define ZL_KERNEL(name) name
#endif

namespace zl {

static constexpr int kTileW = 16, kTileH = 8;      // pixels per CTA: 2x2 warps of 8x4 pixels
static constexpr int kPixelBlock = 128;
static constexpr int kLightBlock = 128;
static constexpr int kTraceBlock = 128;

// Sobol row of this pass into shared memory (sampleOffset is the same for every pixel)
ZL_DEV void stageSobolRow(const DScene& S, const ZlRenderParams& U, uint32_t* row) {
    if (U.sampler == 0) return;
    const uint32_t index = (uint32_t)(passSampleOffset(U.spp) / 256);
    for (int dim = threadIdx.x; dim < 256; dim += blockDim.x) row[dim] = sobolSample(S.sobol, index, dim);
}

ZL_DEV bool pixelOfThread(const ZlRenderParams& U, int& px, int& py) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    px = blockIdx.x * kTileW + (warp & 1) * 8 + (lane & 7);
    py = blockIdx.y * kTileH + (warp >> 1) * 4 + (lane >> 3);
    return px < U.filmW && py < U.filmH;
}

ZL_DEV SamplerState makeSampler(const DScene& S, const ZlRenderParams& U, const uint32_t* row, int samplerMode) {
    SamplerState st;
    st.randSeed = 0; st.sampleSeed = 0;
    st.sampleOffset = passSampleOffset(U.spp);
    st.uSampler = samplerMode;
    st.s = 0;
    st.row = row;
    st.matrices = S.sobol;
    return st;
}

__global__ void __launch_bounds__(kPixelBlock) ZL_KERNEL(pathPassKernel)(const DScene S, const ZlRenderParams U, float4* __restrict__ film) {
    __shared__ uint32_t row[256];
    stageSobolRow(S, U, row);
    __syncthreads();
    int px, py;
    if (!pixelOfThread(U, px, py)) return;
    float2 scrCoord = f2((float)px, (float)py) / f2((float)U.filmW, (float)U.filmH);
    SamplerState st = makeSampler(S, U, row, U.sampler);
    seedPixel(st, S, U, scrCoord);
    Ray ray = thinLensCameraSampleRay(U, scrCoord, sample4D(st));
    countEvent(S, 5);
    float3 result = pathIntegTrace(S, U, ray, st);
    if (!hasNan(result)) {
        float4* p = film + (size_t)py * U.filmW + px;      // one owner per pixel: plain read-modify-write
        float4 v = *p;
        v.x += result.x; v.y += result.y; v.z += result.z;
        *p = v;
    }
}

__global__ void __launch_bounds__(kPixelBlock) ZL_KERNEL(triplePtPassKernel)(const DScene S, const ZlRenderParams U, float4* __restrict__ film) {
    __shared__ uint32_t row[256];
    stageSobolRow(S, U, row);
    __syncthreads();
    int px, py;
    if (!pixelOfThread(U, px, py)) return;
    float2 scrCoord = f2((float)px, (float)py) / f2((float)U.filmW, (float)U.filmH);
    SamplerState st = makeSampler(S, U, row, U.sampler);
    seedPixel(st, S, U, scrCoord);
    Ray ray = thinLensCameraSampleRay(U, scrCoord, sample4D(st));
    countEvent(S, 5);
    float3 result = traceCameraPath(S, U, ray, st);
    if (!hasNan(result)) {
        // addFilm (triple_path_pass_pt.glsl:33-42) is a non-atomic RMW; LPT splats of the previous
        // pass are complete (stream order), splats of this pass start after this kernel.
        float4* p = film + (size_t)py * U.filmW + px;
        float4 v = *p;
        v.x += result.x; v.y += result.y; v.z += result.z;
        *p = v;
    }
}

__global__ void __launch_bounds__(kLightBlock) ZL_KERNEL(lightPassKernel)(const DScene S, const ZlRenderParams U, float4* __restrict__ film, long long total) {
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    SamplerState st = makeSampler(S, U, nullptr, 0);        // uSampler forced to 0 (LightPath.cpp:48-49)
    st.randSeed = (uint32_t)U.spp * ((uint32_t)ZL_LIGHT_GROUP_SIZE * (uint32_t)U.blocksOnePass) + (uint32_t)id + (uint32_t)U.freeCounter;
    countEvent(S, 5);
    lightIntegTrace(S, U, st, film);
}

__global__ void __launch_bounds__(kLightBlock) ZL_KERNEL(tripleLptPassKernel)(const DScene S, const ZlRenderParams U, float4* __restrict__ film, long long total) {
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    SamplerState st = makeSampler(S, U, nullptr, 0);        // TriplePath.cpp:72
    st.randSeed = (uint32_t)U.spp * ((uint32_t)ZL_LIGHT_GROUP_SIZE * (uint32_t)U.blocksOnePass * (uint32_t)U.loopsPerPass) +
                  (uint32_t)id + (uint32_t)U.freeCounter;
    for (int i = 0; i < U.loopsPerPass; i++) { countEvent(S, 5); traceLightPath(S, U, st, film); }
}

#ifndef ZL_INSTRUMENT
// pixel-centre primary rays, row-major: thinLensCameraSampleRay(scrCoord, u = 0)
__global__ void primaryRaysKernel(const ZlRenderParams U, float4* __restrict__ rays) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)U.filmW * U.filmH) return;
    int px = (int)(i % U.filmW), py = (int)(i / U.filmW);
    float2 scrCoord = f2((float)px, (float)py) / f2((float)U.filmW, (float)U.filmH);
    Ray r = thinLensCameraSampleRay(U, scrCoord, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    rays[2 * i] = make_float4(r.ori.x, r.ori.y, r.ori.z, 1e8f);
    rays[2 * i + 1] = make_float4(r.dir.x, r.dir.y, r.dir.z, 0.0f);
}

// One ray per thread.  If the set is a W x H pixel grid, each warp takes an 8x4 tile.
template <bool ANYHIT, bool COUNT, int LEAN = 0>
__global__ void __launch_bounds__(kTraceBlock) traceRaysKernel(const DScene S, const float4* __restrict__ rays, size_t n, int gridW, int gridH,
                                                               int32_t* __restrict__ outIds, float* __restrict__ outT, int2* __restrict__ outSteps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gridW > 0 && (gridW & 7) == 0 && (gridH & 3) == 0) {
        size_t warp = i >> 5; int lane = (int)(i & 31);
        int tilesX = gridW >> 3;
        int tx = (int)(warp % tilesX), ty = (int)(warp / tilesX);
        int px = tx * 8 + (lane & 7), py = ty * 4 + (lane >> 3);
        if (py >= gridH) return;
        i = (size_t)py * gridW + px;
    }
    if (i >= n) return;
    float4 o = __ldg(rays + 2 * i), d = __ldg(rays + 2 * i + 1);
    Ray r = makeRay(f3(o), f3(d));
    TraceCounters cnt{0, 0};
    float dist = o.w;
    int id = COUNT ? traverse<ANYHIT, COUNT>(S, r, dist, &cnt) : traverseWarp<ANYHIT, LEAN>(S, r, dist);
    outIds[i] = id;
    outT[i] = ANYHIT ? 0.0f : dist;
    if (COUNT) outSteps[i] = make_int2(cnt.nodes, cnt.tris);
}

// Measurement aid for the traversal micro-benchmark: how many DISTINCT 32-byte node records and distinct triangles a warp asks the
// memory system for per lock-step step of the closest-hit walk (lanes that sit on the same record share one request), next to the
// per-lane totals.  The per-lane count is the algorithmic work; the per-warp distinct count is what the L1 is really asked for, and
// only that can be held against a bandwidth.  out = {lane node visits, lane triangle tests, warp-distinct node records, warp-distinct triangles}.
__global__ void __launch_bounds__(kTraceBlock) uniqueSectorKernel(const DScene S, const float4* __restrict__ rays, size_t n, int gridW, int gridH,
                                                                  unsigned long long* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool alive = true;
    if (gridW > 0 && (gridW & 7) == 0 && (gridH & 3) == 0) {
        size_t warp = i >> 5; int lane = (int)(i & 31);
        int tilesX = gridW >> 3;
        int tx = (int)(warp % tilesX), ty = (int)(warp / tilesX);
        int px = tx * 8 + (lane & 7), py = ty * 4 + (lane >> 3);
        alive = py < gridH;
        i = (size_t)py * gridW + px;
    }
    alive = alive && i < n;
    Ray r = makeRay(f3(0.0f), f3(0.0f, 0.0f, 1.0f));
    if (alive) { float4 o = __ldg(rays + 2 * i), d = __ldg(rays + 2 * i + 1); r = makeRay(f3(o), f3(d)); }
    const RayPrep rp = prepareRay(r);
    const int nNodes = S.bvhSize;
    const float4* __restrict__ nodes = S.nodes + (size_t)cubemapFace(-r.dir) * (size_t)nNodes * 2;
    const int face = cubemapFace(-r.dir);
    float dist = 1e8f;
    int k = 0;
    unsigned long long laneNodes = 0, laneTris = 0, warpNodes = 0, warpTris = 0;
    alive = alive && nNodes > 0;
    while (__ballot_sync(0xffffffffu, alive)) {
        if (alive) {
            const unsigned same = __match_any_sync(__activemask(), face * nNodes + k);
            if ((__ffs(same) - 1) == (int)(threadIdx.x & 31)) warpNodes++;
            laneNodes++;
            float4 lo, hi;
            loadNode(nodes, k, lo, hi);
            float boxDist;
            const bool bHit = boxHit<false>(f3(lo), f3(hi), rp, boxDist);
            const int prim = __float_as_int(lo.w);
            const bool enter = bHit && !(boxDist > dist);
            const bool leaf = enter && prim >= 0;
            if (leaf) {
                const unsigned sameT = __match_any_sync(__activemask(), prim);
                if ((__ffs(sameT) - 1) == (int)(threadIdx.x & 31)) warpTris++;
                laneTris++;
                const float4* __restrict__ tp = S.triPos + 3 * (size_t)prim;
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                float t;
                if (intersectTriangle(f3(a), f3(b), f3(c), rp.o, rp.d, t) && t < dist) dist = t;
            }
            k = enter ? k + 1 : __float_as_int(hi.w);
            if (k == nNodes) alive = false;
        }
    }
    atomicAdd(out + 0, laneNodes); atomicAdd(out + 1, laneTris); atomicAdd(out + 2, warpNodes); atomicAdd(out + 3, warpTris);
}

__global__ void streamReadKernel(const uint4* __restrict__ buf, size_t n16, unsigned* sink) {
    unsigned acc = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        uint4 v = __ldg(buf + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;   // never true for the 0x01 fill; keeps the loads alive
}

// ---- per-function KAT evaluation (op table in include/zillum_cuda.h) ----
__global__ void katKernel(const DScene S, const ZlRenderParams U, int op, const float* __restrict__ in, int inStride,
                          float* __restrict__ out, int outStride, size_t n) {
    __shared__ uint32_t row[256];
    stageSobolRow(S, U, row);
    __syncthreads();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = in + i * inStride;
    float* o = out + i * outStride;
    auto B = [](float f) { return __float_as_int(f); };
    auto F = [](uint32_t u) { return __uint_as_float(u); };
    auto V3 = [](const float* p) { return f3(p[0], p[1], p[2]); };
    auto put3 = [](float* q, float3 v) { q[0] = v.x; q[1] = v.y; q[2] = v.z; };
    SamplerState st = makeSampler(S, U, row, U.sampler);
    switch (op) {
    case ZL_KAT_HASH: o[0] = F(hash((uint32_t)B(a[0]))); break;
    case ZL_KAT_SOBOL: o[0] = F(sobolSample(S.sobol, (uint32_t)B(a[0]), B(a[1]))); break;
    case ZL_KAT_CUBEMAP_FACE: o[0] = F((uint32_t)cubemapFace(V3(a))); break;
    case ZL_KAT_BOXHIT: {
        // node id -> the record of face 0 that refers to it is not addressable directly, so the
        // bounds come from any face entry whose link table maps to `node`; KAT inputs pass the
        // THREADED index of face `cubemapFace(-dir)` instead (see tests): entry k of that face.
        Ray r = makeRay(V3(a + 1), V3(a + 4));
        const float4* nodes = S.nodes + (size_t)cubemapFace(-r.dir) * (size_t)S.bvhSize * 2;
        int k = B(a[0]);
        float4 lo, hi;
        unpackNodeRecord(nodes[2 * (size_t)k], nodes[2 * (size_t)k + 1], lo, hi);
        float t = 0.0f;
        bool h = boxHit<false>(f3(lo), f3(hi), prepareRay(r), t);
        o[0] = h ? 1.0f : 0.0f; o[1] = h ? t : 0.0f;
        break; }
    case ZL_KAT_TRIANGLE: {
        TriVerts tv = loadTriangle(S, B(a[0]));
        float t = 0.0f;
        bool h = intersectTriangle(tv.a, tv.b, tv.c, V3(a + 1), V3(a + 4), t);
        o[0] = h ? 1.0f : 0.0f; o[1] = h ? t : 0.0f;
        break; }
    case ZL_KAT_SURFACE: {
        SurfaceInfo s = triangleSurfaceInfo(S, B(a[0]), V3(a + 1));
        put3(o, s.ns); put3(o + 3, s.ng); o[6] = s.uv.x; o[7] = s.uv.y;
        break; }
    case ZL_KAT_CAMERA_RAY: {
        Ray r = thinLensCameraSampleRay(U, f2(a[0], a[1]), make_float4(a[2], a[3], a[4], a[5]));
        put3(o, r.ori); put3(o + 3, r.dir);
        break; }
    case ZL_KAT_CAMERA_II: {
        CameraIiSample c = thinLensCameraSampleIi(U, V3(a), f2(a[3], a[4]));
        put3(o, c.wi); put3(o + 3, c.Ii); o[6] = c.dist; o[7] = c.uv.x; o[8] = c.uv.y; o[9] = c.pdf;
        break; }
    case ZL_KAT_CAMERA_PDF: {
        CameraPdf c = thinLensCameraPdfIe(U, makeRay(V3(a), V3(a + 3)));
        o[0] = c.pdfPos; o[1] = c.pdfDir;
        break; }
    case ZL_KAT_BSDF_EVAL: {
        int mat = B(a[0]), tex = B(a[1]);
        uint32_t type = loadMaterialType(S, mat);
        BSDFParam p = loadMaterial(S, type, mat, tex, f2(a[2], a[3]));
        float4 r = materialBSDFAndPdf(type, p, V3(a + 4), V3(a + 7), V3(a + 10), (uint32_t)B(a[13]));
        o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
        break; }
    case ZL_KAT_BSDF_SAMPLE: {
        int mat = B(a[0]), tex = B(a[1]);
        uint32_t type = loadMaterialType(S, mat);
        BSDFParam p = loadMaterial(S, type, mat, tex, f2(a[2], a[3]));
        st.randSeed = (uint32_t)B(a[14]);
        BSDFSample s = materialSample(type, p, V3(a + 7), V3(a + 4), (uint32_t)B(a[10]), V3(a + 11), st);
        put3(o, s.wi); o[3] = s.pdf; put3(o + 4, s.bsdf); o[7] = s.eta; o[8] = F(s.flag);
        break; }
    case ZL_KAT_ENV_LE: { put3(o, envLe(S, U, V3(a))); o[3] = envPdfLi(S, U, V3(a)); break; }
    case ZL_KAT_ENV_SAMPLE: { float4 r = envSampleWi(S, U, make_float4(a[0], a[1], a[2], a[3])); o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; break; }
    case ZL_KAT_LIGHT_LE: {
        int id = B(a[0]);
        put3(o, lightLe(S, id, V3(a + 1), V3(a + 4))); o[3] = lightPdfLi(S, id, V3(a + 1), V3(a + 7));
        break; }
    case ZL_KAT_LIGHT_SAMPLE_LE: {
        LightLeSample l = lightSampleOneLe(S, B(a[0]), make_float4(a[1], a[2], a[3], a[4]));
        put3(o, l.ray.ori); put3(o + 3, l.ray.dir); put3(o + 6, l.Le); o[9] = l.pdfPos; o[10] = l.pdfDir;
        break; }
    case ZL_KAT_SAMPLE_LIGHT_ENV: {
        LightLiSample l = sampleLightAndEnv(S, U, V3(a), a[3], make_float4(a[4], a[5], a[6], a[7]));
        put3(o, l.wi); put3(o + 3, l.coef); o[6] = l.pdf;
        break; }
    case ZL_KAT_LIBM: {
        const float x = a[1], y = a[2];
        switch (B(a[0])) {
        case 0: o[0] = zl_sinf(x); break; case 1: o[0] = zl_cosf(x); break; case 2: o[0] = zl_atan2f(y, x); break;
        case 3: o[0] = zl_asinf(x); break; case 4: o[0] = zl_acosf(x); break; case 5: o[0] = zl_logf(x); break;
        case 6: o[0] = zl_powf(x, y); break; default: o[0] = zl_expf(x); break;
        }
        break; }
    default: break;
    }
}

#endif  // !ZL_INSTRUMENT

// film * scale with alpha = 1 into a staging buffer (util/img_copy_1x32f_4x32f.glsl + resultScale)
#ifndef ZL_INSTRUMENT
// Pipelined light tracer: a pass splats into its own zeroed buffer; when the pass is complete this kernel adds the buffer to the film (on the
// film stream, in pass order) and zeroes it for the pass after next.  Frame reads on the film stream then see whole passes only, without any
// pass having to wait for a read.
__global__ void mergeSplatKernel(float4* __restrict__ film, float4* __restrict__ splats, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = splats[i];
        if (a.x != 0.0f || a.y != 0.0f || a.z != 0.0f || a.w != 0.0f) {      // (also true for NaN components)
            float4 v = film[i];
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            film[i] = v;
            splats[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
    }
}
__global__ void resolveFilmKernel(const float4* __restrict__ film, float4* __restrict__ out, size_t n, float scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = film[i];
    out[i] = make_float4(v.x * scale, v.y * scale, v.z * scale, 1.0f);
}
// the same frame without its constant alpha: packed RGB, 12 bytes per pixel (read-back traffic -25 %)
__global__ void resolveFilmRgbKernel(const float4* __restrict__ film, float* __restrict__ out, size_t n, float scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = film[i];
    out[3 * i] = v.x * scale; out[3 * i + 1] = v.y * scale; out[3 * i + 2] = v.z * scale;
}

// Scene upload: resolve the vertex indices on the device (the host loop of Scene::createGLContext's consumers, Scene.cpp:197-243, gathered
// 96 bytes per triangle corner before a 600 MB copy).  One thread per triangle corner; uv rides in the w lanes (0 for light triangles,
// whose vertices lie past the texcoord array); an index outside the vertex array raises *bad and reads vertex 0.
__global__ void gatherTrianglesKernel(const float* __restrict__ vertices, const float* __restrict__ normals, const float* __restrict__ texcoords,
                                      const unsigned numVertices, const unsigned numTexcoords, const uint32_t* __restrict__ indices, const size_t corners,
                                      float4* __restrict__ triPos, float4* __restrict__ triNrm, int* __restrict__ bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= corners) return;
    uint32_t vi = indices[i];
    if (vi >= numVertices) { *bad = 1; vi = 0; }
    const float* v = vertices + 3 * (size_t)vi;
    const float* n = normals + 3 * (size_t)vi;
    float tu = 0.0f, tv = 0.0f;
    if (vi < numTexcoords) { tu = texcoords[2 * (size_t)vi]; tv = texcoords[2 * (size_t)vi + 1]; }
    triPos[i] = make_float4(v[0], v[1], v[2], tu);
    triNrm[i] = make_float4(n[0], n[1], n[2], tv);
}

// BVH::buildHitTable (BVH.cpp:298-346) on the device.  The reference walks the pre-order tree six times with a
// stack; here every node finds its own position in all six orderings by descending from the root: at an interior
// node x (left child L = x + 1, right child R = L + size(L)) face f visits L first iff cmp_f(centroid(L),
// centroid(R)) ('>' on the + faces, '<' on the - faces, strict: ties visit R first), the first child sits at
// pos + 1 and the second at pos + 1 + size(first).  One thread per node, ~tree-depth steps each, neighbouring
// threads share their path prefix (cache hits); the 32-byte records (packNodeRecord: bounds, prim | -1, pos + size) are
// written straight into the six threaded arrays.  Same float operations as AABB::centroid ((pMin + pMax) * 0.5f).
__global__ void threadMtbvhKernel(const float* __restrict__ bounds, const int* __restrict__ sizeIndices, const int n, float4* __restrict__ nodes) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int pos[6] = {0, 0, 0, 0, 0, 0};
    int x = 0;
    int sx = __ldg(sizeIndices);                       // size word of the current node
    while (x != t) {
        const int L = x + 1;
        int sL = __ldg(sizeIndices + L);
        if (sL < 0) sL = 1;                            // leaf: primIndex | 0x80000000
        const int R = L + sL;
        const int sR = sx - 1 - sL;
        const float* bl = bounds + 6 * (size_t)L;
        const float* br = bounds + 6 * (size_t)R;
        const bool goLeft = t < R;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float cl = (__ldg(bl + a) + __ldg(bl + 3 + a)) * 0.5f, cr = (__ldg(br + a) + __ldg(br + 3 + a)) * 0.5f;
            const bool lFirstPlus = cl > cr, lFirstMinus = cl < cr;
            pos[2 * a] += goLeft ? (lFirstPlus ? 1 : 1 + sR) : (lFirstPlus ? 1 + sL : 1);
            pos[2 * a + 1] += goLeft ? (lFirstMinus ? 1 : 1 + sR) : (lFirstMinus ? 1 + sL : 1);
        }
        x = goLeft ? L : R;
        sx = goLeft ? sL : sR;                         // (a leaf reached here ends the loop: x == t)
    }
    const int word = __ldg(sizeIndices + t);
    const bool leaf = word < 0;
    const int prim = leaf ? (word ^ (int)0x80000000) : -1;
    const int size = leaf ? 1 : word;
    const float* b = bounds + 6 * (size_t)t;
    const float3 lo = f3(__ldg(b), __ldg(b + 1), __ldg(b + 2));
    const float3 hi = f3(__ldg(b + 3), __ldg(b + 4), __ldg(b + 5));
#pragma unroll
    for (int f = 0; f < 6; f++) {
        float4* rec = nodes + 2 * ((size_t)f * n + pos[f]);
        packNodeRecord(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, prim, pos[f] + size, rec[0], rec[1]);
    }
}

// post_proc.glsl:12-59 — the display stage: scale, clamp to [0, 1e30], tone map (0 none, 1 filmic, 2 ACES), gamma 1/2.2.
// One thread per pixel, film row 0 = bottom (as uIn); outF is the reference's rgba32f result texture, out8 its
// GL_RGB / GL_UNSIGNED_BYTE read-back (Texture2D::readFromDevice, Texture.cpp:96-102: clamp to [0,1], x255, round).
ZL_DEV float3 postCalc(float3 x) {                      // post_proc.glsl:22-26
    const float A = 0.22f, B = 0.3f, C = 0.1f, D = 0.2f, E = 0.01f, F = 0.3f;
    return (x * (x * A + f3(B * C)) + f3(D * E)) / (x * (x * A + f3(B)) + f3(D * F)) - f3(E / F);
}
__global__ void postProcKernel(const float4* __restrict__ film, float4* __restrict__ outF, unsigned char* __restrict__ out8,
                               size_t n, float scale, int toneMapper) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = film[i];
    float3 color = f3(v.x * scale, v.y * scale, v.z * scale);
    color = f3(fminf(fmaxf(color.x, 0.0f), 1e30f), fminf(fmaxf(color.y, 0.0f), 1e30f), fminf(fmaxf(color.z, 0.0f), 1e30f));   // clamp(color, 0.0, 1e30)
    float3 mapped = color;
    if (toneMapper == 1) mapped = postCalc(color * 1.6f) / postCalc(f3(11.2f));                                              // filmic, :28-32
    else if (toneMapper == 2) mapped = (color * (color * 2.51f + f3(0.03f))) / (color * (color * 2.43f + f3(0.59f)) + f3(0.14f));   // ACES, :34-37
    const float g = 1.0f / 2.2f;
    mapped = f3(zl_powf(mapped.x, g), zl_powf(mapped.y, g), zl_powf(mapped.z, g));
    if (outF) outF[i] = make_float4(mapped.x, mapped.y, mapped.z, 1.0f);
    if (out8) {
        out8[3 * i + 0] = (unsigned char)rintf(fminf(fmaxf(mapped.x, 0.0f), 1.0f) * 255.0f);
        out8[3 * i + 1] = (unsigned char)rintf(fminf(fmaxf(mapped.y, 0.0f), 1.0f) * 255.0f);
        out8[3 * i + 2] = (unsigned char)rintf(fminf(fmaxf(mapped.z, 0.0f), 1.0f) * 255.0f);
    }
}
#endif

}  // namespace zl
