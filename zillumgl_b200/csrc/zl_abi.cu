// zl_abi.cu — the C ABI of include/zillum_cuda.h: scene upload (re-packing the reference's
// buffer-texture arrays into the 16-byte device layout of zl_scene.cuh), film objects, the
// four pass launches, explicit ray-set traversal, per-function KAT evaluation.
// Everything here is plumbing around the kernels in zl_kernels.cuh; there is no CPU
// fallback: without a CUDA device every entry point returns an error.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "zl_kernels.cuh"
#include "zl_bvh_build.cuh"
#include "zl_wavefront.cuh"

using namespace zl;

static thread_local std::string g_lastError;
static std::atomic<unsigned long long> g_launches{0};

static int fail(int code, const std::string& msg) { g_lastError = msg; return code; }
#define ZL_CK(call)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail((int)e_, std::string(#call) + ": " + cudaGetErrorString(e_));                 \
    } while (0)
#define ZL_LAUNCHED()                                                                                 \
    do {                                                                                              \
        g_launches++;                                                                                 \
        cudaError_t e_ = cudaGetLastError();                                                          \
        if (e_ != cudaSuccess) return fail((int)e_, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
    } while (0)

// ---- per-stage device timing (zl_stage_timing_*): CUDA events on the launch stream around each group of
// launches of a pass.  Off by default (two event records per group); bench.py switches it on for the
// roofline of the dominant kernel, outside its headline timed region.
struct StageTimer {
    struct Span { int stage; cudaEvent_t a, b; unsigned long long launches; };
    bool enabled = false;
    std::vector<Span> spans;
    void clear() { for (auto& s : spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); } spans.clear(); }
};
static StageTimer g_stageTimer;
struct StageScope {
    int stage; cudaStream_t stream; cudaEvent_t a = nullptr; unsigned long long l0 = 0;
    StageScope(int stage_, cudaStream_t stream_) : stage(stage_), stream(stream_) {
        if (!g_stageTimer.enabled) return;
        cudaEventCreate(&a);
        cudaEventRecord(a, stream);
        l0 = g_launches.load();
    }
    ~StageScope() {
        if (!a) return;
        cudaEvent_t b;
        cudaEventCreate(&b);
        cudaEventRecord(b, stream);
        g_stageTimer.spans.push_back({stage, a, b, g_launches.load() - l0});
    }
};

struct ZlScene {
    DScene d{};
    std::vector<void*> allocs;
    size_t totalBytes = 0, nodeBytes = 0;
    unsigned binMask = 0;          // material-type bins (materialBin) present in the scene: which shade kernels to launch
    int traceLoop = 0;             // trace kernel form chosen for this scene (WfOptions::loop = -1): 0 = one ray per lane, 5 = two rays per lane
    const float4* bvh2 = nullptr; int bvh2Depth = 0;      // BVH2 records (kept here; DScene::bvh2 is set per launch from the walk switch)
    double cudaInitMs = 0.0;     // one-time lazy loading of the device-build kernels, when this scene creation paid for it (zl_scene_cuda_init_ms)
    double bvhBuildMs = 0.0, mtbvhThreadMs = 0.0; int bvhLevels = 0;   // device-side scene preparation (0 when done on the host)
    ~ZlScene() { for (void* p : allocs) cudaFree(p); }
};
// wavefront workspace of a film (allocated on first use of variant 1): slot-indexed path state + queues
struct WfWorkspace {
    WfState st{};
    int gridTracePipelined = 0;     // grid of the default trace kernel when another pass is in flight next to this one (variant 2)
    int gridShadeDense = 0, gridResolveDense = 0;      // persistent grids of the 56-register stage kernels (zl_wavefront.cuh "Dense")
    void* block = nullptr;
    size_t bytes = 0;
    size_t capacity = 0;            // slots the arrays and queues can hold
    size_t histInts = 0;            // ints of the sort histogram block
    // overlapped pass schedule (launchWavefrontPathPass): resolve(b) and the shade kernels of the minor material types run on
    // side streams next to the main stream's shade / sort / trace
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t evFork = nullptr, evJoin[3] = {nullptr, nullptr, nullptr}, evTraced = nullptr, evResolved[2] = {nullptr, nullptr};
    bool resolvePending[2] = {false, false};
    // pipelined passes (variant 2): the stream this workspace's pass runs on, and "every resolve of that pass is done"
    cudaStream_t chain = nullptr;
    cudaEvent_t evPassResolved = nullptr;
    bool passInFlight = false;
    // pipelined light tracer: this workspace's pass splats here (film-sized, zero between passes); merged into the film on the film stream
    float4* splats = nullptr; size_t splatPixels = 0; cudaEvent_t evMerged = nullptr; bool mergePending = false;
    ~WfWorkspace() {
        if (chain) { cudaStreamSynchronize(chain); cudaStreamDestroy(chain); }
        if (evPassResolved) cudaEventDestroy(evPassResolved);
        if (splats) cudaFree(splats);
        if (evMerged) cudaEventDestroy(evMerged);
        for (auto& st_ : side) if (st_) { cudaStreamSynchronize(st_); cudaStreamDestroy(st_); }
        if (evFork) cudaEventDestroy(evFork);
        if (evTraced) cudaEventDestroy(evTraced);
        for (auto& e : evJoin) if (e) cudaEventDestroy(e);
        for (auto& e : evResolved) if (e) cudaEventDestroy(e);
    }
    int gridLightShade[kWfBins] = {0, 0, 0, 0, 0}, gridTripleShade[kWfBins] = {0, 0, 0, 0, 0}, gridTripleLightShade[kWfBins] = {0, 0, 0, 0, 0}, gridTripleResolve = 0;
    int gridShade[kWfBins] = {0, 0, 0, 0, 0}, gridTrace = 0, gridTraceSimple[3] = {0, 0, 0}, gridResolve = 0, sms = 148;
};
struct ZlFilm {
    float4* d = nullptr; float4* stage = nullptr; unsigned char* stage8 = nullptr; int w = 0, h = 0; bool owned = true; WfWorkspace* wf = nullptr;
    // pipelined passes (zl_launch_path_pass variant 2): second workspace, the film stream R that carries every film write and
    // read while passes are in flight, and the bookkeeping of zl_film_flush
    unsigned pathRing = 0;
    WfWorkspace* wf2 = nullptr; WfWorkspace* wf3 = nullptr; WfWorkspace* wf4 = nullptr;   /* wf3, wf4: path tracer with three / four passes in flight (pipeDepth) */ cudaStream_t filmStream = nullptr; cudaEvent_t evUser = nullptr, evTail = nullptr;
    bool pipeDirty = false; unsigned long long pipePasses = 0;
    bool readSincePass = false;     // a frame read was queued on the film stream after the last pipelined pass (splat passes must follow it)
    bool tripleHalf = false;        // variant-2 triple tracer: the camera pass of the current pass pair is launched, its light pass is not yet
    // zl_film_download_async: up to kDlSlots read-backs in flight, each with its own staging buffer (allocated on first use; FIFO: dlOldest .. dlOldest + dlPending - 1)
    static constexpr int kDlSlots = 4;
    struct Download { float4* stage = nullptr; cudaEvent_t evResolved = nullptr, evCopied = nullptr; void* pendingDst = nullptr; size_t pendingBytes = 0; };
    unsigned long long passesLaunched = 0;     // > 0: the D2H copy of a read-back is issued behind the NEXT pass launch (filmIssuePendingCopies)
    cudaStream_t copyStream = nullptr; Download dl[kDlSlots]; int dlOldest = 0, dlPending = 0;
    cudaEvent_t evSnap = nullptr, evSnapUser = nullptr;     // zl_film_snapshot_async
    // kernelVariant 3: one captured graph per pass kind (0 path, 1 light, 2 triple PT, 3 triple LPT), replayed on graphStream (graphPass)
    struct PassGraph { cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; cudaGraphNode_t setNode = nullptr; ZlRenderParams key{}; const void* scene = nullptr; unsigned binMask = 0;
                       std::string optKey; unsigned long long launches = 0; bool warm = false, failed = false; };
    PassGraph graphs[4]; cudaStream_t graphStream = nullptr; cudaEvent_t evGraphIn = nullptr, evGraphOut = nullptr; int* dPassCounters = nullptr;
};
// pipelined passes (variant 2, launchWavefrontPathPassPipelined): make `stream` wait for every pass in flight on the film; afterwards the film may be used from `stream` like any buffer
static int pipeFlush(ZlFilm* f, cudaStream_t stream) {
    if (!f || !f->pipeDirty) return 0;
    ZL_CK(cudaStreamWaitEvent(stream, f->evTail, 0));       // R is in order: the last pass's last resolve covers everything before it
    f->pipeDirty = false;
    return 0;
}
namespace zlc { struct DScene; int launchCountedPass(int kind, const DScene& S, const ZlRenderParams& U, float4* film, cudaStream_t stream); }
struct ZlRaySet {
    float4* rays = nullptr;     // 2 float4 per ray: {ori.xyz, tMax}, {dir.xyz, 0}
    int32_t* ids = nullptr; float* t = nullptr;
    size_t n = 0; int tileW = 0, tileH = 0;   // > 0: rays form a tileW x tileH pixel grid (row-major)
};

template <typename T, typename P>
static int upload(ZlScene* s, const std::vector<T>& host, P* dev) {
    *dev = nullptr;
    if (host.empty()) return 0;
    void* p = nullptr;
    ZL_CK(cudaMalloc(&p, host.size() * sizeof(T)));
    s->allocs.push_back(p);
    s->totalBytes += host.size() * sizeof(T);
    ZL_CK(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dev = (P)p;
    return 0;
}

static unsigned binMaskOf(const float* materials, int count) {
    unsigned mask = 0;
    for (int i = 0; i < count; i++) {
        uint32_t type;
        std::memcpy(&type, materials + 16 * (size_t)i + 13, 4);      // Material.h:32-53: texel 3 = {ior, type bits, pad, pad}
        mask |= 1u << ((type >= 1u && type <= 4u) ? type : 0u);
    }
    return mask;
}

extern "C" {

int zl_abi_version(void) { return ZL_ABI_VERSION; }
const char* zl_last_error_string(void) { return g_lastError.c_str(); }
int zl_device_count(int* count) { ZL_CK(cudaGetDeviceCount(count)); return 0; }
int zl_set_device(int device) { ZL_CK(cudaSetDevice(device)); return 0; }
int zl_device_synchronize(void) { ZL_CK(cudaDeviceSynchronize()); return 0; }
unsigned long long zl_launch_count(void) { return g_launches.load(); }

// The first device BVH build of a process pays for CUDA's lazy loading of a dozen kernels (the level kernels + CUB's scan): 0.4-0.7 s
// that round 1 reported as "444 ms to build 262 k triangles".  A 4-triangle build loads them; returns the milliseconds it took
// (sub-millisecond once loaded).
// A/B switch ZL_BVH2_WALK (default 0): 1 = pure rays walk the child-boxes-in-the-parent records with a short stack (traverseBvh2: same results,
// half the dependent fetches, but 11-19 % slower on B200: profiles/r2_trace_sweep.md); 0 = the threaded records
static bool bvh2WalkEnabled() {
    const char* e = std::getenv("ZL_BVH2_WALK");
    return e ? std::atoi(e) != 0 : false;
}
static double warmUpBvhBuild() {
    static bool done = false;
    if (done) return 0.0;
    done = true;
    const auto t0 = std::chrono::steady_clock::now();
    float4* tri = nullptr; float* b = nullptr; int* sz = nullptr;
    const int T = 4;
    if (cudaMalloc((void**)&tri, 3 * T * sizeof(float4)) == cudaSuccess && cudaMalloc((void**)&b, (2 * T - 1) * 6 * sizeof(float)) == cudaSuccess &&
        cudaMalloc((void**)&sz, (2 * T - 1) * sizeof(int)) == cudaSuccess) {
        float4 h[3 * T];
        for (int i = 0; i < T; i++) { h[3 * i] = make_float4((float)i, 0, 0, 0); h[3 * i + 1] = make_float4((float)i + 1, 0, 0, 0); h[3 * i + 2] = make_float4((float)i, 1, (float)i, 0); }
        cudaMemcpy(tri, h, sizeof h, cudaMemcpyHostToDevice);
        buildBvhOnDevice(tri, T, b, sz, nullptr);
    }
    cudaFree(tri); cudaFree(b); cudaFree(sz);
    cudaGetLastError();
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

int zl_scene_create(const ZlSceneDesc* desc, ZlScene** out) {
    if (!desc || !out) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_create: null argument");
    if (desc->numTriangles <= 0 || desc->bvhSize != 2 * desc->numTriangles - 1)
        return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_create: bvhSize must equal 2*numTriangles-1");
    if (!desc->vertices || !desc->normals || !desc->indices || !desc->materials || !desc->sobolMatrices)
        return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_create: missing required array");
    if (desc->bounds && !desc->hitTable && !desc->sizeIndices)
        return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_create: bounds given without hitTable or sizeIndices");
    if (desc->numVertices <= 0 || desc->numMaterials <= 0 || desc->objPrimCount < 0 || desc->numLightTriangles < 0 ||
        desc->objPrimCount + desc->numLightTriangles != desc->numTriangles)
        return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_create: counts are inconsistent (objPrimCount + numLightTriangles must equal numTriangles)");
    if (desc->objPrimCount > 0 && !desc->matTexIndices)
        return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_create: matTexIndices missing");
    if (desc->numLightTriangles > 0 && (!desc->lightPower || !desc->lightAlias || !desc->lightProb))
        return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_create: light tables missing");
    if (desc->numTextures > 0 && desc->texels && !desc->texUVScale)
        return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_create: texUVScale missing");
    if (desc->envMap && desc->envW > 0 && desc->envH > 0 && (!desc->envAlias || !desc->envAliasProb))
        return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_create: environment alias tables missing");
    // (vertex indices are range-checked by the device-side gather below)
    for (int i = 0; i < desc->objPrimCount; i++) {
        const int m = desc->matTexIndices[i] & 0xffff, t = desc->matTexIndices[i] >> 16;
        if (m >= desc->numMaterials || t >= desc->numTextures) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_create: material / texture index out of range");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(ZL_ERR_NO_DEVICE, "zl_scene_create: no CUDA device");
    auto* s = new ZlScene();
    DScene& d = s->d;
    const ZlSceneDesc& h = *desc;
    const size_t n = (size_t)h.bvhSize, T = (size_t)h.numTriangles;
    int rc = 0;
    {   // per-triangle gathered positions / normals, uv in the w lanes: the indexed arrays go up as they are (176 MB instead of 604 MB at
        // 6.29 M triangles) and gatherTrianglesKernel resolves the indices on the device, range-checking them on the way
        const size_t V = (size_t)h.numVertices, TC = h.texcoords ? (size_t)h.numTexcoords : 0;
        float *dv = nullptr, *dn = nullptr, *dt = nullptr; uint32_t* di = nullptr; int* dBad = nullptr;
        void *pPos = nullptr, *pNrm = nullptr;
        cudaError_t e = cudaMalloc(&pPos, 3 * T * sizeof(float4));
        if (e == cudaSuccess) { s->allocs.push_back(pPos); e = cudaMalloc(&pNrm, 3 * T * sizeof(float4)); }
        if (e == cudaSuccess) { s->allocs.push_back(pNrm); s->totalBytes += 6 * T * sizeof(float4); }
        if (e == cudaSuccess) e = cudaMalloc((void**)&dv, V * 3 * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc((void**)&dn, V * 3 * sizeof(float));
        if (e == cudaSuccess && TC) e = cudaMalloc((void**)&dt, TC * 2 * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc((void**)&di, 3 * T * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc((void**)&dBad, sizeof(int));
        if (e == cudaSuccess) e = cudaMemcpy(dv, h.vertices, V * 3 * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dn, h.normals, V * 3 * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess && TC) e = cudaMemcpy(dt, h.texcoords, TC * 2 * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(di, h.indices, 3 * T * sizeof(uint32_t), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemset(dBad, 0, sizeof(int));
        int bad = 0;
        if (e == cudaSuccess) {
            gatherTrianglesKernel<<<(unsigned)((3 * T + 255) / 256), 256>>>(dv, dn, dt, (unsigned)V, (unsigned)TC, di, 3 * T, (float4*)pPos, (float4*)pNrm, dBad);
            g_launches++;
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpy(&bad, dBad, sizeof(int), cudaMemcpyDeviceToHost);
        }
        cudaFree(dv); cudaFree(dn); cudaFree(dt); cudaFree(di); cudaFree(dBad);
        if (e != cudaSuccess) { delete s; return fail((int)e, std::string("zl_scene_create: triangle upload: ") + cudaGetErrorString(e)); }
        if (bad) { delete s; return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_create: vertex index out of range"); }
        d.triPos = (const float4*)pPos; d.triNrm = (const float4*)pNrm;
    }
    {   // threaded node records, one face at a time (bounded staging memory)
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, (6 * n + 1) * 2 * sizeof(float4));   // + one pad record: the look-ahead loads of traverseSpec may read entry n of the last face
        if (e != cudaSuccess) { delete s; return fail((int)e, "zl_scene_create: cudaMalloc(nodes)"); }
        cudaMemset((float4*)p + 6 * n * 2, 0, 2 * sizeof(float4));
        s->allocs.push_back(p);
        s->nodeBytes = 6 * n * 2 * sizeof(float4);
        s->totalBytes += s->nodeBytes;
        // the builder's pre-order tree on the device: input of the MTBVH threading kernel and of the BVH2 records
        const bool hostTable = h.hitTable && h.bounds;
        float* dBounds = nullptr; int* dSizes = nullptr;
        e = cudaMalloc((void**)&dBounds, n * 6 * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc((void**)&dSizes, n * sizeof(int));
        if (h.bounds) {
            if (e == cudaSuccess) e = cudaMemcpy(dBounds, h.bounds, n * 6 * sizeof(float), cudaMemcpyHostToDevice);
            if (h.sizeIndices) {
                if (e == cudaSuccess) e = cudaMemcpy(dSizes, h.sizeIndices, n * sizeof(int), cudaMemcpyHostToDevice);
            } else {    // only the hit table came: entry j of any face = (node, prim | -1, j + subtree size) (BVH.cpp:324-326)
                std::vector<int> sizes(n);
                for (size_t j = 0; j < n; j++) {
                    const int node = h.hitTable[3 * j], prim = h.hitTable[3 * j + 1], miss = h.hitTable[3 * j + 2];
                    if ((size_t)(unsigned)node >= n) { cudaFree(dBounds); cudaFree(dSizes); delete s; return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_create: hit table names a node outside the tree"); }
                    sizes[node] = prim >= 0 ? (int)((unsigned)prim | 0x80000000u) : miss - (int)j;
                }
                if (e == cudaSuccess) e = cudaMemcpy(dSizes, sizes.data(), n * sizeof(int), cudaMemcpyHostToDevice);
            }
        } else if (e == cudaSuccess) {      // no tree at all: BVH::build on the device (zl_bvh_build.cuh)
            s->cudaInitMs = warmUpBvhBuild();         // first use in a process: CUDA loads the build kernels' modules lazily (hundreds of ms); keep that out of the build time
            const auto t0 = std::chrono::steady_clock::now();
            e = buildBvhOnDevice(d.triPos, (int)T, dBounds, dSizes, &s->bvhLevels);
            s->bvhBuildMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            g_launches += 8ull * (unsigned long long)std::max(s->bvhLevels, 1);
        }
        if (e == cudaSuccess && !hostTable) {   // thread the six orderings on the device (threadMtbvhKernel), from the host's tree or the one built here
            const auto t0 = std::chrono::steady_clock::now();
            threadMtbvhKernel<<<(unsigned)((n + 127) / 128), 128>>>(dBounds, dSizes, (int)n, (float4*)p);
            g_launches++;
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            s->mtbvhThreadMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
        d.bvh2 = nullptr;
        // child-boxes-in-the-parent records of the same tree (traverseBvh2): T - 1 interior nodes x 64 bytes.  The walk is an A/B switch that is off
        // by default (profiles/r2_trace_sweep.md): large scenes get the records only when it is switched on at creation (0.4 GB at 6.3 M triangles),
        // small ones always, so that ZL_BVH2_WALK can be toggled per launch on an existing scene
        if (e == cudaSuccess && T >= 2 && (bvh2WalkEnabled() || T <= ((size_t)1 << 20))) {
            void* q = nullptr;
            e = cudaMalloc(&q, (T - 1) * 4 * sizeof(float4));
            if (e == cudaSuccess) {
                s->allocs.push_back(q);
                int maxDepth = 0;
                e = buildBvh2OnDevice(dBounds, dSizes, (int)n, (float4*)q, &maxDepth);
                g_launches += 3;
                s->bvh2Depth = maxDepth;
                float rb[6];
                if (e == cudaSuccess) e = cudaMemcpy(rb, dBounds, sizeof rb, cudaMemcpyDeviceToHost);
                d.rootLo = make_float3(rb[0], rb[1], rb[2]); d.rootHi = make_float3(rb[3], rb[4], rb[5]);
                // a lane's stack holds at most one remembered sibling per level: deeper trees (chains of coincident centroids) keep the threaded walk
                if (e == cudaSuccess && maxDepth < kBvh2Stack) { s->bvh2 = (const float4*)q; s->totalBytes += (T - 1) * 4 * sizeof(float4); }
            }
        }
        cudaFree(dBounds); cudaFree(dSizes);
        if (e != cudaSuccess) { delete s; return fail((int)e, std::string("zl_scene_create: device tree preparation: ") + cudaGetErrorString(e)); }
        std::vector<float4> stage(hostTable ? n * 2 : 0);
        for (int f = 0; f < 6 && hostTable; f++) {
            const int32_t* table = h.hitTable + (size_t)f * n * 3;
            for (size_t k = 0; k < n; k++) {
                int node = table[3 * k], prim = table[3 * k + 1], miss = table[3 * k + 2];
                if ((size_t)(unsigned)node >= n) { delete s; return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_create: hit table names a node outside the tree"); }
                const float* b = h.bounds + 6 * (size_t)node;
                packNodeRecord(b[0], b[1], b[2], b[3], b[4], b[5], prim, miss, stage[2 * k], stage[2 * k + 1]);
            }
            e = cudaMemcpy((float4*)p + (size_t)f * n * 2, stage.data(), n * 2 * sizeof(float4), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { delete s; return fail((int)e, "zl_scene_create: cudaMemcpy(nodes)"); }
        }
        d.nodes = (const float4*)p;
        {   // compact copy of the top levels of the six orderings for the shared-memory staged walk (zl_traverse.cuh, buildStagedTopKernel)
            void* t = nullptr;
            e = cudaMalloc(&t, kTopBytes);
            if (e != cudaSuccess) { delete s; return fail((int)e, "zl_scene_create: cudaMalloc(top)"); }
            s->allocs.push_back(t);
            buildStagedTopKernel<<<6, 32>>>(d.nodes, (int)n, (float4*)t);
            g_launches++;
            if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) { delete s; return fail((int)e, std::string("zl_scene_create: staged top: ") + cudaGetErrorString(e)); }
            d.top = (const float4*)t;
        }
    }
    {
        std::vector<int> mt(h.matTexIndices, h.matTexIndices + (h.objPrimCount > 0 ? h.objPrimCount : 0));
        std::vector<float4> mats(4 * (size_t)h.numMaterials);
        std::memcpy(mats.data(), h.materials, mats.size() * sizeof(float4));
        std::vector<float4> lpp((size_t)h.numLightTriangles);
        std::vector<int> la((size_t)h.numLightTriangles);
        for (int i = 0; i < h.numLightTriangles; i++) {
            lpp[i] = make_float4(h.lightPower[3 * i], h.lightPower[3 * i + 1], h.lightPower[3 * i + 2], h.lightProb[i]);
            la[i] = h.lightAlias[i];
        }
        if ((rc = upload(s, mt, &d.matTex)) || (rc = upload(s, mats, &d.materials)) || (rc = upload(s, lpp, &d.lightPowProb)) ||
            (rc = upload(s, la, &d.lightAlias))) { delete s; return rc; }
    }
    {   // albedo layers (sRGB8 -> uchar4) + decode LUT (GL_SRGB, evaluated in double)
        std::vector<uchar4> tex;
        std::vector<float2> scale;
        if (h.numTextures > 0 && h.texels) {
            size_t px = (size_t)h.numTextures * h.texMaxW * h.texMaxH;
            tex.resize(px);
            for (size_t i = 0; i < px; i++) tex[i] = make_uchar4(h.texels[3 * i], h.texels[3 * i + 1], h.texels[3 * i + 2], 255);
            scale.resize(h.numTextures);
            for (int i = 0; i < h.numTextures; i++) scale[i] = make_float2(h.texUVScale[2 * i], h.texUVScale[2 * i + 1]);
        }
        std::vector<float> lut(256);
        for (int i = 0; i < 256; i++) {
            double c = i / 255.0;
            lut[i] = (float)((c <= 0.04045) ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4));
        }
        if ((rc = upload(s, tex, &d.texels)) || (rc = upload(s, scale, &d.texScale)) || (rc = upload(s, lut, &d.srgbLut))) { delete s; return rc; }
    }
    {   // environment map -> RGB16F, alias/prob packed as int2
        std::vector<ushort4> env;
        std::vector<int2> ea;
        if (h.envMap && h.envW > 0 && h.envH > 0) {
            d.envW = h.envW; d.envH = h.envH; d.envSum = h.envSum;
            size_t px = (size_t)h.envW * h.envH;
            env.resize(px);
            for (size_t i = 0; i < px; i++) {
                __half r = __float2half_rn(h.envMap[3 * i]), g = __float2half_rn(h.envMap[3 * i + 1]), b = __float2half_rn(h.envMap[3 * i + 2]);
                env[i] = make_ushort4(__half_as_ushort(r), __half_as_ushort(g), __half_as_ushort(b), 0);
            }
            size_t ne = (size_t)(h.envW + 1) * h.envH;
            ea.resize(ne);
            for (size_t i = 0; i < ne; i++) { int pb; std::memcpy(&pb, &h.envAliasProb[i], 4); ea[i] = make_int2(h.envAlias[i], pb); }
        } else {   // the reference always binds a map; a missing one behaves as 1x1 black
            d.envW = d.envH = 1; d.envSum = 0.0f;
            env.assign(1, make_ushort4(0, 0, 0, 0));
            int one; float onef = 1.0f; std::memcpy(&one, &onef, 4);
            ea.assign(2, make_int2(0, one));
        }
        if ((rc = upload(s, env, &d.env)) || (rc = upload(s, ea, &d.envAlias))) { delete s; return rc; }
    }
    {
        std::vector<float2> noise;
        if (h.noise && h.noiseW > 0 && h.noiseH > 0) {
            d.noiseW = h.noiseW; d.noiseH = h.noiseH;
            noise.resize((size_t)h.noiseW * h.noiseH);
            std::memcpy(noise.data(), h.noise, noise.size() * sizeof(float2));
        } else { d.noiseW = d.noiseH = 1; noise.assign(1, make_float2(0.5f, 0.5f)); }
        std::vector<uint32_t> sob(h.sobolMatrices, h.sobolMatrices + 256 * 32);
        if ((rc = upload(s, noise, &d.noise)) || (rc = upload(s, sob, &d.sobol))) { delete s; return rc; }
    }
    d.bvhSize = h.bvhSize; d.numTriangles = h.numTriangles; d.objPrimCount = h.objPrimCount;
    d.numLightTriangles = h.numLightTriangles; d.numMaterials = h.numMaterials;
    d.numTextures = (h.numTextures > 0 && h.texels) ? h.numTextures : 0; d.texMaxW = h.texMaxW; d.texMaxH = h.texMaxH;
    d.lightSum = h.lightSum;
    // octant-specialised packed walks when the node records exceed the L2 (the step waits on DRAM / far L2 and issue slots matter);
    // the scalar walk with its shorter dependent chain when one face's records are L2-resident (profiles/r2_trace_sweep.md)
    d.octantWalk = (n * 32 <= (size_t)64 << 20) ? 2 : 1;
    // Two rays per lane (wfTraceDualKernel) when all six orderings of the table fit in a corner of the L1 (Cornell-class scenes): there the
    // walk is issue-bound at 10-11 live lanes per instruction (profiles/r2_pass_full_cornell.csv), and a lane that carries two rays idles less
    // (trace stage 1.71 -> 1.39 ms, profiles/r2_sweep_loops_cornell.json).  Larger scenes wait on L2 / DRAM and lose with it (r1_sweep_dual_*).
    s->traceLoop = (6 * n * 32 <= (size_t)64 << 10) ? 5 : 0;
    d.nodePolicy = 0; d.statePolicy = 0;     // set per launch from WfOptions (ZL_NODE_POLICY / ZL_STATE_POLICY)
    d.bvh2 = bvh2WalkEnabled() ? s->bvh2 : nullptr;
    s->binMask = binMaskOf(h.materials, h.numMaterials);
    *out = s;
    return 0;
}

int zl_scene_destroy(ZlScene* scene) { delete scene; return 0; }


int zl_scene_update_materials(ZlScene* scene, int first, int count, const float* materials) {
    cudaDeviceSynchronize();      // passes may be in flight on internal streams (variant 2); the materials are rewritten in place
    if (!scene || !materials || first < 0 || count < 0 || first + count > scene->d.numMaterials)
        return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_update_materials: null argument or range out of bounds");
    ZL_CK(cudaMemcpy((void*)(scene->d.materials + 4 * (size_t)first), materials, (size_t)count * 64, cudaMemcpyHostToDevice));
    scene->binMask |= binMaskOf(materials, count);                     // a type may have been added; stale bins only cost an empty launch
    return 0;
}

int zl_build_bvh(const float* vertices, int numVertices, const uint32_t* indices, int numTriangles, float* boundsOut, int32_t* sizeIndicesOut, int* levelsOut) {
    if (!vertices || !indices || !boundsOut || !sizeIndicesOut || numTriangles <= 0 || numVertices <= 0)
        return fail(ZL_ERR_INVALID_ARGUMENT, "zl_build_bvh: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(ZL_ERR_NO_DEVICE, "zl_build_bvh: no CUDA device");
    const size_t T = (size_t)numTriangles, n = 2 * T - 1;
    std::vector<float4> pos(3 * T);
    for (size_t t = 0; t < 3 * T; t++) {
        if (indices[t] >= (uint32_t)numVertices) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_build_bvh: vertex index out of range");
        const float* v = vertices + 3 * (size_t)indices[t];
        pos[t] = make_float4(v[0], v[1], v[2], 0.0f);
    }
    float4* dPos = nullptr; float* dBounds = nullptr; int* dSizes = nullptr;
    cudaError_t e = cudaMalloc((void**)&dPos, pos.size() * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dBounds, n * 6 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dSizes, n * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(dPos, pos.data(), pos.size() * sizeof(float4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { warmUpBvhBuild(); e = buildBvhOnDevice(dPos, numTriangles, dBounds, dSizes, levelsOut); }
    if (e == cudaSuccess) e = cudaMemcpy(boundsOut, dBounds, n * 6 * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(sizeIndicesOut, dSizes, n * sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(dPos); cudaFree(dBounds); cudaFree(dSizes);
    if (e != cudaSuccess) return fail((int)e, std::string("zl_build_bvh: ") + cudaGetErrorString(e));
    return 0;
}
int zl_scene_read_nodes(const ZlScene* scene, int face, size_t first, size_t count, float* boundsOut, int32_t* linksOut) {
    if (!scene || !boundsOut || !linksOut || face < 0 || face > 5) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_read_nodes: bad argument");
    const size_t n = (size_t)scene->d.bvhSize;
    if (first + count > n) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_read_nodes: range exceeds bvhSize");
    std::vector<float4> rec(2 * count);
    ZL_CK(cudaMemcpy(rec.data(), scene->d.nodes + 2 * ((size_t)face * n + first), rec.size() * sizeof(float4), cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < count; k++) {
        float4 lo, hi;
        unpackNodeRecord(rec[2 * k], rec[2 * k + 1], lo, hi);
        float* b = boundsOut + 6 * k;
        b[0] = lo.x; b[1] = lo.y; b[2] = lo.z; b[3] = hi.x; b[4] = hi.y; b[5] = hi.z;
        std::memcpy(linksOut + 2 * k, &lo.w, 4); std::memcpy(linksOut + 2 * k + 1, &hi.w, 4);
    }
    return 0;
}
double zl_scene_cuda_init_ms(const ZlScene* scene) { return scene ? scene->cudaInitMs : 0.0; }
int zl_scene_prep_times(const ZlScene* scene, double* bvhBuildMs, double* mtbvhThreadMs, int* bvhLevels) {
    if (!scene) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_prep_times: null scene");
    if (bvhBuildMs) *bvhBuildMs = scene->bvhBuildMs;
    if (mtbvhThreadMs) *mtbvhThreadMs = scene->mtbvhThreadMs;
    if (bvhLevels) *bvhLevels = scene->bvhLevels;
    return 0;
}
int zl_scene_memory(const ZlScene* scene, size_t* totalBytes, size_t* nodeBytes) {
    if (!scene) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_scene_memory: null scene");
    if (totalBytes) *totalBytes = scene->totalBytes;
    if (nodeBytes) *nodeBytes = scene->nodeBytes;
    return 0;
}

// ---- film ----
int zl_film_create(int width, int height, ZlFilm** out) {
    if (width <= 0 || height <= 0 || !out) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_film_create: bad size");
    auto* f = new ZlFilm();
    f->w = width; f->h = height; f->owned = true;
    cudaError_t e = cudaMalloc((void**)&f->d, (size_t)width * height * sizeof(float4));
    if (e != cudaSuccess) { delete f; return fail((int)e, std::string("zl_film_create: ") + cudaGetErrorString(e)); }
    e = cudaMemset(f->d, 0, (size_t)width * height * sizeof(float4));
    if (e != cudaSuccess) { cudaFree(f->d); delete f; return fail((int)e, "zl_film_create: cudaMemset"); }
    *out = f;
    return 0;
}
int zl_film_create_external(int width, int height, void* devicePtr, ZlFilm** out) {
    if (width <= 0 || height <= 0 || !devicePtr || !out) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_film_create_external: bad argument");
    auto* f = new ZlFilm();
    f->w = width; f->h = height; f->owned = false; f->d = (float4*)devicePtr;
    *out = f;
    return 0;
}
int zl_film_destroy(ZlFilm* film) {
    if (film && (film->pipeDirty || film->wf2)) cudaDeviceSynchronize();
    if (film && film->wf2) { cudaFree(film->wf2->block); delete film->wf2; }
    if (film && film->wf3) { cudaFree(film->wf3->block); delete film->wf3; }
    if (film && film->wf4) { cudaFree(film->wf4->block); delete film->wf4; }
    if (film && film->filmStream) { cudaStreamDestroy(film->filmStream); cudaEventDestroy(film->evUser); cudaEventDestroy(film->evTail); }
    if (film && film->owned && film->d) cudaFree(film->d);
    if (film && film->evSnap) { cudaEventDestroy(film->evSnap); cudaEventDestroy(film->evSnapUser); }
    if (film && film->graphStream) {
        cudaStreamSynchronize(film->graphStream);
        for (auto& g : film->graphs) if (g.exec) { cudaGraphExecDestroy(g.exec); cudaGraphDestroy(g.graph); }
        cudaStreamDestroy(film->graphStream); cudaEventDestroy(film->evGraphIn); cudaEventDestroy(film->evGraphOut); cudaFree(film->dPassCounters);
    }
    if (film && film->stage) cudaFree(film->stage);
    if (film && film->stage8) cudaFree(film->stage8);
    if (film && film->wf) { cudaFree(film->wf->block); delete film->wf; }
    if (film && film->copyStream) {
        cudaStreamSynchronize(film->copyStream); cudaStreamDestroy(film->copyStream);
        for (auto& d : film->dl) { if (d.stage) cudaFree(d.stage); if (d.evResolved) cudaEventDestroy(d.evResolved); if (d.evCopied) cudaEventDestroy(d.evCopied); }
    }
    delete film;
    return 0;
}
int zl_film_flush(ZlFilm* film, void* stream) {
    if (!film) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_film_flush: null film");
    return pipeFlush(film, (cudaStream_t)stream);
}
int zl_film_clear(ZlFilm* film, void* stream) {
    if (!film) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_film_clear: null film");
    if (int rc = pipeFlush(film, (cudaStream_t)stream)) return rc;
    ZL_CK(cudaMemsetAsync(film->d, 0, (size_t)film->w * film->h * sizeof(float4), (cudaStream_t)stream));
    return 0;
}
void* zl_film_device_ptr(ZlFilm* film) { return film ? film->d : nullptr; }
int zl_film_download(ZlFilm* film, float scale, float* rgbaHost, void* stream) {
    if (!film || !rgbaHost) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_film_download: null argument");
    if (int rc = pipeFlush(film, (cudaStream_t)stream)) return rc;
    size_t n = (size_t)film->w * film->h;
    if (!film->stage) ZL_CK(cudaMalloc((void**)&film->stage, n * sizeof(float4)));
    resolveFilmKernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(film->d, film->stage, n, scale);
    ZL_LAUNCHED();
    ZL_CK(cudaMemcpyAsync(rgbaHost, film->stage, n * sizeof(float4), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    ZL_CK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
int zl_film_postprocess(ZlFilm* film, float resultScale, int toneMapper, float* rgbaHost, unsigned char* rgb8Host, void* stream) {
    if (!film || (!rgbaHost && !rgb8Host)) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_film_postprocess: null argument");
    if (toneMapper < 0 || toneMapper > 2) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_film_postprocess: toneMapper must be 0 (none), 1 (filmic) or 2 (ACES)");
    const size_t n = (size_t)film->w * film->h;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = pipeFlush(film, st)) return rc;
    if (!film->stage) ZL_CK(cudaMalloc((void**)&film->stage, n * sizeof(float4)));
    if (rgb8Host && !film->stage8) ZL_CK(cudaMalloc((void**)&film->stage8, n * 3));
    postProcKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(film->d, rgbaHost ? film->stage : nullptr, rgb8Host ? film->stage8 : nullptr, n, resultScale, toneMapper);
    ZL_LAUNCHED();
    if (rgbaHost) ZL_CK(cudaMemcpyAsync(rgbaHost, film->stage, n * sizeof(float4), cudaMemcpyDeviceToHost, st));
    if (rgb8Host) ZL_CK(cudaMemcpyAsync(rgb8Host, film->stage8, n * 3, cudaMemcpyDeviceToHost, st));
    ZL_CK(cudaStreamSynchronize(st));
    return 0;
}
// issue the D2H copies of the read-backs whose frame is resolved into its staging buffer but not yet on its way (FIFO order)
static int filmIssuePendingCopies(ZlFilm* film) {
    for (int k = 0; k < film->dlPending; k++) {
        ZlFilm::Download& d = film->dl[(film->dlOldest + k) % ZlFilm::kDlSlots];
        if (!d.pendingDst) continue;
        ZL_CK(cudaStreamWaitEvent(film->copyStream, d.evResolved, 0));
        ZL_CK(cudaMemcpyAsync(d.pendingDst, d.stage, d.pendingBytes, cudaMemcpyDeviceToHost, film->copyStream));
        ZL_CK(cudaEventRecord(d.evCopied, film->copyStream));
        d.pendingDst = nullptr;
    }
    return 0;
}
static int filmDownloadAsync(ZlFilm* film, float scale, float* rgbaHostPinned, void* stream, int channels);
int zl_film_download_async(ZlFilm* film, float scale, float* rgbaHostPinned, void* stream) { return filmDownloadAsync(film, scale, rgbaHostPinned, stream, 4); }
int zl_film_download_rgb_async(ZlFilm* film, float scale, float* rgbHostPinned, void* stream) { return filmDownloadAsync(film, scale, rgbHostPinned, stream, 3); }
static int filmDownloadAsync(ZlFilm* film, float scale, float* rgbaHostPinned, void* stream, int channels) {
    if (!film || !rgbaHostPinned) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_film_download_async: null argument");
    const size_t n = (size_t)film->w * film->h;
    // with pipelined passes in flight the frame is resolved on the film stream, behind the resolves of the passes launched so
    // far and ahead of those launched later: a consistent snapshot that does not hold the next pass back
    cudaStream_t st = film->pipeDirty ? film->filmStream : (cudaStream_t)stream;
    if (!film->copyStream) ZL_CK(cudaStreamCreateWithFlags(&film->copyStream, cudaStreamNonBlocking));
    if (int rc = filmIssuePendingCopies(film)) return rc;           // (a slot that is taken over below must have its copy on the way)
    if (film->dlPending == 0) film->dlOldest = 0;                   // nothing in flight: start the ring over (a film with one read-back at a time uses one staging buffer)
    const bool reuse = film->dlPending == ZlFilm::kDlSlots;         // one read-back more than there are slots takes over the oldest slot
    ZlFilm::Download& d = film->dl[reuse ? film->dlOldest : (film->dlOldest + film->dlPending) % ZlFilm::kDlSlots];
    if (!d.stage) {
        ZL_CK(cudaMalloc((void**)&d.stage, n * sizeof(float4)));
        ZL_CK(cudaEventCreateWithFlags(&d.evResolved, cudaEventDisableTiming));
        ZL_CK(cudaEventCreateWithFlags(&d.evCopied, cudaEventDisableTiming));
    }
    // ZL_DEBUG_DOWNLOAD_TIMING=1: host time of each call below, printed every 16 frames (a read-back that blocks the host shows up here)
    static const bool dbg = std::getenv("ZL_DEBUG_DOWNLOAD_TIMING") != nullptr;
    static double acc[5] = {0, 0, 0, 0, 0}; static int accN = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t0 = now();
    if (reuse) { ZL_CK(cudaStreamWaitEvent(st, d.evCopied, 0)); film->dlOldest = (film->dlOldest + 1) % ZlFilm::kDlSlots; film->dlPending = ZlFilm::kDlSlots - 1; }   // its staging buffer is still being read
    if (channels == 4) resolveFilmKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(film->d, d.stage, n, scale);
    else resolveFilmRgbKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(film->d, (float*)d.stage, n, scale);
    ZL_LAUNCHED();
    const auto t1 = now();
    ZL_CK(cudaEventRecord(d.evResolved, st));
    if (film->pipeDirty) { ZL_CK(cudaEventRecord(film->evTail, st)); film->readSincePass = true; }      // a later flush also waits for this read of the film
    const auto t2 = now();
    // The copy itself.  On a film that passes are launched on, the cudaMemcpyAsync call is issued behind the NEXT pass launch (or by the wait):
    // on some hosts of the pod the call does not return before the copy has completed (0.03 ms on most boxes, 2 ms on others, same
    // binaries, page-locked destination) — issued right here it kept the host from queueing the next pass while this one was still
    // running, and the GPU idled for the length of the copy every step.  Behind the next launch it costs nothing either way.
    d.pendingDst = rgbaHostPinned; d.pendingBytes = n * sizeof(float) * channels;
    const auto t3 = now();
    film->dlPending++;
    if (film->passesLaunched == 0) { if (int rc = filmIssuePendingCopies(film)) return rc; }
    const auto t4 = now();
    const auto t5 = t4;
    if (dbg) {
        acc[0] += ms(t0, t1); acc[1] += ms(t1, t2); acc[2] += ms(t2, t3); acc[3] += ms(t3, t4); acc[4] += ms(t4, t5);
        if (++accN % 16 == 0) {
            cudaPointerAttributes pa{};
            cudaPointerGetAttributes(&pa, rgbaHostPinned);
            std::fprintf(stderr, "[zl download timing] per frame ms: launch %.3f, record %.3f, wait-event %.3f, memcpyAsync %.3f, record %.3f; host pointer type %d (1 = page-locked host)\n",
                         acc[0] / 16, acc[1] / 16, acc[2] / 16, acc[3] / 16, acc[4] / 16, (int)pa.type);
            for (double& a : acc) a = 0;
        }
    }
    return 0;
}
// Device-side snapshot of the film (the input of a reduce-before-copy frame path over several GPUs): film -> dstDevice (w*h float4) as
// ONE consistent state, i.e. behind the resolves of every pass launched so far and ahead of those launched later (on the film stream
// while pipelined passes are in flight, like the frame read-backs), after everything already queued on `stream` (the previous consumer
// of dstDevice); `stream` then waits for the copy, so the caller may hand dstDevice to a collective on `stream`.
int zl_film_snapshot_async(ZlFilm* film, void* dstDevice, void* stream) {
    if (!film || !dstDevice) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_film_snapshot_async: null argument");
    const size_t bytes = (size_t)film->w * film->h * sizeof(float4);
    cudaStream_t user = (cudaStream_t)stream;
    if (!film->pipeDirty) { ZL_CK(cudaMemcpyAsync(dstDevice, film->d, bytes, cudaMemcpyDeviceToDevice, user)); return 0; }
    cudaStream_t st = film->filmStream;
    if (!film->evSnap) { ZL_CK(cudaEventCreateWithFlags(&film->evSnap, cudaEventDisableTiming)); ZL_CK(cudaEventCreateWithFlags(&film->evSnapUser, cudaEventDisableTiming)); }
    ZL_CK(cudaEventRecord(film->evSnapUser, user));
    ZL_CK(cudaStreamWaitEvent(st, film->evSnapUser, 0));
    ZL_CK(cudaMemcpyAsync(dstDevice, film->d, bytes, cudaMemcpyDeviceToDevice, st));
    ZL_CK(cudaEventRecord(film->evSnap, st));
    ZL_CK(cudaEventRecord(film->evTail, st)); film->readSincePass = true;      // a later flush also waits for this read of the film
    ZL_CK(cudaStreamWaitEvent(user, film->evSnap, 0));
    return 0;
}
int zl_film_download_wait(ZlFilm* film) {
    if (!film) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_film_download_wait: null film");
    if (film->dlPending > 0) {                                      // the oldest read-back in flight
        if (int rc = filmIssuePendingCopies(film)) return rc;
        ZL_CK(cudaEventSynchronize(film->dl[film->dlOldest].evCopied));
        film->dlOldest = (film->dlOldest + 1) % ZlFilm::kDlSlots;
        film->dlPending--;
    }
    return 0;
}
int zl_film_allreduce(ZlFilm* film, void* ncclComm, void* stream) {
    // ncclAllReduce(sendbuff, recvbuff, count, ncclFloat32 = 7, ncclSum = 0, comm, stream), resolved at run time
    typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
    static AllReduceFn fn = nullptr;
    if (!fn) {
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (lib) fn = (AllReduceFn)dlsym(lib, "ncclAllReduce");
        if (!fn) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_film_allreduce: libnccl.so.2 / ncclAllReduce not found");
    }
    if (!film || !ncclComm) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_film_allreduce: null argument");
    if (int rc = pipeFlush(film, (cudaStream_t)stream)) return rc;
    int rc = fn(film->d, film->d, (size_t)film->w * film->h * 4, 7, 0, ncclComm, (cudaStream_t)stream);
    if (rc != 0) return fail(20000 + rc, "zl_film_allreduce: ncclAllReduce failed");
    return 0;
}

// ---- pass launches ----
static int checkPass(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, const char* who) {
    if (!s || !f || !p) return fail(ZL_ERR_INVALID_ARGUMENT, std::string(who) + ": null argument");
    if (p->filmW != f->w || p->filmH != f->h) return fail(ZL_ERR_INVALID_ARGUMENT, std::string(who) + ": params film size differs from the film");
    return 0;
}

static constexpr int kWfTraceBlock = 128;
// Ray sorting pays for itself (~1 ms per 4K pass) only when the MTBVH is large enough for incoherent lanes to miss
// the caches: +8 % on the Rungholt-class pass, +3 % Sponza-class, -29 % on the 36-triangle Cornell box, -14 % on the
// 6 k-triangle default scene (profiles/r1_trace_sweep.md).  ZL_WF_SORT=0/1 overrides.
static constexpr int kWfSortMinTriangles = 65536;

static int wfEnsure(ZlFilm* f, size_t needSlots = 0, bool second = false, int extra = 0) {
    WfWorkspace*& slot = extra == 2 ? f->wf4 : (extra == 1 ? f->wf3 : (second ? f->wf2 : f->wf));
    if (slot && slot->capacity >= needSlots) return 0;
    if (slot) { cudaDeviceSynchronize(); cudaFree(slot->block); delete slot; slot = nullptr; }
    auto* w = new WfWorkspace();
    WfState& st = w->st;
    st.tilesX = (f->w + 7) / 8; st.tilesY = (f->h + 3) / 4;
    st.nSlots = st.tilesX * st.tilesY * 32;
    const size_t n = ((std::max((size_t)st.nSlots, needSlots) + 31) / 32) * 32;
    w->capacity = n;
    const size_t vec = n * sizeof(float4), q = (n * sizeof(int) + 255) / 256 * 256;
    int sortBits = kWfSortBitsDefault;
    if (const char* e = std::getenv("ZL_WF_SORT_BITS")) sortBits = std::min(kWfSortBitsMax, std::max(4, std::atoi(e)));   // >= 4: each of the two histograms must be a whole number of 8192-bin scan tiles
    w->histInts = 2 * (size_t)wfSortBins(sortBits) + (size_t)wfScanBlocks(sortBits) + 64;     // two histograms, scan block bases, ticket
    w->bytes = 11 * vec + (kWfBins + 3 + 2 + 2 + 1) * q + kWfCounters * sizeof(int) + w->histInts * sizeof(int);
    cudaError_t e = cudaMalloc(&w->block, w->bytes);
    if (e != cudaSuccess) { delete w; return fail((int)e, std::string("wavefront workspace: ") + cudaGetErrorString(e)); }
    char* p = (char*)w->block;
    auto take = [&](size_t b) { char* r = p; p += b; return r; };
    {   // the eight common fields: one 128-byte record per slot (kWfRecordVecs = 8), or eight arrays (ZL_WF_AOS=0)
        float4* rec = (float4*)take(8 * vec);
        const size_t step = kWfRecordVecs == 8 ? 1 : n;
        // 32-byte sectors of a record: {hit0, hit1} {dir, sh} {thr, res} {smp, shc}: sort and trace touch only the first two
        st.hit[0].p = rec + 0 * step; st.hit[1].p = rec + 1 * step; st.dir.p = rec + 2 * step; st.sh.p = rec + 3 * step;
        st.thr.p = rec + 4 * step; st.res.p = rec + 5 * step; st.smp.p = (uint4*)(rec + 6 * step); st.shc.p = rec + 7 * step;
    }
    st.sho = (float4*)take(vec);
    st.aux = (float4*)take(vec); st.nrm = (float4*)take(vec); st.tdist = (float*)take(q);
    for (int t = 0; t < kWfBins; t++) st.qIn[t] = (int*)take(q);
    st.qS = (int*)take(q); st.qE = (int*)take(q); st.qT = (int*)take(q);
    st.qSs = (int*)take(q); st.qEs = (int*)take(q); st.keyTmp = (int*)take(2 * q);
    st.hist = (int*)take(w->histInts * sizeof(int));
    st.sortBits = sortBits; st.sortBins = wfSortBins(sortBits);
    st.cnt = (int*)take(kWfCounters * sizeof(int));
    st.sortMode = 0;
    st.capacity = (int)n;
    st.fusedKeys = 0;
    int dev = 0, sms = 148, perSm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    w->sms = sms;
    // persistent grids: every SM filled to the occupancy limit of the kernel
    auto fill = [&](auto kernel, int block) {
        perSm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, block, 0);
        return sms * (perSm > 0 ? perSm : 1);
    };
    w->gridShade[0] = fill(wfShadeKernel<0>, 128); w->gridShade[1] = fill(wfShadeKernel<1>, 128); w->gridShade[2] = fill(wfShadeKernel<2>, 128);
    w->gridShade[3] = fill(wfShadeKernel<3>, 128); w->gridShade[4] = fill(wfShadeKernel<4>, 128);
    w->gridTrace = fill(wfTraceKernel<kWfTraceBlock>, kWfTraceBlock);
    w->gridTraceSimple[2] = fill(wfTraceSimpleKernel<kWfTraceBlock, 12, 0>, kWfTraceBlock);
    // Two passes in flight (variant 2): 12 trace CTAs fill an SM's registers, so the other pass's generate / shade / resolve kernels only run in
    // the tails.  With 10 the trace stage itself is 4 % slower and the pass 2 % faster (7.31 -> 7.17 ms, profiles/r2_trace_sweep.md): one stage
    // CTA of the other pass fits next to them.  ZL_WF_TRACE_CTAS_PER_SM overrides both.
    w->gridTracePipelined = std::min(w->gridTraceSimple[2], w->sms * 10);
    if (const char* e = std::getenv("ZL_WF_TRACE_CTAS_PER_SM")) {
        const int c = std::atoi(e);
        if (c >= 1 && c <= 12) w->gridTraceSimple[2] = w->gridTracePipelined = std::min(w->gridTraceSimple[2], w->sms * c);
    }
    w->gridResolve = fill(wfResolveKernel, 128);
    w->gridResolveDense = fill(wfResolveDenseKernel, 128);
    w->gridShadeDense = fill(wfShadeDenseKernel, 128);
    w->gridLightShade[0] = fill(wfLightShadeKernel<0>, 128); w->gridLightShade[1] = fill(wfLightShadeKernel<1>, 128); w->gridLightShade[2] = fill(wfLightShadeKernel<2>, 128);
    w->gridLightShade[3] = fill(wfLightShadeKernel<3>, 128); w->gridLightShade[4] = fill(wfLightShadeKernel<4>, 128);
    w->gridTripleShade[0] = fill(wfTripleShadeKernel<0>, 128); w->gridTripleShade[1] = fill(wfTripleShadeKernel<1>, 128); w->gridTripleShade[2] = fill(wfTripleShadeKernel<2>, 128);
    w->gridTripleShade[3] = fill(wfTripleShadeKernel<3>, 128); w->gridTripleShade[4] = fill(wfTripleShadeKernel<4>, 128);
    w->gridTripleLightShade[0] = fill(wfTripleLightShadeKernel<0>, 128); w->gridTripleLightShade[1] = fill(wfTripleLightShadeKernel<1>, 128);
    w->gridTripleLightShade[2] = fill(wfTripleLightShadeKernel<2>, 128); w->gridTripleLightShade[3] = fill(wfTripleLightShadeKernel<3>, 128);
    w->gridTripleLightShade[4] = fill(wfTripleLightShadeKernel<4>, 128);
    w->gridTripleResolve = fill(wfTripleResolveKernel, 128);
    slot = w;
    return 0;
}

}  // extern "C" (templates below need C++ linkage)

struct WfOptions {
    int simpleMask = 3;      // A/B switch: bit 0 = plain-loop kernel for the first rays (b = 0), bit 1 = for every other bounce (0 = regenerating kernel; camera paths only)
    int sortMode = 0;
    int minBlocksSet = 0;    // ZL_WF_TRACE_MINB given (each trace kernel has its own default otherwise)
    int minBlocks = 12;      // 40 registers, 48 warps per SM: best of 8/10/12/14/16 (profiles/r1_trace_sweep.md)
    int sortRays = -1;       // -1 = by scene size (kWfSortMinTriangles), 0 = never, 1 = always
    int loop = -1;           // -1 = the scene's choice (ZlScene::traceLoop: 5 when the whole MTBVH table is L1-resident, else 0); 0 = wfTraceSimpleKernel; 1 = look-ahead node loads; 2 = deferred leaf tests; 3 = both (wfTraceDeferKernel)
    int flushAt = 12;        // deferred leaf tests: run them once this many lanes hold one
    int fuseSortKeys = 1;    // path tracer: sort keys + histogram recorded by the shade kernels (A/B: 0 = separate wfSortCountKernel)
    int roundSteps = 16;     // loop 4 (wfTraceRefillKernel): steps per lane between two warp-wide retire / refill points
    int refillAt = 8;        // loop 4: hand out new rays once this many lanes are idle
    int refillFrom = 1;      // loop 4: first bounce traced by the refill kernel (camera rays are coherent: plain loop)
    int overlap = 1;         // path tracer: resolve(b) and the minor-type shade kernels on side streams (A/B: 0 = one stream)
    int octantWalk = -1;     // -1 = the scene's choice (DScene::octantWalk, by size); 1 = octant-specialised packed walks; 0 = general packed walk; 2 = scalar walk
    int bvh2Walk = 0;        // 1: pure rays walk the child-boxes-in-the-parent records with a short stack (traverseBvh2); ZL_BVH2_WALK=0: threaded records
    int nodePolicy = 0;      // node-record loads with evict_last in L1 / L2 (ZL_NODE_POLICY=1)
    int statePolicy = 0;     // trace kernel's path-state accesses through the streaming operators (ZL_STATE_POLICY=1)
    int tracePipe = 0;       // ZL_WF_TRACE_PIPE=1: chunk heads (claim / queue entry / path state) software-pipelined under the previous walks
    WfOptions() {
        bvh2Walk = bvh2WalkEnabled() ? 1 : 0;
        if (const char* e = std::getenv("ZL_NODE_POLICY")) nodePolicy = std::atoi(e) != 0 ? 1 : 0;
        if (const char* e = std::getenv("ZL_STATE_POLICY")) statePolicy = std::atoi(e) != 0 ? 1 : 0;
        if (const char* e = std::getenv("ZL_WF_TRACE_PIPE")) tracePipe = std::atoi(e) != 0 ? 1 : 0;
        if (const char* e = std::getenv("ZL_OCTANT_WALK")) octantWalk = std::min(2, std::max(-1, std::atoi(e)));      // -1 (default): by scene size
        if (const char* e = std::getenv("ZL_WF_ROUND_STEPS")) roundSteps = std::max(1, std::atoi(e));
        if (const char* e = std::getenv("ZL_WF_REFILL_AT")) refillAt = std::min(32, std::max(1, std::atoi(e)));
        if (const char* e = std::getenv("ZL_WF_REFILL_FROM")) refillFrom = std::atoi(e);
        if (const char* e = std::getenv("ZL_WF_OVERLAP")) overlap = std::atoi(e);
        if (const char* e = std::getenv("ZL_WF_TRACE_LOOP")) loop = std::atoi(e);
        if (const char* e = std::getenv("ZL_WF_FLUSH_AT")) flushAt = std::atoi(e);
        if (const char* e = std::getenv("ZL_WF_FUSE_SORT_KEYS")) fuseSortKeys = std::atoi(e);
        if (const char* e = std::getenv("ZL_WF_TRACE_SIMPLE")) simpleMask = std::atoi(e);
        if (const char* e = std::getenv("ZL_WF_SORT_MODE")) sortMode = std::atoi(e);
        if (const char* e = std::getenv("ZL_WF_TRACE_MINB")) { minBlocks = std::atoi(e); minBlocksSet = 1; }
        if (const char* e = std::getenv("ZL_WF_SORT")) sortRays = std::atoi(e) != 0 ? 1 : 0;
    }
};

// persistent grid of a trace kernel instantiation: SMs x resident CTAs (cached per instantiation by the caller)
template <typename K>
static int wfGridOf(K kernel, int sms) {
    int perSm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kWfTraceBlock, 0);
    return sms * (perSm > 0 ? perSm : 1);
}
template <int MINB, int MODE, int LOOP>
static void wfLaunchDefer(const ZlScene* s, const ZlFilm* f, const WfState& wt, int b, int last, float shadowEps, int flushAt, cudaStream_t stream) {
    static int grid = 0;
    if (grid == 0) grid = wfGridOf(wfTraceDeferKernel<kWfTraceBlock, MINB, MODE, LOOP>, f->wf->sms);
    wfTraceDeferKernel<kWfTraceBlock, MINB, MODE, LOOP><<<grid, kWfTraceBlock, 0, stream>>>(s->d, wt, b, last, shadowEps, f->d, f->w, f->h, flushAt);
}
template <int MINB, int MODE>
static void wfLaunchRefill(const ZlScene* s, const ZlFilm* f, const WfState& wt, int b, int last, float shadowEps, int roundSteps, int refillAt, cudaStream_t stream) {
    static int grid = 0;
    if (grid == 0) grid = wfGridOf(wfTraceRefillKernel<kWfTraceBlock, MINB, MODE>, f->wf->sms);
    wfTraceRefillKernel<kWfTraceBlock, MINB, MODE><<<grid, kWfTraceBlock, 0, stream>>>(s->d, wt, b, last, shadowEps, f->d, f->w, f->h, roundSteps, refillAt);
    g_launches++;
    wfTraceSimpleKernel<kWfTraceBlock, 12, MODE, true><<<f->wf->sms, kWfTraceBlock, 0, stream>>>(s->d, wt, b, last, shadowEps, f->d, f->w, f->h);   // the non-"pure" rays it listed
}
template <int MINB, int MODE>
static void wfLaunchDual(const ZlScene* s, const ZlFilm* f, const WfState& wt, const DScene& dS, int b, int last, float shadowEps, cudaStream_t stream) {
    static int grid = 0;
    if (grid == 0) grid = wfGridOf(wfTraceDualKernel<kWfTraceBlock, MINB, MODE>, f->wf->sms);
    wfTraceDualKernel<kWfTraceBlock, MINB, MODE><<<grid, kWfTraceBlock, 0, stream>>>(dS, wt, b, last, shadowEps, f->d, f->w, f->h);
}
template <int MODE>
static void wfLaunchDualMinb(const ZlScene* s, const ZlFilm* f, const WfState& wt, const DScene& dS, int minb, int b, int last, float shadowEps, cudaStream_t stream) {
    (void)minb;      // 8 / 10 / 12 blocks per SM were measured and dropped (profiles/r1_trace_sweep.md): 9 = 56 registers, 36 warps
    wfLaunchDual<9, MODE>(s, f, wt, dS, b, last, shadowEps, stream);
}
template <int MODE>
static void wfLaunchRefillMinb(const ZlScene* s, const ZlFilm* f, const WfState& wt, int minb, int b, int last, float shadowEps, int roundSteps, int refillAt, cudaStream_t stream) {
    switch (minb) {
    case 8: wfLaunchRefill<8, MODE>(s, f, wt, b, last, shadowEps, roundSteps, refillAt, stream); break;
    case 9: wfLaunchRefill<9, MODE>(s, f, wt, b, last, shadowEps, roundSteps, refillAt, stream); break;
    case 10: wfLaunchRefill<10, MODE>(s, f, wt, b, last, shadowEps, roundSteps, refillAt, stream); break;
    default: wfLaunchRefill<12, MODE>(s, f, wt, b, last, shadowEps, roundSteps, refillAt, stream); break;
    }
}
template <int MODE, int LOOP>
static void wfLaunchDeferMinb(const ZlScene* s, const ZlFilm* f, const WfState& wt, int minb, int b, int last, float shadowEps, int flushAt, cudaStream_t stream) {
    switch (minb) {
    case 8: wfLaunchDefer<8, MODE, LOOP>(s, f, wt, b, last, shadowEps, flushAt, stream); break;
    case 10: wfLaunchDefer<10, MODE, LOOP>(s, f, wt, b, last, shadowEps, flushAt, stream); break;
    default: wfLaunchDefer<12, MODE, LOOP>(s, f, wt, b, last, shadowEps, flushAt, stream); break;
    }
}

// sort (optional) + trace of the S and E queues of bounce b.  MODE 0: camera paths, MODE 1: light paths (splats).
static bool wfSortEnabled(const ZlScene* s, const WfOptions& o) {
    return o.sortRays < 0 ? s->d.numTriangles >= kWfSortMinTriangles : o.sortRays != 0;
}
static int wfSortClearHistogram(const WfWorkspace& w, cudaStream_t stream) {
    ZL_CK(cudaMemsetAsync(w.st.hist, 0, w.histInts * sizeof(int), stream));
    return 0;
}
// keysReady: the shade kernels of this bounce already recorded keys + histogram (WfState::fusedKeys)
template <int MODE>
static int wfTraceStage(ZlScene* s, ZlFilm* f, const WfOptions& o, int b, int last, bool sortThis, float shadowEps, cudaStream_t stream, bool keysReady = false,
                        const WfWorkspace* ws_ = nullptr) {
    const WfWorkspace& w = ws_ ? *ws_ : *f->wf;
    WfState wt = w.st;
    wt.sortMode = o.sortMode;
    if (wfSortEnabled(s, o) && sortThis) {
        StageScope scope(ZL_STAGE_SORT, stream);
        if (!keysReady) {
            if (int rc = wfSortClearHistogram(w, stream)) return rc;
            wfSortCountKernel<<<w.sms * 8, 256, 0, stream>>>(s->d, wt, b, MODE == 1 ? 1 : 0);
            ZL_LAUNCHED();
        }
        wfSortScanKernel<<<wfScanBlocks(w.st.sortBits), 1024, 0, stream>>>(w.st);
        ZL_LAUNCHED();
        wfSortScatterKernel<<<w.sms * 8, 256, 0, stream>>>(w.st, b);
        ZL_LAUNCHED();
        wt.qS = w.st.qSs; wt.qE = w.st.qEs;
    }
    StageScope scope(ZL_STAGE_TRACE, stream);
    const int loop = o.loop >= 0 ? o.loop : s->traceLoop;
    if (loop == 5) {
        DScene dS = s->d;
        if (o.octantWalk >= 0) dS.octantWalk = o.octantWalk;
        wfLaunchDualMinb<MODE>(s, f, wt, dS, o.minBlocksSet ? o.minBlocks : 9, b, last, shadowEps, stream);
    } else if (loop == 4 && b >= o.refillFrom) {
        wfLaunchRefillMinb<MODE>(s, f, wt, o.minBlocks, b, last, shadowEps, o.roundSteps, o.refillAt, stream);
    } else if (loop >= 1 && loop <= 3) {
        if (loop == 1) wfLaunchDeferMinb<MODE, 1>(s, f, wt, o.minBlocks, b, last, shadowEps, o.flushAt, stream);
        else if (loop == 2) wfLaunchDeferMinb<MODE, 2>(s, f, wt, o.minBlocks, b, last, shadowEps, o.flushAt, stream);
        else wfLaunchDeferMinb<MODE, 3>(s, f, wt, o.minBlocks, b, last, shadowEps, o.flushAt, stream);
    } else if (loop == 7) {      // intra-warp ray compaction (traversePureCompact)
        static int grid7[2] = {0, 0};
        const bool lean = o.minBlocksSet && o.minBlocks >= 10;      // ZL_WF_TRACE_MINB=10: 48 registers (spills); default 8 blocks = 64 registers
        auto kern = lean ? wfTraceCompactKernel<kWfTraceBlock, 10, MODE> : wfTraceCompactKernel<kWfTraceBlock, 8, MODE>;
        if (!grid7[lean]) grid7[lean] = wfGridOf(kern, w.sms);
        DScene dS = s->d;
        if (o.octantWalk >= 0) dS.octantWalk = o.octantWalk;
        dS.nodePolicy = o.nodePolicy; dS.statePolicy = o.statePolicy;
        kern<<<grid7[lean], kWfTraceBlock, 0, stream>>>(dS, wt, b, last, shadowEps, f->d, f->w, f->h);
    } else if (loop == 6) {      // shared-memory staged top levels (TMA bulk copy per CTA), 256 threads x 6 CTAs = 48 warps per SM
        static int grid6 = 0;
        auto kern = wfTraceSimpleKernel<256, 6, MODE, false, true>;
        if (!grid6) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTopBytes);
            int perSm = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, 256, kTopBytes);
            grid6 = w.sms * std::max(perSm, 1);
        }
        DScene dS = s->d;
        dS.nodePolicy = o.nodePolicy; dS.statePolicy = o.statePolicy;
        kern<<<grid6, 256, kTopBytes, stream>>>(dS, wt, b, last, shadowEps, f->d, f->w, f->h);
    } else if (MODE == 1 || (o.simpleMask & (b == 0 ? 1 : 2))) {
        DScene dS = s->d;
        if (o.octantWalk >= 0) dS.octantWalk = o.octantWalk;
        dS.nodePolicy = o.nodePolicy; dS.statePolicy = o.statePolicy;
        dS.bvh2 = o.bvh2Walk ? s->bvh2 : nullptr;
        static int carveout = -2;    // A/B switch ZL_WF_L1_CARVEOUT: preferred shared-memory carve-out (percent) of the default trace kernel; unset = driver default
        if (carveout == -2) {
            const char* e = std::getenv("ZL_WF_L1_CARVEOUT");
            carveout = e ? std::atoi(e) : -1;
            if (carveout >= 0) {
                cudaFuncSetAttribute(wfTraceSimpleKernel<kWfTraceBlock, 12, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
                cudaFuncSetAttribute(wfTraceSimpleKernel<kWfTraceBlock, 12, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
            }
        }
        // (8 / 10 / 14 / 16 blocks per SM were measured and dropped, profiles/r1_trace_sweep.md: 12 = 40 registers, 48 warps)
        const char* eMinb = std::getenv("ZL_BVH2_MINB");
        const int bvh2Minb = eMinb ? std::atoi(eMinb) : 0;
        if (dS.bvh2 && bvh2Minb == 8) {
            static int g8 = 0;
            if (!g8) g8 = wfGridOf(wfTraceSimpleKernel<kWfTraceBlock, 8, MODE>, w.sms);
            wfTraceSimpleKernel<kWfTraceBlock, 8, MODE><<<g8, kWfTraceBlock, 0, stream>>>(dS, wt, b, last, shadowEps, f->d, f->w, f->h);
        } else if (dS.bvh2 && bvh2Minb == 10) {
            static int g10 = 0;
            if (!g10) g10 = wfGridOf(wfTraceSimpleKernel<kWfTraceBlock, 10, MODE>, w.sms);
            wfTraceSimpleKernel<kWfTraceBlock, 10, MODE><<<g10, kWfTraceBlock, 0, stream>>>(dS, wt, b, last, shadowEps, f->d, f->w, f->h);
        } else
        if ((!dS.bvh2 && !dS.nodePolicy && !dS.statePolicy) && o.minBlocksSet && (o.minBlocks == 10 || o.minBlocks == 11) && dS.octantWalk >= 1) {      // A/B: ZL_WF_TRACE_MINB=10|11 (48 / 46 registers)
            static int gridAlt[2][2] = {{0, 0}, {0, 0}};
            const int mi = o.minBlocks - 10, wi = dS.octantWalk - 1;
            if (o.minBlocks == 10) {
                if (wi == 0) { auto k = wfTraceSimpleKernel<kWfTraceBlock, 10, MODE, false, false, 1>; if (!gridAlt[mi][wi]) gridAlt[mi][wi] = wfGridOf(k, w.sms); k<<<gridAlt[mi][wi], kWfTraceBlock, 0, stream>>>(dS, wt, b, last, shadowEps, f->d, f->w, f->h); }
                else { auto k = wfTraceSimpleKernel<kWfTraceBlock, 10, MODE, false, false, 2>; if (!gridAlt[mi][wi]) gridAlt[mi][wi] = wfGridOf(k, w.sms); k<<<gridAlt[mi][wi], kWfTraceBlock, 0, stream>>>(dS, wt, b, last, shadowEps, f->d, f->w, f->h); }
            } else {
                if (wi == 0) { auto k = wfTraceSimpleKernel<kWfTraceBlock, 11, MODE, false, false, 1>; if (!gridAlt[mi][wi]) gridAlt[mi][wi] = wfGridOf(k, w.sms); k<<<gridAlt[mi][wi], kWfTraceBlock, 0, stream>>>(dS, wt, b, last, shadowEps, f->d, f->w, f->h); }
                else { auto k = wfTraceSimpleKernel<kWfTraceBlock, 11, MODE, false, false, 2>; if (!gridAlt[mi][wi]) gridAlt[mi][wi] = wfGridOf(k, w.sms); k<<<gridAlt[mi][wi], kWfTraceBlock, 0, stream>>>(dS, wt, b, last, shadowEps, f->d, f->w, f->h); }
            }
        } else
        if ((!dS.bvh2 && !dS.nodePolicy && !dS.statePolicy) && o.tracePipe && dS.octantWalk >= 1) {      // A/B: ZL_WF_TRACE_PIPE=1, software-pipelined chunk heads
            if (dS.octantWalk == 1) wfTraceSimpleKernel<kWfTraceBlock, 12, MODE, false, false, 1, true><<<w.gridTraceSimple[2], kWfTraceBlock, 0, stream>>>(dS, wt, b, last, shadowEps, f->d, f->w, f->h);
            else wfTraceSimpleKernel<kWfTraceBlock, 12, MODE, false, false, 2, true><<<w.gridTraceSimple[2], kWfTraceBlock, 0, stream>>>(dS, wt, b, last, shadowEps, f->d, f->w, f->h);
        } else
        if (!dS.bvh2 && !dS.nodePolicy && !dS.statePolicy && dS.octantWalk == 1)      // the default configurations: instantiations without the switched-off A/B walks
            wfTraceSimpleKernel<kWfTraceBlock, 12, MODE, false, false, 1><<<(ws_ && MODE == 0) ? w.gridTracePipelined : w.gridTraceSimple[2], kWfTraceBlock, 0, stream>>>(dS, wt, b, last, shadowEps, f->d, f->w, f->h);
        else if (!dS.bvh2 && !dS.nodePolicy && !dS.statePolicy && dS.octantWalk == 2)
            wfTraceSimpleKernel<kWfTraceBlock, 12, MODE, false, false, 2><<<(ws_ && MODE == 0) ? w.gridTracePipelined : w.gridTraceSimple[2], kWfTraceBlock, 0, stream>>>(dS, wt, b, last, shadowEps, f->d, f->w, f->h);
        else
        wfTraceSimpleKernel<kWfTraceBlock, 12, MODE><<<w.gridTraceSimple[2], kWfTraceBlock, 0, stream>>>(dS, wt, b, last, shadowEps, f->d, f->w, f->h);
    } else wfTraceKernel<kWfTraceBlock><<<w.gridTrace, kWfTraceBlock, 0, stream>>>(s->d, wt, b, last);
    ZL_LAUNCHED();
    return 0;
}

// ---- side streams of the overlapped pass schedule ----
static int wfEnsureSide(WfWorkspace& w) {
    if (w.side[0]) return 0;
    for (auto& st_ : w.side) ZL_CK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    ZL_CK(cudaEventCreateWithFlags(&w.evFork, cudaEventDisableTiming));
    ZL_CK(cudaEventCreateWithFlags(&w.evTraced, cudaEventDisableTiming));
    for (auto& e : w.evJoin) ZL_CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : w.evResolved) ZL_CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return 0;
}
// The shade kernels of one bounce (one per material type, disjoint paths, atomic queue appends): the first goes to the main
// stream, the others alternate over side streams 1 and 2 behind a fork event; join() makes the main stream wait for them.
struct WfFan {
    WfWorkspace& w; cudaStream_t main; bool overlap; int n = 0; bool used[3] = {false, false, false};
    WfFan(WfWorkspace& w_, bool overlap_, cudaStream_t main_) : w(w_), main(main_), overlap(overlap_) {
        if (overlap) cudaEventRecord(w.evFork, main);
    }
    cudaStream_t next() {
        const int l = n++;
        if (!overlap || l == 0) return main;
        const int k = 1 + (l - 1) % 2;
        if (!used[k]) { cudaStreamWaitEvent(w.side[k], w.evFork, 0); used[k] = true; }
        return w.side[k];
    }
    int join() {
        for (int k = 1; k < 3; k++)
            if (used[k]) { ZL_CK(cudaEventRecord(w.evJoin[k], w.side[k])); ZL_CK(cudaStreamWaitEvent(main, w.evJoin[k], 0)); used[k] = false; }
        return 0;
    }
};
// resolve(b) on side stream 0 behind trace(b); nothing later in the pass touches what it reads (the ended-paths queue is one
// array for the whole pass, wfEndedBase), so the main stream waits for it only at the end of the pass
struct WfResolveSide {
    static int before(WfWorkspace& w, cudaStream_t main) {              // returns through ZL_CK
        ZL_CK(cudaEventRecord(w.evTraced, main));
        ZL_CK(cudaStreamWaitEvent(w.side[0], w.evTraced, 0));
        return 0;
    }
    static int after(WfWorkspace& w, int b) {
        ZL_CK(cudaEventRecord(w.evResolved[b & 1], w.side[0]));
        w.resolvePending[b & 1] = true;
        return 0;
    }
    static int waitParity(WfWorkspace& w, int parity, cudaStream_t main) {
        if (w.resolvePending[parity]) { ZL_CK(cudaStreamWaitEvent(main, w.evResolved[parity], 0)); w.resolvePending[parity] = false; }
        return 0;
    }
};

// ---- a wavefront pass as a replayed CUDA graph (kernelVariant 3) ----
// A pass of the sequential wavefront schedule is 30-60 launches plus the fork / join events of its side streams; on small films (720p:
// 0.6 ms of GPU work per pass) the host cannot issue them as fast as the GPU retires them.  The first pass with a given configuration
// runs with plain launches (allocations, one-time attribute calls), the second is stream-captured into a graph, every later one is ONE
// cudaGraphLaunch.  Kernel arguments are baked into a graph, and the only values that change from pass to pass are uSpp and
// uFreeCounter: the graph's first node is setPassCountersKernel(counters, spp, freeCounter), whose arguments are rewritten before each
// replay (cudaGraphExecKernelNodeSetParams), and the kernels that need the pair read it from there (WfState::passCounters,
// wfPassParams).  Same kernels, same order, same arguments otherwise: films are bit-identical to variant 1.
static thread_local const int* g_wfPassCounters = nullptr;
static std::string wfOptionsKey() {
    std::string k;
    for (const char* n : {"ZL_OCTANT_WALK", "ZL_WF_ROUND_STEPS", "ZL_WF_REFILL_AT", "ZL_WF_REFILL_FROM", "ZL_WF_OVERLAP", "ZL_WF_TRACE_LOOP", "ZL_WF_FLUSH_AT",
                          "ZL_WF_FUSE_SORT_KEYS", "ZL_WF_TRACE_SIMPLE", "ZL_WF_SORT_MODE", "ZL_WF_TRACE_MINB", "ZL_WF_SORT", "ZL_NODE_POLICY", "ZL_STATE_POLICY",
                          "ZL_BVH2_WALK", "ZL_BVH2_MINB", "ZL_WF_L1_CARVEOUT", "ZL_WF_TRACE_PIPE"}) {
        const char* e = std::getenv(n);
        k += e ? e : "-";
        k += ';';
    }
    return k;
}
template <class Launch>
static int graphPass(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, cudaStream_t user, int which, Launch launch) {
    if (g_stageTimer.enabled) return launch(user);                 // stage timing brackets every launch group with events: plain launches
    ZlFilm::PassGraph& g = f->graphs[which];
    ZlRenderParams key = *p;
    key.spp = 0; key.freeCounter = 0;
    const std::string ok = wfOptionsKey();
    if (g.scene != s || g.binMask != s->binMask || std::memcmp(&g.key, &key, sizeof key) != 0 || g.optKey != ok) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); cudaGraphDestroy(g.graph); }
        g = ZlFilm::PassGraph();
        g.scene = s; g.binMask = s->binMask; g.key = key; g.optKey = ok;
    }
    if (!g.warm || g.failed) { g.warm = true; return launch(user); }      // (a failed capture was reported once; the same plain launches after it)
    if (!f->graphStream) {
        ZL_CK(cudaStreamCreateWithFlags(&f->graphStream, cudaStreamNonBlocking));
        ZL_CK(cudaEventCreateWithFlags(&f->evGraphIn, cudaEventDisableTiming));
        ZL_CK(cudaEventCreateWithFlags(&f->evGraphOut, cudaEventDisableTiming));
        ZL_CK(cudaMalloc((void**)&f->dPassCounters, 2 * sizeof(int)));
    }
    const cudaStream_t G = f->graphStream;       // the caller's stream may be the legacy default stream, which cannot be captured
    ZL_CK(cudaEventRecord(f->evGraphIn, user));
    ZL_CK(cudaStreamWaitEvent(G, f->evGraphIn, 0));
    if (!g.exec) {
        const unsigned long long l0 = g_launches.load();
        ZL_CK(cudaStreamBeginCapture(G, cudaStreamCaptureModeThreadLocal));
        setPassCountersKernel<<<1, 1, 0, G>>>(f->dPassCounters, p->spp, p->freeCounter);
        g_launches++;
        g_wfPassCounters = f->dPassCounters;
        const int rc = launch(G);
        g_wfPassCounters = nullptr;
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(G, &graph);
        g.launches = g_launches.load() - l0;
        g_launches -= g.launches;                                    // nothing ran yet: replays are counted when they are launched
        if (rc != 0 || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            g.failed = true;
            return rc != 0 ? rc : fail((int)ce, std::string("graphPass: stream capture failed: ") + cudaGetErrorString(ce));
        }
        size_t nNodes = 0;
        ZL_CK(cudaGraphGetNodes(graph, nullptr, &nNodes));
        std::vector<cudaGraphNode_t> nodes(nNodes);
        ZL_CK(cudaGraphGetNodes(graph, nodes.data(), &nNodes));
        for (cudaGraphNode_t nd : nodes) {
            cudaGraphNodeType ty;
            if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
            cudaKernelNodeParams kp{};
            if (cudaGraphKernelNodeGetParams(nd, &kp) == cudaSuccess && kp.func == (void*)setPassCountersKernel) { g.setNode = nd; break; }
        }
        cudaError_t ie = g.setNode ? cudaGraphInstantiate(&g.exec, graph, 0) : cudaErrorUnknown;
        if (ie != cudaSuccess) { cudaGraphDestroy(graph); g.exec = nullptr; g.failed = true; return fail((int)ie, "graphPass: cudaGraphInstantiate failed"); }
        g.graph = graph;      // setNode is a handle into the source graph: it lives as long as the executable graph does
    }
    int* counters = f->dPassCounters;
    int spp = p->spp, fc = p->freeCounter;
    void* args[3] = {&counters, &spp, &fc};
    cudaKernelNodeParams np{};
    np.func = (void*)setPassCountersKernel; np.gridDim = dim3(1); np.blockDim = dim3(1); np.sharedMemBytes = 0; np.kernelParams = args; np.extra = nullptr;
    ZL_CK(cudaGraphExecKernelNodeSetParams(g.exec, g.setNode, &np));
    ZL_CK(cudaGraphLaunch(g.exec, G));
    g_launches += g.launches;
    ZL_CK(cudaEventRecord(f->evGraphOut, G));
    ZL_CK(cudaStreamWaitEvent(user, f->evGraphOut, 0));
    return 0;
}

// One pass of the wavefront path tracer.  Dependencies between the stages of bounce b:
//     trace(b-1) -> shade<type>(b) [independent of each other: disjoint paths, atomic queue appends] -> sort(b) -> trace(b)
//     trace(b)   -> resolve(b)     [paths that ended: disjoint from everything later in the pass; film pixels are owned by one path]
// With `overlap` the main stream carries shade<first type> / sort / trace, the other shade kernels and resolve(b) run on side
// streams (fork / join with events), so the small latency-bound kernels and the tails of the persistent grids fill each
// other's idle SMs.  Per-path arithmetic and the one film write per pixel are unchanged: the film is bit-identical.
static int launchWavefrontPathPass(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, cudaStream_t stream) {
    if (int rc = wfEnsure(f)) return rc;
    WfWorkspace& w = *f->wf;
    const WfOptions o;
    const bool fused = wfSortEnabled(s, o) && o.fuseSortKeys;      // shade kernels record the sort keys of the rays they queue
    const bool overlap = o.overlap != 0 && !g_stageTimer.enabled;  // stage timing needs the stages back to back on one stream
    if (overlap) { if (int rc = wfEnsureSide(w)) return rc; }
    WfState ws = w.st;
    ws.sortMode = o.sortMode;
    ws.fusedKeys = fused ? 1 : 0;
    ws.passCounters = g_wfPassCounters;      // non-null while this pass is being captured into a CUDA graph (graphPass)
    ZL_CK(cudaMemsetAsync(w.st.cnt, 0, kWfCounters * sizeof(int), stream));
    {   StageScope scope(ZL_STAGE_GENERATE, stream);
        wfGenerateKernel<<<(w.st.nSlots + 127) / 128, 128, 0, stream>>>(s->d, *p, ws);
        ZL_LAUNCHED();
    }
    for (int b = 0; b <= p->maxDepth; b++) {
        if (b > 0) {    // one shade kernel per material-type bin present in the scene
            StageScope scope(ZL_STAGE_SHADE, stream);
            if (fused) { if (int rc = wfSortClearHistogram(w, stream)) return rc; }
            WfFan fan(w, overlap, stream);
            if (s->binMask & 1u) { wfShadeKernel<0><<<w.gridShade[0], 128, 0, fan.next()>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 2u) { wfShadeKernel<1><<<w.gridShade[1], 128, 0, fan.next()>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 4u) { wfShadeKernel<2><<<w.gridShade[2], 128, 0, fan.next()>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 8u) { wfShadeKernel<3><<<w.gridShade[3], 128, 0, fan.next()>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 16u) { wfShadeKernel<4><<<w.gridShade[4], 128, 0, fan.next()>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (int rc = fan.join()) return rc;
        }
        // camera rays are generated in tile order: already coherent, not sorted
        if (int rc = wfTraceStage<0>(s, f, o, b, b == p->maxDepth ? 1 : 0, b > 0, 1e-4f, stream, fused)) return rc;
        StageScope scope(ZL_STAGE_RESOLVE, stream);
        if (overlap) {
            if (int rc = WfResolveSide::before(w, stream)) return rc;
            wfResolveKernel<<<w.gridResolve, 128, 0, w.side[0]>>>(s->d, *p, ws, f->d, b);
            ZL_LAUNCHED();
            if (int rc = WfResolveSide::after(w, b)) return rc;
        } else {
            wfResolveKernel<<<w.gridResolve, 128, 0, stream>>>(s->d, *p, ws, f->d, b);
            ZL_LAUNCHED();
        }
    }
    if (overlap)        // join: whatever follows on `stream` (next pass, film read-back) sees every film write of this pass
        for (int k = 0; k < 2; k++) { if (int rc = WfResolveSide::waitParity(w, k, stream)) return rc; }
    return 0;
}

// ---- pipelined passes (variant 2) ----
// Two passes in flight.  Pass k runs on workspace / chain stream k & 1 (generate, shade, sort, trace), so the short, tail-bound
// late bounces of pass k overlap the wide first bounces of pass k + 1: measured upper bound with two independent integrators
// +15 % on the Rungholt-class 4K pass (tools/probe_pass_pipelining.py).  EVERY film write of a pass is in its resolve kernels,
// and all resolve kernels of all passes — and every read of the film (zl_film_download*, zl_film_postprocess) — go to ONE
// stream, the film stream R, in pass order: each pixel receives its adds in the order of the sequential schedule, so the film
// is bit-identical to variant 1 / the megakernel, and a frame read between two passes is a consistent snapshot.
//   chain c:  [wait: the previous pass on this workspace is resolved]  memset, generate, { shade(b), sort(b), trace(b), record evTraced }...
//   R      :  { wait evTraced(b), resolve(b) }...  record evPassResolved(c), evTail
// `stream` (the caller's) is involved only at the edges: the first pass after a flush waits for what the caller enqueued before
// (evUser), and zl_film_flush / the film calls make `stream` wait for evTail.
// Passes in flight of the pipelined path tracer.  The stage kernels of one pass fill what the trace kernels of the others leave idle (tails,
// and the SM resources the 10-CTA trace grid leaves free); measured (profiles/r2_trace_sweep.md): 2 -> 3 passes +2 % on the 4K Rungholt-class
// pass and the 1080p Sponza-class pass, +21 % on the 720p default scene, whose kernels are too short to fill the GPU; a fourth pass another
// +11 % at 720p and -0.3 % on the larger films.  Default: 3, and 4 for films below 2^20 pixels.
// ZL_WF_PIPE_DEPTH=2|3|4 overrides.  One 1.7 GB workspace per pass in flight at 4K.
static int pipeDepth(const ZlFilm* f) {
    if (const char* e = std::getenv("ZL_WF_PIPE_DEPTH")) { const int d = std::atoi(e); if (d >= 2 && d <= 4) return d; }
    return ((size_t)f->w * f->h < ((size_t)1 << 20)) ? 4 : 3;
}
// triple tracer: three pass pairs in flight (Sponza-class 1080p 294.7 -> 300.6 Msamples/s against two); ZL_WF_TRIPLE_DEPTH=2 is the A/B switch
static int tripleDepth(const ZlFilm*) {
    const char* e = std::getenv("ZL_WF_TRIPLE_DEPTH");
    return (e && std::atoi(e) == 2) ? 2 : 3;
}
static int pipeEnsure(ZlFilm* f, int depth = 2) {
    if (int rc = wfEnsure(f)) return rc;
    if (int rc = wfEnsure(f, 0, true)) return rc;
    if (depth >= 3) { if (int rc = wfEnsure(f, 0, false, 1)) return rc; }
    if (depth >= 4) { if (int rc = wfEnsure(f, 0, false, 2)) return rc; }
    if (!f->filmStream) {
        ZL_CK(cudaStreamCreateWithFlags(&f->filmStream, cudaStreamNonBlocking));
        ZL_CK(cudaEventCreateWithFlags(&f->evUser, cudaEventDisableTiming));
        ZL_CK(cudaEventCreateWithFlags(&f->evTail, cudaEventDisableTiming));
    }
    for (WfWorkspace* w : {f->wf, f->wf2, f->wf3, f->wf4}) {
        if (!w) continue;
        if (!w->chain) {
            ZL_CK(cudaStreamCreateWithFlags(&w->chain, cudaStreamNonBlocking));
            ZL_CK(cudaEventCreateWithFlags(&w->evPassResolved, cudaEventDisableTiming));
        }
        if (!w->evTraced) ZL_CK(cudaEventCreateWithFlags(&w->evTraced, cudaEventDisableTiming));
    }
    return 0;
}
static int launchWavefrontPathPassPipelined(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, cudaStream_t stream) {
    const int depth = pipeDepth(f);
    if (int rc = pipeEnsure(f, depth)) return rc;
    WfWorkspace* const ring[4] = {f->wf, f->wf2, f->wf3, f->wf4};
    WfWorkspace& w = *ring[f->pathRing++ % (unsigned)depth];      // (its own counter: the ring survives a change of depth or of integrator on this film)
    const cudaStream_t M = w.chain, R = f->filmStream;
    const WfOptions o;
    const bool fused = wfSortEnabled(s, o) && o.fuseSortKeys;
    if (!f->pipeDirty) {            // first pass since the film was last used from the caller's stream: order behind that use
        ZL_CK(cudaEventRecord(f->evUser, stream));
        ZL_CK(cudaStreamWaitEvent(f->wf->chain, f->evUser, 0));
        ZL_CK(cudaStreamWaitEvent(f->wf2->chain, f->evUser, 0));
        if (f->wf3 && f->wf3->chain) ZL_CK(cudaStreamWaitEvent(f->wf3->chain, f->evUser, 0));
        if (f->wf4 && f->wf4->chain) ZL_CK(cudaStreamWaitEvent(f->wf4->chain, f->evUser, 0));
        ZL_CK(cudaStreamWaitEvent(R, f->evUser, 0));
        f->pipeDirty = true;
    }
    f->readSincePass = false;
    if (w.passInFlight) ZL_CK(cudaStreamWaitEvent(M, w.evPassResolved, 0));     // the pass before last read this workspace until its last resolve
    WfState ws = w.st;
    ws.sortMode = o.sortMode;
    ws.fusedKeys = fused ? 1 : 0;
    ZL_CK(cudaMemsetAsync(w.st.cnt, 0, kWfCounters * sizeof(int), M));
    const bool dense = (size_t)f->w * f->h >= ((size_t)1 << 20) && !std::getenv("ZL_WF_NO_DENSE");      // 56-register stage kernels (zl_wavefront.cuh)
    if (dense) wfGenerateDenseKernel<<<(w.st.nSlots + 127) / 128, 128, 0, M>>>(s->d, *p, w.st);
    else wfGenerateKernel<<<(w.st.nSlots + 127) / 128, 128, 0, M>>>(s->d, *p, w.st);
    ZL_LAUNCHED();
    for (int b = 0; b <= p->maxDepth; b++) {
        if (b > 0) {
            if (fused) ZL_CK(cudaMemsetAsync(w.st.hist, 0, w.histInts * sizeof(int), M));
            if (s->binMask & 1u) {
                if (dense) wfShadeDenseKernel<<<w.gridShadeDense, 128, 0, M>>>(s->d, *p, ws, f->d, b);
                else wfShadeKernel<0><<<w.gridShade[0], 128, 0, M>>>(s->d, *p, ws, f->d, b);
                ZL_LAUNCHED();
            }
            if (s->binMask & 2u) { wfShadeKernel<1><<<w.gridShade[1], 128, 0, M>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 4u) { wfShadeKernel<2><<<w.gridShade[2], 128, 0, M>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 8u) { wfShadeKernel<3><<<w.gridShade[3], 128, 0, M>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 16u) { wfShadeKernel<4><<<w.gridShade[4], 128, 0, M>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
        }
        if (int rc = wfTraceStage<0>(s, f, o, b, b == p->maxDepth ? 1 : 0, b > 0, 1e-4f, M, fused, &w)) return rc;
        ZL_CK(cudaEventRecord(w.evTraced, M));
        ZL_CK(cudaStreamWaitEvent(R, w.evTraced, 0));
        if (dense) wfResolveDenseKernel<<<w.gridResolveDense, 128, 0, R>>>(s->d, *p, ws, f->d, b);
        else wfResolveKernel<<<w.gridResolve, 128, 0, R>>>(s->d, *p, ws, f->d, b);
        ZL_LAUNCHED();
    }
    ZL_CK(cudaEventRecord(w.evPassResolved, R));
    ZL_CK(cudaEventRecord(f->evTail, R));
    w.passInFlight = true;
    f->pipePasses++;
    return 0;
}

// adjoint light tracer, wavefront form (zl_wavefront_light.cuh)
static int launchWavefrontLightPass(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, cudaStream_t stream) {
    const long long total = (long long)ZL_LIGHT_GROUP_SIZE * p->blocksOnePass;
    if (int rc = wfEnsure(f, (size_t)total)) return rc;
    WfWorkspace& w = *f->wf;
    const WfOptions o;
    const bool fused = wfSortEnabled(s, o) && o.fuseSortKeys;      // shade kernels record the sort keys of the rays they queue
    const bool overlap = o.overlap != 0 && !g_stageTimer.enabled;  // shade kernels of the material types side by side (WfFan)
    if (overlap) { if (int rc = wfEnsureSide(w)) return rc; }
    WfState ws = w.st;
    ws.sortMode = o.sortMode;
    ws.fusedKeys = fused ? 1 : 0;
    ws.passCounters = g_wfPassCounters;      // non-null while this pass is being captured into a CUDA graph (graphPass)
    ZL_CK(cudaMemsetAsync(w.st.cnt, 0, kWfCounters * sizeof(int), stream));
    { StageScope scope(ZL_STAGE_GENERATE, stream);
    wfLightGenerateKernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(s->d, *p, ws, total,
                                                                              (uint32_t)ZL_LIGHT_GROUP_SIZE * (uint32_t)p->blocksOnePass, 0);
    ZL_LAUNCHED(); }
    for (int b = 0; b <= p->maxDepth; b++) {
        if (b > 0) {
            StageScope scope(ZL_STAGE_SHADE, stream);
            if (fused) { if (int rc = wfSortClearHistogram(w, stream)) return rc; }
            WfFan fan(w, overlap, stream);
            if (s->binMask & 1u) { wfLightShadeKernel<0><<<w.gridLightShade[0], 128, 0, fan.next()>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
            if (s->binMask & 2u) { wfLightShadeKernel<1><<<w.gridLightShade[1], 128, 0, fan.next()>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
            if (s->binMask & 4u) { wfLightShadeKernel<2><<<w.gridLightShade[2], 128, 0, fan.next()>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
            if (s->binMask & 8u) { wfLightShadeKernel<3><<<w.gridLightShade[3], 128, 0, fan.next()>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
            if (s->binMask & 16u) { wfLightShadeKernel<4><<<w.gridLightShade[4], 128, 0, fan.next()>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
            if (int rc = fan.join()) return rc;
        }
        if (int rc = wfTraceStage<1>(s, f, o, b, 0, true, 0.0f, stream, fused && b > 0)) return rc;
    }
    return 0;
}

// Light tracer with two passes in flight (variant 2).  Its film writes are the splats of the trace kernels — float atomics.  Each pass
// splats into a film-sized buffer of its workspace (zero between passes); when the pass is complete, the film stream adds the buffer to
// the film (mergeSplatKernel, in pass order).  So the film only ever holds whole passes, a frame read on the film stream is a consistent
// snapshot, and — unlike round 1's schedule, where a pass launched after a read had to wait for it and for the pass before it — a frame
// read back every pass no longer serialises the passes (Cornell 1080p end to end: 859 -> see DESIGN.md §5).
static int launchWavefrontLightPassPipelined(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, cudaStream_t stream) {
    const long long total = (long long)ZL_LIGHT_GROUP_SIZE * p->blocksOnePass;
    const int depth = pipeDepth(f);
    if (int rc = wfEnsure(f, (size_t)total)) return rc;
    if (int rc = wfEnsure(f, (size_t)total, true)) return rc;
    if (depth >= 3) { if (int rc = wfEnsure(f, (size_t)total, false, 1)) return rc; }
    if (depth >= 4) { if (int rc = wfEnsure(f, (size_t)total, false, 2)) return rc; }
    if (int rc = pipeEnsure(f, depth)) return rc;
    WfWorkspace* const ring[4] = {f->wf, f->wf2, f->wf3, f->wf4};
    WfWorkspace& w = *ring[f->pathRing++ % (unsigned)depth];
    const cudaStream_t M = w.chain, R = f->filmStream;
    const size_t pixels = (size_t)f->w * f->h;
    if (!w.splats || w.splatPixels != pixels) {
        if (w.splats) { cudaDeviceSynchronize(); cudaFree(w.splats); w.splats = nullptr; }
        ZL_CK(cudaMalloc((void**)&w.splats, pixels * sizeof(float4)));
        ZL_CK(cudaMemset(w.splats, 0, pixels * sizeof(float4)));
        ZL_CK(cudaDeviceSynchronize());      // the chain streams are non-blocking: they do not order themselves behind the legacy stream's memset
        w.splatPixels = pixels;
        if (!w.evMerged) ZL_CK(cudaEventCreateWithFlags(&w.evMerged, cudaEventDisableTiming));
        w.mergePending = false;
    }
    const WfOptions o;
    const bool fused = wfSortEnabled(s, o) && o.fuseSortKeys;
    if (!f->pipeDirty) {
        ZL_CK(cudaEventRecord(f->evUser, stream));
        for (WfWorkspace* x : ring) if (x && x->chain) ZL_CK(cudaStreamWaitEvent(x->chain, f->evUser, 0));
        ZL_CK(cudaStreamWaitEvent(R, f->evUser, 0));
        f->pipeDirty = true;
    }
    f->readSincePass = false;
    if (w.mergePending) ZL_CK(cudaStreamWaitEvent(M, w.evMerged, 0));      // the pass before last on this workspace: its splats are merged and the buffer is zero again
    WfState ws = w.st;
    ws.sortMode = o.sortMode;
    ws.fusedKeys = fused ? 1 : 0;
    ZL_CK(cudaMemsetAsync(w.st.cnt, 0, kWfCounters * sizeof(int), M));
    wfLightGenerateKernel<<<(unsigned)((total + 127) / 128), 128, 0, M>>>(s->d, *p, w.st, total, (uint32_t)ZL_LIGHT_GROUP_SIZE * (uint32_t)p->blocksOnePass, 0);
    ZL_LAUNCHED();
    float4* const filmPtr = f->d;
    f->d = w.splats;                    // the trace kernels of this pass take their splat target from f->d at launch time
    int rcTrace = 0;
    for (int b = 0; b <= p->maxDepth && rcTrace == 0; b++) {
        if (b > 0) {
            if (fused) { if (cudaMemsetAsync(w.st.hist, 0, w.histInts * sizeof(int), M) != cudaSuccess) { rcTrace = ZL_ERR_INVALID_ARGUMENT; break; } }
            if (s->binMask & 1u) { wfLightShadeKernel<0><<<w.gridLightShade[0], 128, 0, M>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
            if (s->binMask & 2u) { wfLightShadeKernel<1><<<w.gridLightShade[1], 128, 0, M>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
            if (s->binMask & 4u) { wfLightShadeKernel<2><<<w.gridLightShade[2], 128, 0, M>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
            if (s->binMask & 8u) { wfLightShadeKernel<3><<<w.gridLightShade[3], 128, 0, M>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
            if (s->binMask & 16u) { wfLightShadeKernel<4><<<w.gridLightShade[4], 128, 0, M>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
        }
        rcTrace = wfTraceStage<1>(s, f, o, b, 0, true, 0.0f, M, fused && b > 0, &w);
    }
    f->d = filmPtr;
    if (rcTrace) return rcTrace;
    ZL_CK(cudaEventRecord(w.evPassResolved, M));            // "this pass has splatted everything"
    ZL_CK(cudaStreamWaitEvent(R, w.evPassResolved, 0));
    mergeSplatKernel<<<w.sms * 8, 256, 0, R>>>(f->d, w.splats, pixels);
    ZL_LAUNCHED();
    ZL_CK(cudaEventRecord(w.evMerged, R));
    ZL_CK(cudaEventRecord(f->evTail, R));
    w.mergePending = true;
    w.passInFlight = true;
    f->pipePasses++;
    return 0;
}

// triple tracer, camera pass (s = 0 / s = 1 strategies), wavefront form (zl_wavefront_triple.cuh)
static int launchWavefrontTriplePtPass(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, cudaStream_t stream) {
    if (int rc = wfEnsure(f)) return rc;
    WfWorkspace& w = *f->wf;
    const WfOptions o;
    const bool fused = wfSortEnabled(s, o) && o.fuseSortKeys;      // shade kernels record the sort keys of the rays they queue
    const bool overlap = o.overlap != 0 && !g_stageTimer.enabled;  // same schedule as launchWavefrontPathPass
    if (overlap) { if (int rc = wfEnsureSide(w)) return rc; }
    WfState ws = w.st;
    ws.sortMode = o.sortMode;
    ws.fusedKeys = fused ? 1 : 0;
    ws.passCounters = g_wfPassCounters;      // non-null while this pass is being captured into a CUDA graph (graphPass)
    ZL_CK(cudaMemsetAsync(w.st.cnt, 0, kWfCounters * sizeof(int), stream));
    { StageScope scope(ZL_STAGE_GENERATE, stream);
    wfGenerateKernel<<<(w.st.nSlots + 127) / 128, 128, 0, stream>>>(s->d, *p, ws);
    ZL_LAUNCHED(); }
    for (int b = 0; b <= p->maxDepth; b++) {
        if (b > 0) {
            StageScope scope(ZL_STAGE_SHADE, stream);
            if (fused) { if (int rc = wfSortClearHistogram(w, stream)) return rc; }
            WfFan fan(w, overlap, stream);
            if (s->binMask & 1u) { wfTripleShadeKernel<0><<<w.gridTripleShade[0], 128, 0, fan.next()>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 2u) { wfTripleShadeKernel<1><<<w.gridTripleShade[1], 128, 0, fan.next()>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 4u) { wfTripleShadeKernel<2><<<w.gridTripleShade[2], 128, 0, fan.next()>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 8u) { wfTripleShadeKernel<3><<<w.gridTripleShade[3], 128, 0, fan.next()>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 16u) { wfTripleShadeKernel<4><<<w.gridTripleShade[4], 128, 0, fan.next()>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (int rc = fan.join()) return rc;
        }
        WfOptions ob = o;
        ob.simpleMask = 3;     // the regenerating kernel knows only the path tracer's 1e-4 shadow offset
        if (int rc = wfTraceStage<0>(s, f, ob, b, b == p->maxDepth ? 1 : 0, b > 0, 1e-5f, stream, fused)) return rc;    // visible(): origin + 1e-5 * dir
        StageScope scope(ZL_STAGE_RESOLVE, stream);
        const cudaStream_t rs = overlap ? w.side[0] : stream;
        if (overlap) { if (int rc = WfResolveSide::before(w, stream)) return rc; }
        if (b == 0) wfResolveKernel<<<w.gridResolve, 128, 0, rs>>>(s->d, *p, ws, f->d, b);                 // primary miss -> envLe, emitter -> lightLe
        else wfTripleResolveKernel<<<w.gridTripleResolve, 128, 0, rs>>>(s->d, *p, ws, f->d, b);
        ZL_LAUNCHED();
        if (overlap) { if (int rc = WfResolveSide::after(w, b)) return rc; }
    }
    if (overlap)
        for (int k = 0; k < 2; k++) { if (int rc = WfResolveSide::waitParity(w, k, stream)) return rc; }
    return 0;
}

// triple tracer, light pass (t = 1 strategy), wavefront form: uLoopsPerPass rounds, RNG streams carried in smp
static int launchWavefrontTripleLptPass(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, cudaStream_t stream) {
    const long long total = (long long)ZL_LIGHT_GROUP_SIZE * p->blocksOnePass;
    if (int rc = wfEnsure(f, (size_t)total)) return rc;
    WfWorkspace& w = *f->wf;
    const WfOptions o;
    const bool fused = wfSortEnabled(s, o) && o.fuseSortKeys;      // shade kernels record the sort keys of the rays they queue
    const bool overlap = o.overlap != 0 && !g_stageTimer.enabled;  // shade kernels of the material types side by side (WfFan)
    if (overlap) { if (int rc = wfEnsureSide(w)) return rc; }
    WfState ws = w.st;
    ws.sortMode = o.sortMode;
    ws.fusedKeys = fused ? 1 : 0;
    ws.passCounters = g_wfPassCounters;      // non-null while this pass is being captured into a CUDA graph (graphPass)
    const uint32_t seedMul = (uint32_t)ZL_LIGHT_GROUP_SIZE * (uint32_t)p->blocksOnePass * (uint32_t)p->loopsPerPass;
    for (int loop = 0; loop < p->loopsPerPass; loop++) {
        ZL_CK(cudaMemsetAsync(w.st.cnt, 0, kWfCounters * sizeof(int), stream));
        { StageScope scope(ZL_STAGE_GENERATE, stream);
        wfTripleLightGenerateKernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(s->d, *p, ws, total, seedMul, loop > 0 ? 1 : 0);
        ZL_LAUNCHED(); }
        for (int b = 0; b <= p->maxDepth; b++) {
            if (b > 0) {
                StageScope scope(ZL_STAGE_SHADE, stream);
                if (fused) { if (int rc = wfSortClearHistogram(w, stream)) return rc; }
                WfFan fan(w, overlap, stream);
                if (s->binMask & 1u) { wfTripleLightShadeKernel<0><<<w.gridTripleLightShade[0], 128, 0, fan.next()>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
                if (s->binMask & 2u) { wfTripleLightShadeKernel<1><<<w.gridTripleLightShade[1], 128, 0, fan.next()>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
                if (s->binMask & 4u) { wfTripleLightShadeKernel<2><<<w.gridTripleLightShade[2], 128, 0, fan.next()>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
                if (s->binMask & 8u) { wfTripleLightShadeKernel<3><<<w.gridTripleLightShade[3], 128, 0, fan.next()>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
                if (s->binMask & 16u) { wfTripleLightShadeKernel<4><<<w.gridTripleLightShade[4], 128, 0, fan.next()>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
                if (int rc = fan.join()) return rc;
            }
            if (int rc = wfTraceStage<1>(s, f, o, b, 0, true, 0.0f, stream, fused && b > 0)) return rc;
        }
    }
    return 0;
}

// Triple tracer with two passes in flight (variant 2).  A pass = camera pass (PT) + light pass (LPT) on ONE chain / workspace:
//   chain c:  PT(k) stages ............ [wait: PT(k) resolved]  LPT(k) stages (splats: float atomics)   record "LPT(k) done"
//   R      :  resolve(b) of PT(k) ...                           wait "LPT(k) done"  | later: frame reads, resolves of PT(k+1)
// so the plain read-modify-write film adds of the resolve kernels never run next to the atomic splats of a light pass, a
// frame read sees whole passes only, and the bulk of PT(k+1) (generate / shade / trace on the other chain) overlaps the tail
// of PT(k) and all of LPT(k).
static int launchWavefrontTriplePtPassPipelined(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, cudaStream_t stream) {
    const int depth = tripleDepth(f);
    if (int rc = pipeEnsure(f, depth)) return rc;
    if (f->tripleHalf) { f->pipePasses++; f->tripleHalf = false; }      // a camera pass without its light pass: move on
    WfWorkspace* const ring[4] = {f->wf, f->wf2, f->wf3, f->wf4};
    WfWorkspace& w = *ring[f->pipePasses % (unsigned long long)depth];
    const cudaStream_t M = w.chain, R = f->filmStream;
    const WfOptions o;
    const bool fused = wfSortEnabled(s, o) && o.fuseSortKeys;
    if (!f->pipeDirty) {
        ZL_CK(cudaEventRecord(f->evUser, stream));
        for (WfWorkspace* x : ring) if (x && x->chain) ZL_CK(cudaStreamWaitEvent(x->chain, f->evUser, 0));
        ZL_CK(cudaStreamWaitEvent(R, f->evUser, 0));
        f->pipeDirty = true;
    }
    f->readSincePass = false;
    if (w.passInFlight) ZL_CK(cudaStreamWaitEvent(M, w.evPassResolved, 0));
    WfState ws = w.st;
    ws.sortMode = o.sortMode;
    ws.fusedKeys = fused ? 1 : 0;
    ZL_CK(cudaMemsetAsync(w.st.cnt, 0, kWfCounters * sizeof(int), M));
    wfGenerateKernel<<<(w.st.nSlots + 127) / 128, 128, 0, M>>>(s->d, *p, w.st);
    ZL_LAUNCHED();
    WfOptions ob = o;
    ob.simpleMask = 3;
    for (int b = 0; b <= p->maxDepth; b++) {
        if (b > 0) {
            if (fused) ZL_CK(cudaMemsetAsync(w.st.hist, 0, w.histInts * sizeof(int), M));
            if (s->binMask & 1u) { wfTripleShadeKernel<0><<<w.gridTripleShade[0], 128, 0, M>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 2u) { wfTripleShadeKernel<1><<<w.gridTripleShade[1], 128, 0, M>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 4u) { wfTripleShadeKernel<2><<<w.gridTripleShade[2], 128, 0, M>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 8u) { wfTripleShadeKernel<3><<<w.gridTripleShade[3], 128, 0, M>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
            if (s->binMask & 16u) { wfTripleShadeKernel<4><<<w.gridTripleShade[4], 128, 0, M>>>(s->d, *p, ws, f->d, b); ZL_LAUNCHED(); }
        }
        if (int rc = wfTraceStage<0>(s, f, ob, b, b == p->maxDepth ? 1 : 0, b > 0, 1e-5f, M, fused, &w)) return rc;
        ZL_CK(cudaEventRecord(w.evTraced, M));
        ZL_CK(cudaStreamWaitEvent(R, w.evTraced, 0));
        if (b == 0) wfResolveKernel<<<w.gridResolve, 128, 0, R>>>(s->d, *p, ws, f->d, b);
        else wfTripleResolveKernel<<<w.gridTripleResolve, 128, 0, R>>>(s->d, *p, ws, f->d, b);
        ZL_LAUNCHED();
    }
    ZL_CK(cudaEventRecord(w.evPassResolved, R));
    ZL_CK(cudaEventRecord(f->evTail, R));
    w.passInFlight = true;
    f->tripleHalf = true;
    return 0;
}
static int launchWavefrontTripleLptPassPipelined(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, cudaStream_t stream) {
    const long long total = (long long)ZL_LIGHT_GROUP_SIZE * p->blocksOnePass;
    const int depth = tripleDepth(f);
    if (!f->wf || !f->wf2 || f->wf->capacity < (size_t)total || f->wf2->capacity < (size_t)total ||
        (depth >= 3 && (!f->wf3 || f->wf3->capacity < (size_t)total))) {       // first pass: grow the workspaces (synchronises)
        const bool half = f->tripleHalf;
        if (int rc = wfEnsure(f, (size_t)total)) return rc;
        if (int rc = wfEnsure(f, (size_t)total, true)) return rc;
        if (depth >= 3) { if (int rc = wfEnsure(f, (size_t)total, false, 1)) return rc; }
        f->tripleHalf = half;
    }
    if (int rc = pipeEnsure(f, depth)) return rc;
    WfWorkspace* const ring[4] = {f->wf, f->wf2, f->wf3, f->wf4};
    WfWorkspace& w = *ring[f->pipePasses % (unsigned long long)depth];
    const cudaStream_t M = w.chain, R = f->filmStream;
    const WfOptions o;
    const bool fused = wfSortEnabled(s, o) && o.fuseSortKeys;
    if (!f->pipeDirty) {
        ZL_CK(cudaEventRecord(f->evUser, stream));
        for (WfWorkspace* x : ring) if (x && x->chain) ZL_CK(cudaStreamWaitEvent(x->chain, f->evUser, 0));
        ZL_CK(cudaStreamWaitEvent(R, f->evUser, 0));
        f->pipeDirty = true;
    }
    // the splats follow everything queued on the film stream so far: the resolves of this pass's camera pass, and any frame read
    ZL_CK(cudaEventRecord(f->evTail, R));
    ZL_CK(cudaStreamWaitEvent(M, f->evTail, 0));
    f->readSincePass = false;
    WfState ws = w.st;
    ws.sortMode = o.sortMode;
    ws.fusedKeys = fused ? 1 : 0;
    const uint32_t seedMul = (uint32_t)ZL_LIGHT_GROUP_SIZE * (uint32_t)p->blocksOnePass * (uint32_t)p->loopsPerPass;
    for (int loop = 0; loop < p->loopsPerPass; loop++) {
        ZL_CK(cudaMemsetAsync(w.st.cnt, 0, kWfCounters * sizeof(int), M));
        wfTripleLightGenerateKernel<<<(unsigned)((total + 127) / 128), 128, 0, M>>>(s->d, *p, w.st, total, seedMul, loop > 0 ? 1 : 0);
        ZL_LAUNCHED();
        for (int b = 0; b <= p->maxDepth; b++) {
            if (b > 0) {
                if (fused) ZL_CK(cudaMemsetAsync(w.st.hist, 0, w.histInts * sizeof(int), M));
                if (s->binMask & 1u) { wfTripleLightShadeKernel<0><<<w.gridTripleLightShade[0], 128, 0, M>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
                if (s->binMask & 2u) { wfTripleLightShadeKernel<1><<<w.gridTripleLightShade[1], 128, 0, M>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
                if (s->binMask & 4u) { wfTripleLightShadeKernel<2><<<w.gridTripleLightShade[2], 128, 0, M>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
                if (s->binMask & 8u) { wfTripleLightShadeKernel<3><<<w.gridTripleLightShade[3], 128, 0, M>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
                if (s->binMask & 16u) { wfTripleLightShadeKernel<4><<<w.gridTripleLightShade[4], 128, 0, M>>>(s->d, *p, ws, b); ZL_LAUNCHED(); }
            }
            if (int rc = wfTraceStage<1>(s, f, o, b, 0, true, 0.0f, M, fused && b > 0, &w)) return rc;
        }
    }
    // "this pass has splatted everything": the film stream (frame reads, the next camera pass's resolves) and the next use of
    // this workspace wait for it
    ZL_CK(cudaEventRecord(w.evPassResolved, M));
    ZL_CK(cudaStreamWaitEvent(R, w.evPassResolved, 0));
    ZL_CK(cudaEventRecord(f->evTail, R));
    w.passInFlight = true;
    f->tripleHalf = false;
    f->pipePasses++;
    return 0;
}

extern "C" {

static int launchPathPassImpl(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, int variant, void* stream) {
    if (int rc = checkPass(s, f, p, "zl_launch_path_pass")) return rc;
    if (variant < 0 || variant > 3) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_launch_path_pass: variant must be 0 (megakernel), 1 (wavefront), 2 (wavefront, passes pipelined) or 3 (wavefront, replayed CUDA graph)");
    const bool wavefront = variant >= 1 && p->maxDepth >= 1 && p->maxDepth <= kWfMaxDepth;
    if (wavefront && variant == 2 && !g_stageTimer.enabled) return launchWavefrontPathPassPipelined(s, f, p, (cudaStream_t)stream);
    if (int rc = pipeFlush(f, (cudaStream_t)stream)) return rc;      // a pass of another variant after pipelined ones
    if (wavefront && variant == 3) return graphPass(s, f, p, (cudaStream_t)stream, 0, [&](cudaStream_t q) { return launchWavefrontPathPass(s, f, p, q); });
    if (wavefront) return launchWavefrontPathPass(s, f, p, (cudaStream_t)stream);
    dim3 grid((p->filmW + kTileW - 1) / kTileW, (p->filmH + kTileH - 1) / kTileH);
    StageScope scope(ZL_STAGE_MEGAKERNEL, (cudaStream_t)stream);
    pathPassKernel<<<grid, kPixelBlock, 0, (cudaStream_t)stream>>>(s->d, *p, f->d);
    ZL_LAUNCHED();
    return 0;
}
static int launchTriplePtPassImpl(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, int variant, void* stream) {
    if (int rc = checkPass(s, f, p, "zl_launch_triple_pt_pass")) return rc;
    if (variant < 0 || variant > 3) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_launch_triple_pt_pass: variant must be 0 (megakernel), 1 (wavefront), 2 (wavefront, passes pipelined) or 3 (wavefront, replayed CUDA graph)");
    // triple_path_pass_pt.glsl samples an area light at every vertex unconditionally (:112-130) and the LPT pass has no other
    // emitter: without area lights the reference reads past its light tables.  Refuse loudly instead of rendering nothing.
    if (s->d.numLightTriangles <= 0) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_launch_triple_pt_pass: the triple tracer needs at least one area light");
    const bool wavefront = variant >= 1 && p->maxDepth >= 1 && p->maxDepth <= kWfMaxDepth;
    if (wavefront && variant == 2 && !g_stageTimer.enabled) return launchWavefrontTriplePtPassPipelined(s, f, p, (cudaStream_t)stream);
    if (int rc = pipeFlush(f, (cudaStream_t)stream)) return rc;
    if (wavefront && variant == 3) return graphPass(s, f, p, (cudaStream_t)stream, 2, [&](cudaStream_t q) { return launchWavefrontTriplePtPass(s, f, p, q); });
    if (wavefront) return launchWavefrontTriplePtPass(s, f, p, (cudaStream_t)stream);
    dim3 grid((p->filmW + kTileW - 1) / kTileW, (p->filmH + kTileH - 1) / kTileH);
    StageScope scope(ZL_STAGE_MEGAKERNEL, (cudaStream_t)stream);
    triplePtPassKernel<<<grid, kPixelBlock, 0, (cudaStream_t)stream>>>(s->d, *p, f->d);
    ZL_LAUNCHED();
    return 0;
}
static int launchLightPassImpl(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, int variant, void* stream) {
    if (int rc = checkPass(s, f, p, "zl_launch_light_pass")) return rc;
    if (variant < 0 || variant > 3) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_launch_light_pass: variant must be 0 (megakernel), 1 (wavefront), 2 (wavefront, passes pipelined) or 3 (wavefront, replayed CUDA graph)");
    if (s->d.numLightTriangles <= 0 || p->blocksOnePass <= 0) return 0;
    const bool wavefront = variant >= 1 && p->maxDepth >= 0 && p->maxDepth <= kWfMaxDepth;
    if (wavefront && variant == 2 && !g_stageTimer.enabled) return launchWavefrontLightPassPipelined(s, f, p, (cudaStream_t)stream);
    if (int rc = pipeFlush(f, (cudaStream_t)stream)) return rc;
    if (wavefront && variant == 3) return graphPass(s, f, p, (cudaStream_t)stream, 1, [&](cudaStream_t q) { return launchWavefrontLightPass(s, f, p, q); });
    if (wavefront) return launchWavefrontLightPass(s, f, p, (cudaStream_t)stream);
    long long total = (long long)ZL_LIGHT_GROUP_SIZE * p->blocksOnePass;
    unsigned blocks = (unsigned)((total + kLightBlock - 1) / kLightBlock);
    StageScope scope(ZL_STAGE_MEGAKERNEL, (cudaStream_t)stream);
    lightPassKernel<<<blocks, kLightBlock, 0, (cudaStream_t)stream>>>(s->d, *p, f->d, total);
    ZL_LAUNCHED();
    return 0;
}
static int launchTripleLptPassImpl(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, int variant, void* stream) {
    if (int rc = checkPass(s, f, p, "zl_launch_triple_lpt_pass")) return rc;
    if (variant < 0 || variant > 3) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_launch_triple_lpt_pass: variant must be 0 (megakernel), 1 (wavefront), 2 (wavefront, passes pipelined) or 3 (wavefront, replayed CUDA graph)");
    if (s->d.numLightTriangles <= 0 || p->blocksOnePass <= 0) return 0;
    const bool wavefront = variant >= 1 && p->maxDepth >= 0 && p->maxDepth <= kWfMaxDepth;
    if (wavefront && variant == 2 && !g_stageTimer.enabled) return launchWavefrontTripleLptPassPipelined(s, f, p, (cudaStream_t)stream);
    if (int rc = pipeFlush(f, (cudaStream_t)stream)) return rc;
    if (wavefront && variant == 3) return graphPass(s, f, p, (cudaStream_t)stream, 3, [&](cudaStream_t q) { return launchWavefrontTripleLptPass(s, f, p, q); });
    if (wavefront) return launchWavefrontTripleLptPass(s, f, p, (cudaStream_t)stream);
    long long total = (long long)ZL_LIGHT_GROUP_SIZE * p->blocksOnePass;
    unsigned blocks = (unsigned)((total + kLightBlock - 1) / kLightBlock);
    StageScope scope(ZL_STAGE_MEGAKERNEL, (cudaStream_t)stream);
    tripleLptPassKernel<<<blocks, kLightBlock, 0, (cudaStream_t)stream>>>(s->d, *p, f->d, total);
    ZL_LAUNCHED();
    return 0;
}

// every pass launch: the launch itself, then the D2H copy calls of read-backs whose frames were resolved before it (filmDownloadAsync)
extern "C++" {
template <class Impl>
static int launchPassThenCopies(ZlFilm* f, Impl impl) {
    const int rc = impl();
    if (rc != 0 || !f) return rc;
    f->passesLaunched++;
    return filmIssuePendingCopies(f);
}
}
int zl_launch_path_pass(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, int variant, void* stream) { return launchPassThenCopies(f, [&] { return launchPathPassImpl(s, f, p, variant, stream); }); }
int zl_launch_triple_pt_pass(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, int variant, void* stream) { return launchPassThenCopies(f, [&] { return launchTriplePtPassImpl(s, f, p, variant, stream); }); }
int zl_launch_light_pass(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, int variant, void* stream) { return launchPassThenCopies(f, [&] { return launchLightPassImpl(s, f, p, variant, stream); }); }
int zl_launch_triple_lpt_pass(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, int variant, void* stream) { return launchPassThenCopies(f, [&] { return launchTripleLptPassImpl(s, f, p, variant, stream); }); }

static unsigned long long g_lastSkipped[3] = {0, 0, 0};
int zl_counted_pass_untraced(unsigned long long* counters3) {
    if (!counters3) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_counted_pass_untraced: null counters");
    std::memcpy(counters3, g_lastSkipped, sizeof g_lastSkipped);
    return 0;
}
int zl_counted_pass(ZlScene* s, ZlFilm* f, const ZlRenderParams* p, int kind, unsigned long long* counters6) {
    if (int rc = checkPass(s, f, p, "zl_counted_pass")) return rc;
    if (!counters6) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_counted_pass: null counters");
    if (f->pipeDirty) { ZL_CK(cudaDeviceSynchronize()); f->pipeDirty = false; }      // runs on the null stream: wait for pipelined passes in flight
    if ((kind == 1 || kind == 2 || kind == 3) && s->d.numLightTriangles <= 0) { std::memset(counters6, 0, 48); return 0; }
    unsigned long long* dc = nullptr;
    ZL_CK(cudaMalloc((void**)&dc, 12 * sizeof(unsigned long long)));
    cudaMemset(dc, 0, 12 * sizeof(unsigned long long));
    DScene d = s->d;
    d.counters = dc;
    int bad = zlc::launchCountedPass(kind, reinterpret_cast<const zlc::DScene&>(d), *p, f->d, nullptr);   // same layout, other namespace
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(counters6, dc, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(g_lastSkipped, dc + 6, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(dc);
    if (bad) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_counted_pass: unknown kind");
    if (e != cudaSuccess) return fail((int)e, std::string("zl_counted_pass: ") + cudaGetErrorString(e));
    return 0;
}

// ---- explicit ray sets ----
int zl_rayset_create(const float* raysHost, size_t n, ZlRaySet** out) {
    if (!raysHost || !n || !out) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_rayset_create: bad argument");
    auto* r = new ZlRaySet();
    r->n = n;
    std::vector<float4> packed(2 * n);
    for (size_t i = 0; i < n; i++) {
        packed[2 * i] = make_float4(raysHost[6 * i], raysHost[6 * i + 1], raysHost[6 * i + 2], 1e8f);
        packed[2 * i + 1] = make_float4(raysHost[6 * i + 3], raysHost[6 * i + 4], raysHost[6 * i + 5], 0.0f);
    }
    cudaError_t e;
    if ((e = cudaMalloc((void**)&r->rays, 2 * n * sizeof(float4))) != cudaSuccess || (e = cudaMalloc((void**)&r->ids, n * 4)) != cudaSuccess ||
        (e = cudaMalloc((void**)&r->t, n * 4)) != cudaSuccess ||
        (e = cudaMemcpy(r->rays, packed.data(), 2 * n * sizeof(float4), cudaMemcpyHostToDevice)) != cudaSuccess) {
        zl_rayset_destroy(r);
        return fail((int)e, std::string("zl_rayset_create: ") + cudaGetErrorString(e));
    }
    *out = r;
    return 0;
}
int zl_rayset_destroy(ZlRaySet* r) {
    if (r) { cudaFree(r->rays); cudaFree(r->ids); cudaFree(r->t); delete r; }
    return 0;
}
int zl_rayset_create_primary(const ZlRenderParams* p, ZlRaySet** out) {
    if (!p || !out || p->filmW <= 0 || p->filmH <= 0) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_rayset_create_primary: bad argument");
    auto* r = new ZlRaySet();
    r->n = (size_t)p->filmW * p->filmH; r->tileW = p->filmW; r->tileH = p->filmH;
    cudaError_t e;
    if ((e = cudaMalloc((void**)&r->rays, 2 * r->n * sizeof(float4))) != cudaSuccess || (e = cudaMalloc((void**)&r->ids, r->n * 4)) != cudaSuccess ||
        (e = cudaMalloc((void**)&r->t, r->n * 4)) != cudaSuccess) {
        zl_rayset_destroy(r);
        return fail((int)e, std::string("zl_rayset_create_primary: ") + cudaGetErrorString(e));
    }
    unsigned blocks = (unsigned)((r->n + 255) / 256);
    primaryRaysKernel<<<blocks, 256>>>(*p, r->rays);
    g_launches++;
    if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) {
        zl_rayset_destroy(r);
        return fail((int)e, std::string("zl_rayset_create_primary: ") + cudaGetErrorString(e));
    }
    *out = r;
    return 0;
}
}  // extern "C"
// A/B switch ZL_OCTANT_WALK (read per launch, like the ZL_WF_* switches): 0 = general walk only (zl_traverse.cuh traverseWarp)
static DScene sceneWithWalkSwitch(const ZlScene* s) {
    DScene d = s->d;
    if (const char* e = std::getenv("ZL_OCTANT_WALK")) { const int v = std::atoi(e); if (v >= 0) d.octantWalk = std::min(2, v); }
    d.bvh2 = bvh2WalkEnabled() ? s->bvh2 : nullptr;
    return d;
}
// explicit ray sets: the instantiation of the walk for this scene's default configuration (LEAN, zl_traverse.cuh traverseWarp), the general one otherwise
template <bool ANYHIT>
static void launchTraceRays(const DScene& dS, unsigned blocks, cudaStream_t stream, const float4* rays, size_t n, int tileW, int tileH, int32_t* ids, float* t) {
    const bool plain = !dS.bvh2 && !dS.nodePolicy;
    if (plain && dS.octantWalk == 1) traceRaysKernel<ANYHIT, false, 1><<<blocks, kTraceBlock, 0, stream>>>(dS, rays, n, tileW, tileH, ids, t, nullptr);
    else if (plain && dS.octantWalk == 2) traceRaysKernel<ANYHIT, false, 2><<<blocks, kTraceBlock, 0, stream>>>(dS, rays, n, tileW, tileH, ids, t, nullptr);
    else traceRaysKernel<ANYHIT, false><<<blocks, kTraceBlock, 0, stream>>>(dS, rays, n, tileW, tileH, ids, t, nullptr);
}
extern "C" {
int zl_rayset_set_tmax(ZlRaySet* r, const float* tMaxHost) {
    if (!r || !tMaxHost) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_rayset_set_tmax: null argument");
    std::vector<float4> packed(2 * r->n);
    ZL_CK(cudaMemcpy(packed.data(), r->rays, 2 * r->n * sizeof(float4), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < r->n; i++) packed[2 * i].w = tMaxHost[i];
    ZL_CK(cudaMemcpy(r->rays, packed.data(), 2 * r->n * sizeof(float4), cudaMemcpyHostToDevice));
    return 0;
}
int zl_rayset_trace(ZlScene* s, ZlRaySet* r, int anyhit, int variant, void* stream) {
    if (!s || !r) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_rayset_trace: null argument");
    (void)variant;
    unsigned blocks = (unsigned)((r->n + kTraceBlock - 1) / kTraceBlock);
    const DScene dS = sceneWithWalkSwitch(s);
    if (anyhit) launchTraceRays<true>(dS, blocks, (cudaStream_t)stream, r->rays, r->n, r->tileW, r->tileH, r->ids, r->t);
    else launchTraceRays<false>(dS, blocks, (cudaStream_t)stream, r->rays, r->n, r->tileW, r->tileH, r->ids, r->t);
    ZL_LAUNCHED();
    return 0;
}
int zl_rayset_download(ZlRaySet* r, int32_t* outIds, float* outT) {
    if (!r) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_rayset_download: null argument");
    ZL_CK(cudaDeviceSynchronize());
    if (outIds) ZL_CK(cudaMemcpy(outIds, r->ids, r->n * 4, cudaMemcpyDeviceToHost));
    if (outT) ZL_CK(cudaMemcpy(outT, r->t, r->n * 4, cudaMemcpyDeviceToHost));
    return 0;
}
size_t zl_rayset_size(ZlRaySet* r) { return r ? r->n : 0; }
int zl_rayset_download_rays(ZlRaySet* r, float* raysHost) {
    if (!r || !raysHost) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_rayset_download_rays: null argument");
    std::vector<float4> packed(2 * r->n);
    ZL_CK(cudaMemcpy(packed.data(), r->rays, 2 * r->n * sizeof(float4), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < r->n; i++) {
        raysHost[6 * i] = packed[2 * i].x; raysHost[6 * i + 1] = packed[2 * i].y; raysHost[6 * i + 2] = packed[2 * i].z;
        raysHost[6 * i + 3] = packed[2 * i + 1].x; raysHost[6 * i + 4] = packed[2 * i + 1].y; raysHost[6 * i + 5] = packed[2 * i + 1].z;
    }
    return 0;
}

int zl_trace_rays(ZlScene* s, const float* rays, size_t n, int anyhit, const float* tMax, int32_t* outIds, float* outT, int32_t* outSteps) {
    if (!s || !rays || !outIds) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_trace_rays: null argument");
    if (n == 0) return 0;
    ZlRaySet* r = nullptr;
    if (int rc = zl_rayset_create(rays, n, &r)) return rc;
    int rc = 0;
    if (tMax) rc = zl_rayset_set_tmax(r, tMax);
    int2* steps = nullptr;
    if (!rc && outSteps) {
        cudaError_t e = cudaMalloc((void**)&steps, n * sizeof(int2));
        if (e != cudaSuccess) rc = fail((int)e, "zl_trace_rays: cudaMalloc(steps)");
    }
    if (!rc) {
        unsigned blocks = (unsigned)((n + kTraceBlock - 1) / kTraceBlock);
        if (outSteps) {
            if (anyhit) traceRaysKernel<true, true><<<blocks, kTraceBlock>>>(s->d, r->rays, n, 0, 0, r->ids, r->t, steps);
            else traceRaysKernel<false, true><<<blocks, kTraceBlock>>>(s->d, r->rays, n, 0, 0, r->ids, r->t, steps);
        } else {
            const DScene dS = sceneWithWalkSwitch(s);
            if (anyhit) launchTraceRays<true>(dS, blocks, nullptr, r->rays, n, 0, 0, r->ids, r->t);
            else launchTraceRays<false>(dS, blocks, nullptr, r->rays, n, 0, 0, r->ids, r->t);
        }
        g_launches++;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = fail((int)e, std::string("zl_trace_rays: ") + cudaGetErrorString(e));
    }
    if (!rc) rc = zl_rayset_download(r, outIds, outT);
    if (!rc && outSteps) {
        cudaError_t e = cudaMemcpy(outSteps, steps, n * sizeof(int2), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail((int)e, "zl_trace_rays: cudaMemcpy(steps)");
    }
    cudaFree(steps);
    zl_rayset_destroy(r);
    return rc;
}

// per-lane and per-warp-distinct record / triangle counts of the closest-hit walk over a ray set (uniqueSectorKernel)
int zl_rayset_unique_sectors(ZlScene* s, ZlRaySet* r, unsigned long long* out4) {
    if (!s || !r || !out4) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_rayset_unique_sectors: null argument");
    unsigned long long* d = nullptr;
    ZL_CK(cudaMalloc((void**)&d, 4 * sizeof(unsigned long long)));
    cudaMemset(d, 0, 4 * sizeof(unsigned long long));
    const size_t threads = r->tileW > 0 ? (size_t)((r->tileW + 7) / 8) * ((r->tileH + 3) / 4) * 32 : r->n;
    uniqueSectorKernel<<<(unsigned)((threads + kTraceBlock - 1) / kTraceBlock), kTraceBlock>>>(s->d, r->rays, r->n, r->tileW, r->tileH, d);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out4, d, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail((int)e, std::string("zl_rayset_unique_sectors: ") + cudaGetErrorString(e));
    return 0;
}

// ---- KAT evaluation ----
int zl_debug_eval(ZlScene* s, const ZlRenderParams* p, int op, const float* in, int inStride, float* out, int outStride, size_t n) {
    if (!s || !p || !in || !out || op < 0 || op >= ZL_KAT_COUNT || inStride <= 0 || outStride <= 0) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_debug_eval: bad argument");
    if (n == 0) return 0;
    float *din = nullptr, *dout = nullptr;
    ZL_CK(cudaMalloc((void**)&din, n * inStride * sizeof(float)));
    cudaError_t e = cudaMalloc((void**)&dout, n * outStride * sizeof(float));
    if (e != cudaSuccess) { cudaFree(din); return fail((int)e, "zl_debug_eval: cudaMalloc"); }
    cudaMemcpy(din, in, n * inStride * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, n * outStride * sizeof(float));
    katKernel<<<(unsigned)((n + 63) / 64), 64>>>(s->d, *p, op, din, inStride, dout, outStride, n);
    g_launches++;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, n * outStride * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(din); cudaFree(dout);
    if (e != cudaSuccess) return fail((int)e, std::string("zl_debug_eval: ") + cudaGetErrorString(e));
    return 0;
}

// ---- per-stage device timing ----
int zl_stage_timing_enable(int enable) {
    cudaDeviceSynchronize();
    g_stageTimer.clear();
    g_stageTimer.enabled = enable != 0;
    return 0;
}
int zl_stage_timing_read(double* msPerStage, unsigned long long* launchesPerStage) {
    if (!msPerStage || !launchesPerStage) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_stage_timing_read: null argument");
    ZL_CK(cudaDeviceSynchronize());
    for (int i = 0; i < ZL_STAGE_COUNT; i++) { msPerStage[i] = 0.0; launchesPerStage[i] = 0; }
    for (const auto& sp : g_stageTimer.spans) {
        float ms = 0.0f;
        ZL_CK(cudaEventElapsedTime(&ms, sp.a, sp.b));
        msPerStage[sp.stage] += ms;
        launchesPerStage[sp.stage] += sp.launches;
    }
    g_stageTimer.clear();
    return 0;
}

// ---- measured read bandwidth (L2 when `bytes` fits the 126 MB L2, HBM beyond) ----
int zl_measure_read_bandwidth(size_t bytes, int iters, double* gbPerSec) {
    if (!gbPerSec || bytes < (1u << 20) || iters <= 0) return fail(ZL_ERR_INVALID_ARGUMENT, "zl_measure_read_bandwidth: bad argument");
    size_t n16 = bytes / 16;
    uint4* buf = nullptr; unsigned* sink = nullptr;
    ZL_CK(cudaMalloc((void**)&buf, n16 * 16));
    cudaError_t e = cudaMalloc((void**)&sink, 4);
    if (e != cudaSuccess) { cudaFree(buf); return fail((int)e, "zl_measure_read_bandwidth: cudaMalloc"); }
    cudaMemset(buf, 1, n16 * 16);
    cudaMemset(sink, 0, 4);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    unsigned grid = (unsigned)sms * 8;
    for (int w = 0; w < 3; w++) { streamReadKernel<<<grid, 256>>>(buf, n16, sink); g_launches++; }
    cudaEventRecord(a);
    for (int i = 0; i < iters; i++) { streamReadKernel<<<grid, 256>>>(buf, n16, sink); g_launches++; }
    cudaEventRecord(b);
    e = cudaEventSynchronize(b);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(buf); cudaFree(sink);
    if (e != cudaSuccess) return fail((int)e, std::string("zl_measure_read_bandwidth: ") + cudaGetErrorString(e));
    *gbPerSec = (double)n16 * 16.0 * iters / (ms * 1e-3) / 1e9;
    return 0;
}

}  // extern "C"
