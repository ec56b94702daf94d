// zl_wavefront.cuh — wavefront / streaming variant of the MIS path tracer (variant 1 of
// zl_launch_path_pass).  Same per-path arithmetic as pathIntegTrace (path_integ_naive.glsl:35-143),
// cut at the ray casts into queue-driven stages so that
//   * traversal runs in a kernel of its own (~48 registers, full occupancy) over COMPACTED ray
//     queues: lanes whose path has ended do not ride along (the megakernel executes with 11 of
//     32 lanes active on the Rungholt-class scene, profiles/r1_ncu_pathPassKernel_megakernel.csv),
//   * the register-heavy BSDF code (150 registers) never waits on a BVH walk.
// The reference's own precedent is its unfinished global-queue integrator
// (src/integrator/GlobalQueuePath.cpp:100-155, pt_global_queue_{primary,streaming}.glsl), which
// reads the queue size back to the host at every depth; here the counters stay on the device and
// the stages are persistent grid-stride kernels that read them.
//
// Stages of one pass (b = 1..maxDepth):
//   generate  camera ray per pixel (8x4 pixel tiles) -> queue E of "bounce 0"; trace(0) + resolve(0) handle
//             primary hits / misses with the same code as every other bounce
//   shade(b)  [apply NEE(b-1) result, Russian roulette]  surface + material, NEE sample with the
//             shadow ray DEFERRED (DeferredVis), BSDF sample           -> queues S (shadow), E (extension), T
//   trace(b)  any-hit over S (occluded -> contribution dropped), closest-hit over E -> next queue or T
//   resolve(b) paths that end at this depth: NEE(b), environment / emitter radiance with MIS -> film
// Path state lives in slot-indexed SoA float4 arrays (slot = 8x4 tile order); queues hold slots.
#pragma once
#include "zl_integrators.cuh"

#ifndef ZL_INSTRUMENT
namespace zl {

// resident CTAs per SM the Lambertian shade / generate / resolve kernels are compiled for (register cap = 65536 / (128 * MINB)).
// They are gather-latency bound at low occupancy (ncu, profiles/r1_ncu_stage_kernels_r.csv: wfShadeKernel<0> 122 registers,
// 25 % of the warp slots filled, DRAM 21 % busy); 6 = 80 registers, 24 warps/SM: Rungholt-class 4K pass 8.36 -> 8.19 ms
// (shade 1.73 -> 1.66, resolve 0.61 -> 0.52), 8 = 64 registers gives the same (profiles/r1_trace_sweep.md).
// resident CTAs per SM asked of the non-Lambertian shade kernels (Principled, MetalWorkflow, Dielectric, ThinDielectric).  1 = the compiler's
// choice: 119-142 registers, 12-16 warps per SM.  5 (<= 102 registers, with the out-of-line BSDF functions held to 96 by -maxrregcount,
// 16-40 bytes of spills): shade stage 2.01 -> 1.85 ms Sponza-class, 2.53 -> 2.31 triple, 1.75 -> 1.70 Rungholt-class; 4 and 6 within 1 % of it
// (profiles/r2_sweep_shademinb_*.json)
#ifndef ZL_WF_SHADE_MINB_OTHER
#define ZL_WF_SHADE_MINB_OTHER 5
#endif
#ifndef ZL_WF_STAGE_MINB
#define ZL_WF_STAGE_MINB 6
#endif
static constexpr int kWfMaxDepth = 62;
static constexpr int kWfBins = 5;                      // material-type bins of the shade queues (materialBin)
static constexpr int kWfCntStride = 16;                // counters per bounce
static constexpr int kWfCounters = kWfCntStride * (kWfMaxDepth + 2);
// counter slots of bounce b at cnt[kWfCntStride * b + ...]
enum { kCntIn = 0 /* +bin */, kCntS = 5, kCntE = 6, kCntT = 7, kCntWork = 8, kCntOdd = 9 /* rays wfTraceRefillKernel left to the plain kernel */, kCntOddWork = 10 };

// The eight 16-byte fields every stage touches form ONE 128-byte record per path (slot): a stage that gathers a
// path's state by slot then pulls one cache line (consecutive sectors of one DRAM row) instead of up to eight
// sectors from eight arrays.  WfField keeps the `W.field[slot]` spelling.  ZL_WF_AOS=0 restores the SoA layout (A/B).
#ifndef ZL_WF_AOS
#define ZL_WF_AOS 1
#endif
static constexpr int kWfRecordVecs = ZL_WF_AOS ? 8 : 1;      // float4s between consecutive slots of one field
template <typename T>
struct WfField {
    T* p;
    __host__ __device__ __forceinline__ T& operator[](int slot) const { return p[(size_t)slot * kWfRecordVecs]; }
    __host__ __device__ __forceinline__ T* operator+(int slot) const { return p + (size_t)slot * kWfRecordVecs; }
};
struct WfState {
    WfField<float4> hit[2];   // {pos.xyz, bits(triangle id)}; shading point of bounce b in hit[b & 1], its successor in hit[(b+1) & 1]
    WfField<float4> dir;      // before shade(b): direction the path arrived with (wo = -dir); after: {wi.xyz, bsdfPdf}
    WfField<float4> thr;      // {throughput.xyz, bits(flags)}: bit 0 = delta BSDF sample, bit 1 = path ended at this bounce (bsdfPdf < 1e-8)
    WfField<float4> res;      // {result.xyz, Russian-roulette continue probability of this bounce}
    WfField<uint4>  smp;      // {randSeed, sampleSeed, dimension counter s, 0}
    WfField<float4> sh;       // deferred shadow ray {wi.xyz, max distance}; origin = rayOffseted(pos, wi)
    WfField<float4> shc;      // {NEE contribution.xyz, bits(1 = add it; trace clears it when occluded)}
    float4* sho;      // light-tracer modes only: explicit shadow-ray origin {origin.xyz, max distance}; then sh = {dir.xyz, uv.x}, shc = {contrib.xyz, uv.y}
    float4* aux;      // triple tracer: PT {t1s0, t1s1, coefToPrev, pdfDirToNext}; LPT {s0t1, s1t1, prevPdfDir, -}
    float4* nrm;      // triple tracer: {shading normal of the previous vertex (prevNorm).xyz, -}
    float*  tdist;    // distance returned by the closest-hit traversal of the last extension ray (bvhHit's `dist`)
    int* qIn[kWfBins];// paths to shade at the next shade stage, one queue per material-type bin
    int* qS; int* qE; int* qT;
    int* cnt;
    // ray sorting (wfSort*Kernel): queues S and E re-ordered by (MTBVH face, direction quadrant, Morton cell of the origin)
    int* qSs; int* qEs;   // sorted copies of qS / qE
    int* keyTmp;          // key of work item i (S items first, then E items)
    int* hist;            // 2 * sortBins: histogram, then running offsets, of the S and of the E keys; + scan block bases + ticket
    int tilesX, tilesY, nSlots;
    int capacity;         // slots the arrays and queues can hold (keyTmp holds 2 x capacity keys)
    int sortMode;         // experiment switch for wfSortKey (0 = default)
    int sortBits;         // bits per axis of the Morton cell in the sort key (kWfSortBitsDefault; ZL_WF_SORT_BITS)
    int sortBins;         // wfSortBins(sortBits): keys per queue
    int fusedKeys;        // 1: the shade kernels record sort keys + histogram themselves (no wfSortCountKernel)
    const int* passCounters;  // device {uSpp, uFreeCounter} of the pass when it is replayed as a CUDA graph (the by-value params are baked into the graph); nullptr otherwise
};

// the per-pass uniforms: by value normally, from device memory when the pass is a replayed graph (setPassCountersKernel is the graph's first node)
ZL_DEV ZlRenderParams wfPassParams(ZlRenderParams U, const WfState& W) {
    if (W.passCounters) { U.spp = W.passCounters[0]; U.freeCounter = W.passCounters[1]; }
    return U;
}
__global__ void setPassCountersKernel(int* counters, const int spp, const int freeCounter) { counters[0] = spp; counters[1] = freeCounter; }

ZL_DEV bool wfSlotPixel(const WfState& W, const ZlRenderParams& U, int slot, int& px, int& py) {
    const int tile = slot >> 5, lane = slot & 31;
    const int tx = tile % W.tilesX, ty = tile / W.tilesX;
    px = tx * 8 + (lane & 7);
    py = ty * 4 + (lane >> 3);
    return px < U.filmW && py < U.filmH;
}
// frame[coord] += result if !NaN (path_integ_naive.glsl:170-173); one owner per pixel per pass
ZL_DEV void wfFilmAdd(const WfState& W, const ZlRenderParams& U, float4* __restrict__ film, int slot, float3 result) {
    if (hasNan(result)) return;
    int px, py;
    wfSlotPixel(W, U, slot, px, py);
    float4* p = film + (size_t)py * U.filmW + px;
    float4 v = *p;
    v.x += result.x; v.y += result.y; v.z += result.z;
    *p = v;
}
// warp-aggregated queue append: one atomic per warp, ballot + popc for the lane offsets.
// Must be reached by all 32 lanes of the warp.
ZL_DEV void wfAppend(int* __restrict__ q, int* counter, bool pred, int slot) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) q[base + __popc(m & ((1u << lane) - 1u))] = slot;
}
// same, returning the queue position of the lane's item (-1 without one)
ZL_DEV int wfAppendAt(int* __restrict__ q, int* counter, bool pred, int slot) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return -1;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!pred) return -1;
    const int at = base + __popc(m & ((1u << lane) - 1u));
    q[at] = slot;
    return at;
}
// append to one of several queues chosen per lane (key < 0: none): lanes are grouped by key with
// match.any, one atomic per distinct key per warp.  Must be reached by all 32 lanes.
ZL_DEV void wfAppendKeyed(int* const* queues, int* counters, int key, int slot) {
    const unsigned part = __ballot_sync(0xffffffffu, key >= 0);
    if (key < 0) return;
    const unsigned peers = __match_any_sync(part, key);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counters + key, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    queues[key][base + __popc(peers & ((1u << lane) - 1u))] = slot;
}
// The ended-paths queue T is ONE array for the whole pass: every path ends exactly once, so bounce b appends at
// base(b) = the number of paths that ended in the bounces before it (final by the time any kernel of bounce b runs), and
// resolve(b) reads [base(b), base(b) + count(b)).  No buffer is reused within a pass, so resolve(b) may run at any later
// time (side stream; or, with pipelined passes, behind the previous pass's resolves).
ZL_DEV int wfEndedBase(const WfState& W, int b) {
    int s = 0;
    for (int j = 0; j < b; j++) s += W.cnt[kWfCntStride * j + kCntT];
    return s;
}
ZL_DEV int wfMaterialBinOfTriangle(const DScene& S, int id) {
    return materialBin(loadMaterialType(S, __ldg(&S.matTex[id]) & 0x0000ffff));
}

// Camera stage ("bounce 0"): seeds, camera sample, primary ray.  The ray is traced by wfTraceKernel like
// any extension ray (b = 0: no origin offset) and classified by wfResolveKernel / wfShadeKernel<TYPE>(1):
// with throughput 1 and the delta flag set, resolve's `radiance * throughput * weight` is exactly the
// envLe / lightLe the GLSL returns for a primary miss / emitter hit (path_integ_naive.glsl:38-43).
ZL_DEV void wfGenerateBody(const DScene& S, const ZlRenderParams& Uin, const WfState& W) {
    const ZlRenderParams U = wfPassParams(Uin, W);       // graph replays read the pass index from device memory
    __shared__ uint32_t row[256];
#if ZL_WF_AOS
    // the block's 128 records are assembled field-major in shared memory (row pitch 129: conflict-free both ways)
    // and written out as 16 KB of consecutive float4s; a direct per-thread store would put 16 bytes into each of
    // 32 different lines per instruction (measured 0.68 vs 0.35 ms at 4K).  hit[1] and sh are rewritten by
    // trace(0) / shade(1) before anything reads them.
    __shared__ float4 rec[8 * 129];
#endif
    stageSobolRow(S, U, row);
    __syncthreads();
    const int slot = blockIdx.x * 128 + threadIdx.x;
    int px = 0, py = 0;
    const bool valid = slot < W.nSlots && wfSlotPixel(W, U, slot, px, py);
    float4 vHit = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1)), vDir = make_float4(0.0f, 0.0f, 1.0f, 0.0f);
    uint4 vSmp = make_uint4(0u, 0u, 0u, 0u);
    if (valid) {
        float2 scrCoord = f2((float)px, (float)py) / f2((float)U.filmW, (float)U.filmH);
        SamplerState st = makeSampler(S, U, row, U.sampler);
        seedPixel(st, S, U, scrCoord);
        Ray ray = thinLensCameraSampleRay(U, scrCoord, sample4D(st));
        vHit = make_float4(ray.ori.x, ray.ori.y, ray.ori.z, __int_as_float(-1));
        vDir = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, 0.0f);
        vSmp = make_uint4(st.randSeed, st.sampleSeed, (uint32_t)st.s, 0u);
    }
    const float4 vThr = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(1)), vRes = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    const float4 vShc = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0));
#if ZL_WF_AOS
    const int t = threadIdx.x;
    // record order (wfEnsure): hit0, hit1, dir, sh, thr, res, smp, shc
    rec[0 * 129 + t] = vHit; rec[1 * 129 + t] = vHit; rec[2 * 129 + t] = vDir; rec[3 * 129 + t] = vShc;
    rec[4 * 129 + t] = vThr; rec[5 * 129 + t] = vRes;
    rec[6 * 129 + t] = make_float4(__uint_as_float(vSmp.x), __uint_as_float(vSmp.y), __uint_as_float(vSmp.z), 0.0f);
    rec[7 * 129 + t] = vShc;
    __syncthreads();
    float4* __restrict__ out = W.hit[0].p + (size_t)blockIdx.x * 128 * 8;
    const int live = min(128, W.nSlots - blockIdx.x * 128) * 8;           // float4s of this block that belong to existing slots
    for (int j = t; j < live; j += 128) out[j] = rec[(j & 7) * 129 + (j >> 3)];
#else
    if (valid) {
        W.hit[0][slot] = vHit; W.dir[slot] = vDir; W.thr[slot] = vThr; W.res[slot] = vRes; W.smp[slot] = vSmp; W.shc[slot] = vShc;
    }
#endif
    wfAppend(W.qE, W.cnt + kCntE, valid, slot);
}
__global__ void __launch_bounds__(128, ZL_WF_STAGE_MINB) wfGenerateKernel(const DScene S, const ZlRenderParams Uin, const WfState W) { wfGenerateBody(S, Uin, W); }
// "Dense" instantiations of the path tracer's three big stage kernels: 9 CTAs per SM (56 registers, some spills) instead of 6 (80).  With
// several passes in flight and 10 trace CTAs per SM resident, TWO of these CTAs fit into what the trace kernel leaves free instead of one:
// 4K Rungholt-class 1175 -> 1197 Msamples/s, 1080p Sponza-class 348 -> 353; the 720p default scene, the light and the triple tracer lose
// 1-2 % with them and keep the 80-register kernels (profiles/r2_trace_sweep.md).  Used by the pipelined path tracer on films >= 2^20 pixels.
#ifndef ZL_WF_STAGE_MINB_DENSE
#define ZL_WF_STAGE_MINB_DENSE 9
#endif
__global__ void __launch_bounds__(128, ZL_WF_STAGE_MINB_DENSE) wfGenerateDenseKernel(const DScene S, const ZlRenderParams Uin, const WfState W) { wfGenerateBody(S, Uin, W); }

static constexpr int kWfSortBitsDefault = 5, kWfSortBitsMax = 7;
__host__ __device__ constexpr int wfSortBins(int bits) { return 6 * 4 * (1 << (3 * bits)); }

ZL_DEV uint32_t wfSpread3(uint32_t v) {   // up to 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x30000ffu;
    v = (v | (v << 8)) & 0x300f00fu;
    v = (v | (v << 4)) & 0x30c30c3u;
    v = (v | (v << 2)) & 0x9249249u;
    return v;
}
// scale = 2^bits / extent of the scene (wfSortGrid)
ZL_DEV int wfSortKey(float3 lo, float3 scale, float3 pos, float3 d, int mode, int bits) {
    const int face = cubemapFace(-d);
    const int axis = face >> 1;
    const float m1 = axis == 0 ? d.y : d.x, m2 = axis == 2 ? d.y : d.z;
    const int quad = (m1 < 0.0f ? 1 : 0) | (m2 < 0.0f ? 2 : 0);
    const float3 c = (pos - lo) * scale;
    const float top = (float)((1 << bits) - 1);
    const uint32_t cx = (uint32_t)fminf(fmaxf(c.x, 0.0f), top), cy = (uint32_t)fminf(fmaxf(c.y, 0.0f), top), cz = (uint32_t)fminf(fmaxf(c.z, 0.0f), top);
    const uint32_t morton = wfSpread3(cx) | (wfSpread3(cy) << 1) | (wfSpread3(cz) << 2);
    const int cells = 1 << (3 * bits);
    if (mode == 3) return face * 4 * cells + (int)morton * 4 + quad;      // face, cell, quadrant
    if (mode == 4) return face * 4 * cells + (int)morton;                 // face, cell (no quadrant)
    return (face * 4 + quad) * cells + (int)morton;
}
// key + histogram entry of one queued ray (used by wfSortCountKernel, and by the shade kernels that fuse this step):
// keys of queue S live in keyTmp[0, capacity), keys of queue E in keyTmp[capacity, 2 capacity)
ZL_DEV void wfSortRecordKey(const WfState& W, bool shadowQueue, int at, int key) {
    W.keyTmp[(shadowQueue ? 0 : W.capacity) + at] = key;
    // neighbouring items often share a key (same cell, same face): one atomic per distinct key per converged group
    const int bin = (shadowQueue ? 0 : W.sortBins) + key;
    const unsigned peers = __match_any_sync(__activemask(), bin);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(W.hist + bin, __popc(peers));
}
ZL_DEV void wfSortGrid(const DScene& S, const WfState& W, float3& lo, float3& scale) {
    const float4 rlo = __ldg(S.nodes), rhi = __ldg(S.nodes + 1);          // root bounds (entry 0 of face 0 is the root)
    lo = f3(rlo);
    scale = f3((float)(1 << W.sortBits)) / gmax(f3(rhi) - f3(rlo), f3(1e-20f));
}

// queue appends of a shade kernel + (WfState::fusedKeys) the sort keys of the queued rays, re-read from the record
// the lane has just written (its own stores; L2 hits) — for the shade kernels that do not keep origin and direction
// in registers up to this point.  Must be reached by all 32 lanes.
ZL_DEV void wfAppendRays(const DScene& S, const WfState& W, int* cnt, int b, int slot, bool toS, bool toE, bool explicitShadowOrigin) {
    const int atS = wfAppendAt(W.qS, cnt + kCntS, toS, slot);
    const int atE = wfAppendAt(W.qE, cnt + kCntE, toE, slot);
    if (!W.fusedKeys || !(toS || toE)) return;
    float3 lo, scale;
    wfSortGrid(S, W, lo, scale);
    const float3 pos = f3(W.hit[b & 1][slot]);
    if (toS) wfSortRecordKey(W, true, atS, wfSortKey(lo, scale, explicitShadowOrigin ? f3(W.sho[slot]) : pos, f3(W.sh[slot]), W.sortMode, W.sortBits));
    if (toE) wfSortRecordKey(W, false, atE, wfSortKey(lo, scale, pos, f3(W.dir[slot]), W.sortMode, W.sortBits));
}

// One kernel per material-type bin: TYPE is a compile-time constant, so only that BSDF's code is reachable.
template <uint32_t TYPE>
ZL_DEV void wfShadeBody(const DScene& S, const ZlRenderParams& Uin, const WfState& W, float4* __restrict__ film, const int b) {
    const ZlRenderParams U = wfPassParams(Uin, W);       // graph replays read the pass index from device memory
    __shared__ uint32_t row[256];
    stageSobolRow(S, U, row);
    __syncthreads();
    int* const cnt = W.cnt + kWfCntStride * b;
    const int n = cnt[kCntIn + TYPE];
    const int* __restrict__ qin = W.qIn[TYPE];
    int* const qT = W.qT + wfEndedBase(W, b);
    const int stride = gridDim.x * blockDim.x;
    float3 sortLo = f3(0.0f), sortScale = f3(0.0f);
    if (W.fusedKeys) wfSortGrid(S, W, sortLo, sortScale);
    for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {
        const int i = i0 + (threadIdx.x & 31);
        const bool valid = i < n;
        const int slot = valid ? qin[i] : 0;
        bool toS = false, toE = false, toT = false;
        float3 keyPos = f3(0.0f), keyDirS = f3(0.0f), keyDirE = f3(0.0f);
        if (valid) {
            const float4 h = W.hit[b & 1][slot];
            const float3 pos = f3(h);
            keyPos = pos;
            const int id = __float_as_int(h.w);
            const float3 wo = -f3(W.dir[slot]);
            float3 throughput = f3(W.thr[slot]);
            const float4 r4 = W.res[slot];
            float3 result = f3(r4);
            const uint4 sm = W.smp[slot];
            SamplerState st = makeSampler(S, U, row, U.sampler);
            st.randSeed = sm.x; st.sampleSeed = sm.y; st.s = (int)sm.z;
            bool alive = true;
            if (b > 1) {
                const float4 c = W.shc[slot];                                  // NEE of bounce b-1, visibility now known
                if (__float_as_int(c.w) != 0) result += f3(c);
                if (U.russianRoulette) {                                       // path_integ_naive.glsl:127-133, after bounce b-1
                    const float continueProb = r4.w;
                    if (sample1D(st) >= continueProb) {
                        // ended by roulette: the film write is resolve(b)'s (all film writes of a pass happen in the resolve kernels):
                        // queue T with "ended, nothing more to add" — resolve then adds exactly `result`
                        alive = false; toT = true;
                        W.res[slot] = make_float4(result.x, result.y, result.z, 1.0f);
                        W.shc[slot] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0));
                        W.thr[slot] = make_float4(throughput.x, throughput.y, throughput.z, __int_as_float(2));
                    } else throughput /= continueProb;
                }
            }
            if (alive) {
                // loadShadingPoint (path_integ_naive.glsl:54-67) with the type known at compile time
                SurfaceInfo surf = triangleSurfaceInfo(S, id, pos);
                const int matTexId = __ldg(&S.matTex[id]);
                const int matId = matTexId & 0x0000ffff, texId = matTexId >> 16;
                if (TYPE != Dielectric && TYPE != ThinDielectric) {
                    if (dot(surf.ns, wo) < 0) { surf.ns = -surf.ns; surf.ng = -surf.ng; }
                }
                const BSDFParam mat = loadMaterial(S, TYPE, matId, texId, surf.uv);
                const float3 ns = surf.ns;
                float4 shOut = make_float4(0.0f, 0.0f, 0.0f, 0.0f), shcOut = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0));
                if (U.sampleLight) {
                    float ud = sample1D(st);
                    float4 us = sample4D(st);
                    DeferredVis vis; vis.pending = false; vis.dist = 0.0f; vis.ray = makeRay(pos, f3(0.0f));
                    LightLiSample samp = sampleLightAndEnv(S, U, pos, ud, us, vis);
                    if (samp.pdf > 0.0f) {
                        float4 bsdfAndPdf = materialBSDFAndPdfT<TYPE>(mat, wo, samp.wi, ns, Radiance);
                        float weight = biHeuristic(samp.pdf, bsdfAndPdf.w);
                        float3 contrib = f3(bsdfAndPdf) * throughput * satDot(ns, samp.wi) * samp.coef * weight;
                        shOut = make_float4(vis.ray.dir.x, vis.ray.dir.y, vis.ray.dir.z, vis.dist);
                        shcOut = make_float4(contrib.x, contrib.y, contrib.z, __int_as_float(1));
                        keyDirS = vis.ray.dir;
                        // A contribution of exactly zero (light direction below the shading hemisphere: satDot = 0, or a BSDF that
                        // evaluates to 0) adds +0 whether the shadow ray is blocked or not: keep the add, skip the ray.  The reference
                        // casts it (the visibility test sits inside lightSampleLi / envSampleLi, light.glsl:122-155, 207-219).
                        toS = !(contrib.x == 0.0f && contrib.y == 0.0f && contrib.z == 0.0f);
                    }
                }
                BSDFSample samp = materialSampleT<TYPE>(mat, ns, wo, Radiance, sample3D(st), st);
                keyDirE = samp.wi;
                const float bsdfPdf = samp.pdf;
                const bool deltaBsdf = (samp.flag == SpecRefl || samp.flag == SpecTrans);
                int flags = deltaBsdf ? 1 : 0;
                float rrProb = 1.0f;
                if (bsdfPdf < 1e-8f) { flags |= 2; toT = true; }
                else {
                    throughput *= samp.bsdf / bsdfPdf * (deltaBsdf ? 1.0f : absDot(ns, samp.wi));
                    rrProb = gmin(maxComponent(samp.bsdf / bsdfPdf), 0.95f);
                    toE = true;
                }
                W.dir[slot] = make_float4(samp.wi.x, samp.wi.y, samp.wi.z, bsdfPdf);
                W.thr[slot] = make_float4(throughput.x, throughput.y, throughput.z, __int_as_float(flags));
                W.res[slot] = make_float4(result.x, result.y, result.z, rrProb);
                W.smp[slot] = make_uint4(st.randSeed, st.sampleSeed, (uint32_t)st.s, 0u);
                if (toS) W.sh[slot] = shOut;
                W.shc[slot] = shcOut;
            }
        }
        const int atS = wfAppendAt(W.qS, cnt + kCntS, toS, slot);
        const int atE = wfAppendAt(W.qE, cnt + kCntE, toE, slot);
        wfAppend(qT, cnt + kCntT, toT, slot);
        if (W.fusedKeys) {      // the sort's key + histogram pass, here where origin and direction are still in registers
            if (toS) wfSortRecordKey(W, true, atS, wfSortKey(sortLo, sortScale, keyPos, keyDirS, W.sortMode, W.sortBits));
            if (toE) wfSortRecordKey(W, false, atE, wfSortKey(sortLo, sortScale, keyPos, keyDirE, W.sortMode, W.sortBits));
        }
    }
}
template <uint32_t TYPE>
__global__ void __launch_bounds__(128, (TYPE == 0u ? ZL_WF_STAGE_MINB : ZL_WF_SHADE_MINB_OTHER)) wfShadeKernel(const DScene S, const ZlRenderParams Uin, const WfState W, float4* __restrict__ film, const int b) {
    wfShadeBody<TYPE>(S, Uin, W, film, b);
}
__global__ void __launch_bounds__(128, ZL_WF_STAGE_MINB_DENSE) wfShadeDenseKernel(const DScene S, const ZlRenderParams Uin, const WfState W, float4* __restrict__ film, const int b) {
    wfShadeBody<0u>(S, Uin, W, film, b);
}

// Queue traversal.  Work items [0, |S|) are shadow rays (bvhTest), [|S|, |S| + |E|) extension rays
// (bvhHit); every lane walks its own ray through the stackless MTBVH exactly like traverseCore, but
// the warp is kept busy two ways:
//   * ray regeneration: a lane whose ray has finished takes the next work item from the queue
//     (warps claim chunks of kWfChunk items with one atomic and hand them out with ballot/popc),
//     instead of idling until the longest ray of its warp ends;
//   * postponed leaf tests: a lane that reaches a leaf whose box it hits parks the triangle id and
//     stops walking; the Moeller-Trumbore test runs for all parked lanes together once fewer than
//     2/3 of the live lanes can still walk.  Each lane's own sequence of box tests, triangle tests
//     and distance updates is unchanged, so hits are bit-identical to bvhHit / bvhTest.
static constexpr int kWfChunk = 32;
static constexpr int kWfRefill = 8;      // refill when at least this many lanes are idle

// (Rejected tuning switches, measured in profiles/r1_trace_sweep.md: prefetch.global.L2 of both successor
// records, ld.global.nc.L2::128B, register caps for 40/48 warps per SM, votes every 2/4 steps.)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 8) wfTraceKernel(const DScene S, const WfState W, const int b, const int lastBounce) {
    int* const cnt = W.cnt + kWfCntStride * b;
    const int nS = cnt[kCntS], nE = cnt[kCntE], total = nS + nE;
    int* const work = cnt + kCntWork;
    int* const qT = W.qT + wfEndedBase(W, b);
    const WfField<float4> cur = W.hit[b & 1];
    const WfField<float4> nxt = W.hit[(b + 1) & 1];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned ltMask = (1u << lane) - 1u;
    const int n = S.bvhSize;

    // lane state: `has` = holds a ray; `walk` = can take a box step (has a ray, not parked at a leaf, not at the end)
    bool has = false, walk = false, anyhit = false, occluded = false;
    int slot = 0, k = 0, closest = -1, pend = -1;
    float dist = 0.0f;
    RayPrep rp = prepareRay(makeRay(f3(0.0f), f3(0.0f, 0.0f, 1.0f)));
    const float4* __restrict__ nodes = S.nodes;
    int chunkNext = 0, chunkEnd = 0;   // warp-uniform: the claimed, not yet handed out work items
    bool lastChunk = (total == 0);

    while (true) {
        // ---- results of finished rays (has a ray, cannot walk, nothing parked) ----
        const bool fin = has && !walk && pend < 0;
        if (__ballot_sync(FULL, fin)) {
            int key = -1;
            if (fin) {
                if (anyhit) { if (occluded) reinterpret_cast<int*>(W.shc + slot)[3] = 0; }
                else {
                    const float3 np = rp.o + rp.d * dist;                      // rayPoint(ray, dist)
                    nxt[slot] = make_float4(np.x, np.y, np.z, __int_as_float(closest));
                    if (closest == -1 || closest - S.objPrimCount >= 0 || lastBounce) key = kWfBins;
                    else key = wfMaterialBinOfTriangle(S, closest);
                }
                has = false;
            }
            // bins 0..4 -> qIn[bin] of bounce b+1, key 5 -> qT of bounce b
            const unsigned part = __ballot_sync(FULL, key >= 0);
            if (key >= 0) {
                const unsigned peers = __match_any_sync(part, key);
                const int leader = __ffs(peers) - 1;
                int* counter = (key == kWfBins) ? (cnt + kCntT) : (cnt + kWfCntStride + kCntIn + key);
                int* q = (key == kWfBins) ? qT : W.qIn[key];
                int base = 0;
                if (lane == leader) base = atomicAdd(counter, __popc(peers));
                base = __shfl_sync(peers, base, leader);
                q[base + __popc(peers & ltMask)] = slot;
            }
        }
        // ---- ray regeneration ----
        const unsigned idleMask = __ballot_sync(FULL, !has);
        const bool haveWork = chunkNext < chunkEnd || !lastChunk;
        if (haveWork && (__popc(idleMask) >= kWfRefill || idleMask == FULL)) {
            if (chunkNext >= chunkEnd) {
                int base = 0;
                if (lane == 0) base = atomicAdd(work, kWfChunk);
                base = __shfl_sync(FULL, base, 0);
                chunkNext = min(base, total);
                chunkEnd = min(base + kWfChunk, total);
                if (chunkEnd >= total) lastChunk = true;
            }
            const int take = min(__popc(idleMask), chunkEnd - chunkNext);
            const int rank = __popc(idleMask & ltMask);
            if (!has && rank < take) {
                const int i = chunkNext + rank;
                anyhit = i < nS;
                slot = anyhit ? W.qS[i] : W.qE[i - nS];
                const float3 pos = f3(cur[slot]);
                const float4 d4 = anyhit ? W.sh[slot] : W.dir[slot];
                const Ray r = (b == 0) ? makeRay(pos, f3(d4)) : rayOffseted(pos, f3(d4));   // b = 0: camera rays start at the lens
                rp = prepareRay(r);
                nodes = S.nodes + (size_t)cubemapFace(-r.dir) * (size_t)n * 2;
                dist = anyhit ? d4.w : 1e8f;
                closest = -1; k = 0; pend = -1; occluded = false;
                has = true; walk = (n > 0);
            }
            chunkNext += take;
        }
        const unsigned walkMask = __ballot_sync(FULL, walk);
        if (walkMask == 0 && __ballot_sync(FULL, pend >= 0) == 0) {
            if (__ballot_sync(FULL, has) != 0) continue;                       // only finished rays left: write them out
            if (chunkNext >= chunkEnd && lastChunk) break;
            continue;
        }
        // ---- box phase: walk until fewer than 2/3 of the lanes that could walk at entry still can ----
        const int minWalk = (2 * __popc(walkMask) + 2) / 3;
        if (walkMask != 0) {
            while (true) {
                if (walk) {
                    float4 lo, hi;
                    loadNode(nodes, k, lo, hi);
                    float boxDist;
                    const bool bHit = rp.pure ? boxHitPure(f3(lo), f3(hi), rp, boxDist) : boxHit<true>(f3(lo), f3(hi), rp, boxDist);
                    const int prim = __float_as_int(lo.w);
                    if (!bHit || boxDist > dist) k = __float_as_int(hi.w);
                    else if (prim >= 0) { pend = prim; walk = false; }
                    else k++;
                    if (k == n) walk = false;
                }
                if (__popc(__ballot_sync(FULL, walk)) < minWalk) break;
            }
        }
        // ---- leaf phase: all parked lanes test their triangle ----
        if (pend >= 0) {
            const float4* __restrict__ tp = S.triPos + 3 * (size_t)pend;
            const float4 a = __ldg(tp), bb = __ldg(tp + 1), c = __ldg(tp + 2);
            float t;
            walk = true;
            if (intersectTriangle(f3(a), f3(bb), f3(c), rp.o, rp.d, t) && t < dist) {
                if (anyhit) { occluded = true; walk = false; }
                else { dist = t; closest = pend; }
            }
            pend = -1;
            k++;
            if (k == n) walk = false;
        }
    }
}

// Queue traversal, default kernel.  A warp claims 32 consecutive (sorted) work items with one atomic
// and every lane walks its ray with the plain per-lane loop of traverseCore.  There is deliberately
// NO warp-synchronous instruction inside the walk: with independent thread scheduling the diverged
// groups of a warp (lanes at a box test, lanes in a triangle test) issue independently, so while one
// group waits for a node record another group of the same warp runs.  Measured against the
// regenerating kernel above on the same sorted queues: 11.1 vs 12.8 ms per pass
// (profiles/r1_trace_sweep.md); several rays per lane per claim (bigger batches) are slower again.
//   MODE 0 (camera paths: path tracer, triple-PT): shadow ray = rayOffseted(shading point, wi) with eps `shadowEps`
//          (1e-4 for the NEE of light.glsl, 1e-5 for visible()); an occluded ray clears the "add it" flag of its
//          contribution; extension rays that end the path go to queue T.
//   MODE 1 (light paths: light tracer, triple-LPT): shadow rays are camera connections with explicit origins; an
//          unoccluded one splats its contribution (red.global.add.v4.f32); extension rays that leave the scene or
//          hit an emitter simply end (light_path_integ.glsl:80-83), there is no queue T.
//   ODD: the work items are the queue positions listed in W.keyTmp[0, cnt[kCntOdd]) — the axis-parallel / near-zero-component
//        rays that wfTraceRefillKernel leaves to this kernel's general loop.
// STAGED (ZL_WF_TRACE_LOOP=6): the top kTopDepth levels of the six orderings (DScene::top, buildStagedTopKernel) are brought into
// shared memory by ONE TMA bulk copy per CTA (cp.async.bulk + mbarrier, SASS UBLKCP) and every pure ray starts its walk there
// (traverseWarpStaged, zl_traverse.cuh).  Same visit sequence, same results.
// S.statePolicy: the kernel's own reads and writes of path-state records go through the streaming (evict-first) cache operators.
ZL_DEV float4 wfLoad(const float4* p, const int streaming) { return streaming ? __ldcs(p) : *p; }
ZL_DEV void wfStore(float4* p, const float4 v, const int streaming) { if (streaming) __stcs(p, v); else *p = v; }
template <int BLOCK, int MINB, int MODE, bool ODD = false, bool STAGED = false, int LEAN = 0, bool PIPE = false>
__global__ void __launch_bounds__(BLOCK, MINB) wfTraceSimpleKernel(const DScene S, const WfState W, const int b, const int lastBounce,
                                                                   const float shadowEps, float4* __restrict__ film, const int filmW, const int filmH) {
    extern __shared__ float4 topShared[];
    if (STAGED) {
        __shared__ alignas(8) unsigned long long mbar;
        const unsigned mbarAddr = (unsigned)__cvta_generic_to_shared(&mbar), dstAddr = (unsigned)__cvta_generic_to_shared(topShared);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbarAddr));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbarAddr), "r"((unsigned)kTopBytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dstAddr), "l"(S.top), "r"((unsigned)kTopBytes), "r"(mbarAddr) : "memory");
        }
        asm volatile("{\n\t.reg .pred p;\n\tZL_TOP_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@!p bra ZL_TOP_WAIT;\n\t}" ::"r"(mbarAddr) : "memory");
    }
    const int sp = LEAN ? 0 : S.statePolicy;
    int* const cnt = W.cnt + kWfCntStride * b;
    const int nS = cnt[kCntS], nE = cnt[kCntE], total = ODD ? cnt[kCntOdd] : nS + nE;
    int* const work = cnt + (ODD ? kCntOddWork : kCntWork);
    int* const qT = W.qT + wfEndedBase(W, b);
    const WfField<float4> cur = W.hit[b & 1];
    const WfField<float4> nxt = W.hit[(b + 1) & 1];
    const int lane = threadIdx.x & 31;
    // PIPE (ZL_WF_TRACE_PIPE=1): the head of a chunk — claim (an L2 atomic), queue entry, path state: three dependent round trips during which
    // the warp has no node load in flight — is taken off the walk's critical path.  While chunk t is walked, the claim of chunk t+3 and the
    // queue entries of chunk t+2 (cp.async into shared memory: no register waits for them) are in flight and the path state of chunk t+1
    // is being pulled into L2 (prefetch.global.L2).  Same chunks, same items, same results.
    __shared__ int pipeSlot[PIPE ? BLOCK / 32 : 1][2][32];
    int pBase = 0, pSlot = 0, pB1 = 0, pB2 = 0, pRaw = 0, pPar = 0;
    const int wib = threadIdx.x >> 5;
    if (PIPE && !ODD) {
        int r = 0;
        if (lane == 0) r = atomicAdd(work, 32);
        pBase = __shfl_sync(0xffffffffu, r, 0);
        if (pBase + lane < total) { const int i0 = pBase + lane; pSlot = i0 < nS ? W.qS[i0] : W.qE[i0 - nS]; }
        r = 0;
        if (lane == 0) r = pBase < total ? atomicAdd(work, 32) : pBase;
        pB1 = __shfl_sync(0xffffffffu, r, 0);
        if (pB1 + lane < total) {
            const int i1 = pB1 + lane;
            const int* src = i1 < nS ? W.qS + i1 : W.qE + (i1 - nS);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(&pipeSlot[wib][0][lane])), "l"(src) : "memory");
        }
        if (lane == 0) pRaw = pB1 < total ? atomicAdd(work, 32) : pB1;
    }
    while (true) {
        int base = 0;
        if (PIPE && !ODD) {
            base = pBase;
            if (base >= total) break;
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            if (pB1 + lane < total) {                    // chunk t+1: its queue entries arrived during the last walk; pull its path state into L2
#ifndef ZL_PIPE_PREFETCH
#define ZL_PIPE_PREFETCH 1
#endif
#if ZL_PIPE_PREFETCH
                const int s1 = pipeSlot[wib][pPar][lane];
                asm volatile("prefetch.global.L2 [%0];" ::"l"(cur + s1));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(W.dir + s1));
#endif
            }
            pB2 = __shfl_sync(0xffffffffu, pRaw, 0);     // chunk t+2: claimed during the last walk; its queue entries go to the other buffer
            if (pB2 + lane < total) {
                const int i2 = pB2 + lane;
                const int* src = i2 < nS ? W.qS + i2 : W.qE + (i2 - nS);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(&pipeSlot[wib][pPar ^ 1][lane])), "l"(src) : "memory");
            }
            if (lane == 0) pRaw = pB2 < total ? atomicAdd(work, 32) : pB2;      // chunk t+3
        } else {
            if (lane == 0) base = atomicAdd(work, 32);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= total) break;
        }
        const bool valid = base + lane < total;
        const int i = ODD ? (valid ? W.keyTmp[base + lane] : 0) : base + lane;
        const bool isShadow = i < nS;
        const int slot = (PIPE && !ODD) ? (valid ? pSlot : 0) : (valid ? (isShadow ? W.qS[i] : W.qE[i - nS]) : 0);
        int key = -1;
        if (valid) {
            const float3 pos = f3(wfLoad(cur + slot, sp));
            if (isShadow) {
                const float4 s4 = wfLoad(W.sh + slot, sp);
                if (MODE == 0) {
                    float d = s4.w;
                    const Ray sr = makeRay(pos + f3(s4) * shadowEps, f3(s4));
                    if (STAGED ? traverseWarpStaged<true>(S, topShared, sr, d) : traverseWarp<true, LEAN>(S, sr, d)) reinterpret_cast<int*>(W.shc + slot)[3] = 0;
                } else {
                    const float4 o4 = W.sho[slot];
                    float d = o4.w;
                    const Ray sr = makeRay(f3(o4), f3(s4));
                    if (!(STAGED ? traverseWarpStaged<true>(S, topShared, sr, d) : traverseWarp<true, LEAN>(S, sr, d))) {
                        const float4 c4 = W.shc[slot];                          // accumulateFilm (light_path_integ.glsl:34-43)
                        const int ix = (int)(s4.w * (float)filmW), iy = (int)(c4.w * (float)filmH);
                        if (ix >= 0 && iy >= 0 && ix < filmW && iy < filmH) {
                            float4* p = film + (size_t)iy * filmW + ix;
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(c4.x), "f"(c4.y), "f"(c4.z), "f"(0.0f) : "memory");
                        }
                    }
                }
            } else {
                const float3 dd = f3(wfLoad(W.dir + slot, sp));
                const Ray r = (b == 0) ? makeRay(pos, dd) : rayOffseted(pos, dd);   // b = 0: camera rays start at the lens, emission rays carry their offsets
                float dist;
                const int id = STAGED ? traverseWarpStaged<false>(S, topShared, r, dist) : traverseWarp<false, LEAN>(S, r, dist);
                const float3 np = rayPoint(r, dist);
                wfStore(nxt + slot, make_float4(np.x, np.y, np.z, __int_as_float(id)), sp);
                W.tdist[slot] = dist;
                if (id == -1 || id - S.objPrimCount >= 0 || lastBounce) key = (MODE == 0) ? kWfBins : -1;
                else key = wfMaterialBinOfTriangle(S, id);
            }
        }
        // bins 0..4 -> qIn[bin] of bounce b+1, key 5 -> qT of bounce b; one atomic per distinct key per warp
        const unsigned part = __ballot_sync(0xffffffffu, key >= 0);
        if (key >= 0) {
            const unsigned peers = __match_any_sync(part, key);
            const int leader = __ffs(peers) - 1;
            int* counter = (key == kWfBins) ? (cnt + kCntT) : (cnt + kWfCntStride + kCntIn + key);
            int* q = (key == kWfBins) ? qT : W.qIn[key];
            int off = 0;
            if (lane == leader) off = atomicAdd(counter, __popc(peers));
            off = __shfl_sync(peers, off, leader);
            q[off + __popc(peers & ((1u << lane) - 1u))] = slot;
        }
        if (PIPE && !ODD) {      // rotate: chunk t+1 becomes the current one
            pBase = pB1;
            pSlot = (pB1 + lane < total) ? pipeSlot[wib][pPar][lane] : 0;
            pB1 = pB2;
            pPar ^= 1;
        }
    }
}

// Queue traversal with intra-warp ray compaction (ZL_WF_TRACE_LOOP=7; traversePureCompact, zl_traverse.cuh).  Work distribution, results,
// queue appends and splats are wfTraceSimpleKernel's; a warp whose 32 items are pure rays of one kind and one octant (the sorted queues make
// that the rule) walks them with the compacting loop, any other warp with the plain one.
template <int BLOCK, int MINB, int MODE>
__global__ void __launch_bounds__(BLOCK, MINB) wfTraceCompactKernel(const DScene S, const WfState W, const int b, const int lastBounce,
                                                                    const float shadowEps, float4* __restrict__ film, const int filmW, const int filmH) {
    __shared__ int2 warpResAll[BLOCK / 32][32];
    int2* const warpRes = warpResAll[threadIdx.x >> 5];
    int* const cnt = W.cnt + kWfCntStride * b;
    const int nS = cnt[kCntS], nE = cnt[kCntE], total = nS + nE;
    int* const work = cnt + kCntWork;
    int* const qT = W.qT + wfEndedBase(W, b);
    const WfField<float4> cur = W.hit[b & 1];
    const WfField<float4> nxt = W.hit[(b + 1) & 1];
    const int lane = threadIdx.x & 31;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(work, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) break;
        const bool valid = base + lane < total;
        const int i = base + lane;
        const bool isShadow = i < nS;
        const int slot = valid ? (isShadow ? W.qS[i] : W.qE[i - nS]) : 0;
        Ray r = makeRay(f3(0.0f), f3(0.0f, 0.0f, 1.0f));
        float d = 1e8f;
        float4 s4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (valid) {
            const float3 pos = f3(cur[slot]);
            if (isShadow) {
                s4 = W.sh[slot];
                if (MODE == 0) { r = makeRay(pos + f3(s4) * shadowEps, f3(s4)); d = s4.w; }
                else { const float4 o4 = W.sho[slot]; r = makeRay(f3(o4), f3(s4)); d = o4.w; }
            } else {
                const float3 dd = f3(W.dir[slot]);
                r = (b == 0) ? makeRay(pos, dd) : rayOffseted(pos, dd);
            }
        }
        const bool pure = prepareRay(r).pure;
        const int cls = (valid && pure) ? (rayOctant(r.dir) | (isShadow ? 8 : 0)) : (16 + lane);
        int uniform = 0;
        __match_all_sync(0xffffffffu, cls, &uniform);
        int res = -1;
        if (uniform) {
            if (isShadow) res = traverseWarpCompact<true>(S, r, d, cls & 7, warpRes);
            else res = traverseWarpCompact<false>(S, r, d, cls & 7, warpRes);
        } else if (valid) {
            if (isShadow) res = traverseWarp<true>(S, r, d);
            else res = traverseWarp<false>(S, r, d);
        }
        __syncwarp();
        int key = -1;
        if (valid) {
            if (isShadow) {
                if (MODE == 0) {
                    if (res) reinterpret_cast<int*>(W.shc + slot)[3] = 0;
                } else if (!res) {
                    const float4 c4 = W.shc[slot];                          // accumulateFilm (light_path_integ.glsl:34-43)
                    const int ix = (int)(s4.w * (float)filmW), iy = (int)(c4.w * (float)filmH);
                    if (ix >= 0 && iy >= 0 && ix < filmW && iy < filmH) {
                        float4* p = film + (size_t)iy * filmW + ix;
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(c4.x), "f"(c4.y), "f"(c4.z), "f"(0.0f) : "memory");
                    }
                }
            } else {
                const int id = res;
                const float3 np = rayPoint(r, d);
                nxt[slot] = make_float4(np.x, np.y, np.z, __int_as_float(id));
                W.tdist[slot] = d;
                if (id == -1 || id - S.objPrimCount >= 0 || lastBounce) key = (MODE == 0) ? kWfBins : -1;
                else key = wfMaterialBinOfTriangle(S, id);
            }
        }
        const unsigned part = __ballot_sync(0xffffffffu, key >= 0);
        if (key >= 0) {
            const unsigned peers = __match_any_sync(part, key);
            const int leader = __ffs(peers) - 1;
            int* counter = (key == kWfBins) ? (cnt + kCntT) : (cnt + kWfCntStride + kCntIn + key);
            int* q = (key == kWfBins) ? qT : W.qIn[key];
            int off = 0;
            if (lane == leader) off = atomicAdd(counter, __popc(peers));
            off = __shfl_sync(peers, off, leader);
            q[off + __popc(peers & ((1u << lane) - 1u))] = slot;
        }
    }
}

// Queue traversal, two rays per lane (ZL_WF_TRACE_LOOP=5).  A warp claims 64 consecutive sorted items; lane l walks items
// base + l and base + 32 + l together (traverseDual: both node records requested before either is tested), so the warp has
// two node loads in flight per step instead of one.  Rays that are not "pure" (axis-parallel or a near-zero component) are
// walked first by the general loop; results, queue appends and splats are those of wfTraceSimpleKernel, item by item.
// (Everything per ray is a named scalar: an array of ray states indexed by an unrolled loop variable ended up in local memory.)
template <int MODE>
ZL_DEV void wfDualSetup(const DScene& S, const WfState& W, const WfField<float4>& cur, const int b, const float shadowEps, const int i, const int nS, const int total,
                        WalkRay& r, int& slot, bool& valid, bool& shadow, int& oct) {
    const int n = S.bvhSize;
    valid = i < total;
    shadow = i < nS;
    slot = valid ? (shadow ? W.qS[i] : W.qE[i - nS]) : 0;
    r.k = 0; r.end = 0; r.closest = -1; r.dist = 1e8f; r.anyhit = shadow;
    r.o = f3(0.0f); r.d = f3(0.0f); r.dInv = f3(0.0f);
    oct = 8;
    if (!valid) return;
    const float3 pos = f3(cur[slot]);
    Ray ray;
    if (shadow) {
        const float4 s4 = W.sh[slot];
        if (MODE == 0) { ray = makeRay(pos + f3(s4) * shadowEps, f3(s4)); r.dist = s4.w; }
        else { const float4 o4 = W.sho[slot]; ray = makeRay(f3(o4), f3(s4)); r.dist = o4.w; }
    } else {
        const float3 dd = f3(W.dir[slot]);
        ray = (b == 0) ? makeRay(pos, dd) : rayOffseted(pos, dd);
    }
    const RayPrep rp = prepareRay(ray);
    r.o = ray.ori; r.d = ray.dir; r.dInv = rp.dInv;
    if (rp.pure && n > 0) {
        const int off = cubemapFace(-ray.dir) * n;
        r.k = off; r.end = off + n;
        oct = rayOctant(ray.dir);
    } else {                // measure-zero set: boxHit's other branches, one ray at a time
        float dist = r.dist;
        const float4* __restrict__ nodes = S.nodes + (size_t)cubemapFace(-ray.dir) * (size_t)n * 2;
        if (shadow) r.closest = traversePrepared<true, false>(nodes, S.triPos, n, rp, dist, nullptr) ? 0 : -1;
        else { dist = 1e8f; r.closest = traversePrepared<false, false>(nodes, S.triPos, n, rp, dist, nullptr); r.dist = dist; }
    }
}
// result of one item + its queue append (all 32 lanes)
template <int MODE>
ZL_DEV void wfDualFinish(const DScene& S, const WfState& W, const WfField<float4>& nxt, int* const cnt, int* const qT, const int lastBounce,
                         float4* __restrict__ film, const int filmW, const int filmH, const WalkRay& r, const int slot, const bool valid, const bool shadow) {
    const int lane = threadIdx.x & 31;
    int key = -1;
    if (valid) {
        if (shadow) {
            const bool occluded = r.closest >= 0;
            if (MODE == 0) {
                if (occluded) reinterpret_cast<int*>(W.shc + slot)[3] = 0;
            } else if (!occluded) {
                const float4 s4 = W.sh[slot];
                const float4 c4 = W.shc[slot];                          // accumulateFilm (light_path_integ.glsl:34-43)
                const int ix = (int)(s4.w * (float)filmW), iy = (int)(c4.w * (float)filmH);
                if (ix >= 0 && iy >= 0 && ix < filmW && iy < filmH) {
                    float4* p = film + (size_t)iy * filmW + ix;
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(c4.x), "f"(c4.y), "f"(c4.z), "f"(0.0f) : "memory");
                }
            }
        } else {
            const int id = r.closest;
            const float3 np = rayPoint(makeRay(r.o, r.d), r.dist);
            nxt[slot] = make_float4(np.x, np.y, np.z, __int_as_float(id));
            W.tdist[slot] = r.dist;
            if (id == -1 || id - S.objPrimCount >= 0 || lastBounce) key = (MODE == 0) ? kWfBins : -1;
            else key = wfMaterialBinOfTriangle(S, id);
        }
    }
    __syncwarp();
    const unsigned part = __ballot_sync(0xffffffffu, key >= 0);
    if (key >= 0) {
        const unsigned peers = __match_any_sync(part, key);
        const int leader = __ffs(peers) - 1;
        int* counter = (key == kWfBins) ? (cnt + kCntT) : (cnt + kWfCntStride + kCntIn + key);
        int* q = (key == kWfBins) ? qT : W.qIn[key];
        int off = 0;
        if (lane == leader) off = atomicAdd(counter, __popc(peers));
        off = __shfl_sync(peers, off, leader);
        q[off + __popc(peers & ((1u << lane) - 1u))] = slot;
    }
    __syncwarp();
}
template <int BLOCK, int MINB, int MODE>
__global__ void __launch_bounds__(BLOCK, MINB) wfTraceDualKernel(const DScene S, const WfState W, const int b, const int lastBounce,
                                                                 const float shadowEps, float4* __restrict__ film, const int filmW, const int filmH) {
    int* const cnt = W.cnt + kWfCntStride * b;
    const int nS = cnt[kCntS], nE = cnt[kCntE], total = nS + nE;
    int* const work = cnt + kCntWork;
    int* const qT = W.qT + wfEndedBase(W, b);
    const WfField<float4> cur = W.hit[b & 1];
    const WfField<float4> nxt = W.hit[(b + 1) & 1];
    const int lane = threadIdx.x & 31;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(work, 64);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) break;
        WalkRay r0, r1;
        int slot0, slot1, oct0, oct1;
        bool valid0, valid1, shadow0, shadow1;
        wfDualSetup<MODE>(S, W, cur, b, shadowEps, base + lane, nS, total, r0, slot0, valid0, shadow0, oct0);
        wfDualSetup<MODE>(S, W, cur, b, shadowEps, base + 32 + lane, nS, total, r1, slot1, valid1, shadow1, oct1);
        __syncwarp();
        // one octant for every ray of the warp that takes part?  (sorted queues: nearly always)
        int uni = -1;
        {
            const unsigned m0 = __ballot_sync(0xffffffffu, oct0 < 8), m1 = __ballot_sync(0xffffffffu, oct1 < 8);
            if (S.octantWalk && (m0 | m1)) {
                const int ref = m0 ? __shfl_sync(0xffffffffu, oct0, __ffs(m0) - 1) : __shfl_sync(0xffffffffu, oct1, __ffs(m1) - 1);
                const bool same = (oct0 == 8 || oct0 == ref) && (oct1 == 8 || oct1 == ref);
                if (__all_sync(0xffffffffu, same)) uni = ref;
            }
        }
        traverseDualDispatch(S.nodes, S.triPos, S.bvhSize, r0, r1, uni);
        __syncwarp();
        wfDualFinish<MODE>(S, W, nxt, cnt, qT, lastBounce, film, filmW, filmH, r0, slot0, valid0, shadow0);
        wfDualFinish<MODE>(S, W, nxt, cnt, qT, lastBounce, film, filmW, filmH, r1, slot1, valid1, shadow1);
    }
}

// Queue traversal with DEFERRED LEAF TESTS (LOOP 2, 3) or with look-ahead node loads only (LOOP 1).
//
// Why: in the plain loop a leaf is reached in ~4 % of a lane's steps, so in ~3 of 4 warp-steps SOME lane
// branches into the triangle test (3 dependent LDG.128 + ~60 instructions) with one or two lanes active —
// about 30 % of the issued instructions of wfTraceSimpleKernel run at 1-2 of 32 lanes
// (profiles/r1_ncu_wfTraceSimpleKernel.csv: 12.8-14.4 active lanes per instruction on secondary bounces).
// Here a lane that reaches a leaf whose box passes the cull test parks {triangle, boxDist} in one of two
// register slots and KEEPS WALKING with its current (possibly stale) `dist`.  The warp runs the parked
// tests together when a lane has both slots full, when `flushAt` lanes hold a parked test, or when no
// lane can walk any more.
//
// Exactness.  The reference tests leaf L iff boxHit(L) and boxDist(L) <= dist at the time L is reached,
// where dist is the closest hit among the leaves tested before L.  Parked leaves are replayed in visit
// order and each re-checks `boxDist > dist` against the dist left by its predecessors, which is that same
// value, so exactly the reference's triangles are tested, in the reference's order, with the reference's
// strict `t < dist` update.  Nodes walked with a stale dist are a superset of the reference's walk; a leaf
// under a node the reference would have culled has boxDist(leaf) >= boxDist(node) > dist (the slab
// arithmetic is monotone in the box bounds and a child box lies inside its parent's), so the replay drops
// it.  Any-hit rays never update dist; the first parked triangle that hits ends the ray.
template <int BLOCK, int MINB, int MODE, int LOOP>
__global__ void __launch_bounds__(BLOCK, MINB) wfTraceDeferKernel(const DScene S, const WfState W, const int b, const int lastBounce,
                                                                  const float shadowEps, float4* __restrict__ film, const int filmW, const int filmH,
                                                                  const int flushAt) {
    constexpr bool SPEC = (LOOP & 1) != 0;
    int* const cnt = W.cnt + kWfCntStride * b;
    const int nS = cnt[kCntS], nE = cnt[kCntE], total = nS + nE;
    int* const work = cnt + kCntWork;
    int* const qT = W.qT + wfEndedBase(W, b);
    const WfField<float4> cur = W.hit[b & 1];
    const WfField<float4> nxt = W.hit[(b + 1) & 1];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int n = S.bvhSize;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(work, 32);
        base = __shfl_sync(FULL, base, 0);
        if (base >= total) break;
        const int i = base + lane;
        const bool valid = i < total;
        const bool isShadow = i < nS;
        const int slot = valid ? (isShadow ? W.qS[i] : W.qE[i - nS]) : 0;
        Ray r = makeRay(f3(0.0f), f3(0.0f, 0.0f, 1.0f));
        float dist = 1e8f;
        float4 s4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (valid) {
            const float3 pos = f3(cur[slot]);
            if (isShadow) {
                s4 = W.sh[slot];
                if (MODE == 0) { r = makeRay(pos + f3(s4) * shadowEps, f3(s4)); dist = s4.w; }
                else { const float4 o4 = W.sho[slot]; r = makeRay(f3(o4), f3(s4)); dist = o4.w; }
            } else {
                const float3 dd = f3(W.dir[slot]);
                r = (b == 0) ? makeRay(pos, dd) : rayOffseted(pos, dd);
            }
        }
        // ---- the walk: warp-synchronous, one box step per lane per iteration ----
        const RayPrep rp = prepareRay(r);
        const float4* __restrict__ nodes = S.nodes + (size_t)cubemapFace(-r.dir) * (size_t)n * 2;
        int k = (valid && n > 0) ? 0 : n;
        int closest = -1;
        bool occluded = false;
        int p0 = -1, p1 = -1;
        float d0 = 0.0f, d1 = 0.0f;
        float4 lo, hi, nlo, nhi;
        if (SPEC && k != n) loadNode(nodes, 0, lo, hi);
        while (true) {
            if (k != n) {
                if (SPEC) loadNode(nodes, k + 1, nlo, nhi);
                else loadNode(nodes, k, lo, hi);
                float boxDist;
                const bool bHit = rp.pure ? boxHitPure(f3(lo), f3(hi), rp, boxDist) : boxHit<true>(f3(lo), f3(hi), rp, boxDist);
                if (!bHit || boxDist > dist) {
                    k = __float_as_int(hi.w);
                    if (SPEC && k != n) loadNode(nodes, k, lo, hi);
                } else {
                    const int prim = __float_as_int(lo.w);
                    if (prim >= 0) {
                        if (LOOP >= 2) {
                            if (p0 < 0) { p0 = prim; d0 = boxDist; } else { p1 = prim; d1 = boxDist; }
                        } else {
                            const float4* __restrict__ tp = S.triPos + 3 * (size_t)prim;
                            const float4 a = __ldg(tp), bb = __ldg(tp + 1), c = __ldg(tp + 2);
                            float t;
                            if (intersectTriangle(f3(a), f3(bb), f3(c), rp.o, rp.d, t) && t < dist) {
                                if (isShadow) { occluded = true; k = n - 1; }
                                else { dist = t; closest = prim; }
                            }
                        }
                    }
                    k++;
                    if (SPEC) { lo = nlo; hi = nhi; }
                }
            }
            if (LOOP >= 2) {
                const unsigned walking = __ballot_sync(FULL, k != n);
                const unsigned full = __ballot_sync(FULL, p1 >= 0);
                const unsigned parked = __ballot_sync(FULL, p0 >= 0);
                if (full != 0u || walking == 0u || __popc(parked) >= flushAt) {
                    if (p0 >= 0) {
                        if (!(d0 > dist)) {
                            const float4* __restrict__ tp = S.triPos + 3 * (size_t)p0;
                            const float4 a = __ldg(tp), bb = __ldg(tp + 1), c = __ldg(tp + 2);
                            float t;
                            if (intersectTriangle(f3(a), f3(bb), f3(c), rp.o, rp.d, t) && t < dist) {
                                if (isShadow) { occluded = true; k = n; p1 = -1; }
                                else { dist = t; closest = p0; }
                            }
                        }
                        p0 = -1;
                        if (p1 >= 0) {
                            if (!(d1 > dist)) {
                                const float4* __restrict__ tp = S.triPos + 3 * (size_t)p1;
                                const float4 a = __ldg(tp), bb = __ldg(tp + 1), c = __ldg(tp + 2);
                                float t;
                                if (intersectTriangle(f3(a), f3(bb), f3(c), rp.o, rp.d, t) && t < dist) {
                                    if (isShadow) { occluded = true; k = n; }
                                    else { dist = t; closest = p1; }
                                }
                            }
                            p1 = -1;
                        }
                    }
                    if (walking == 0u) break;
                }
            } else {
                if (k == n) break;
            }
        }
        // ---- results (as in wfTraceSimpleKernel) ----
        int key = -1;
        if (valid) {
            if (isShadow) {
                if (MODE == 0) {
                    if (occluded) reinterpret_cast<int*>(W.shc + slot)[3] = 0;
                } else if (!occluded) {
                    const float4 c4 = W.shc[slot];                          // accumulateFilm (light_path_integ.glsl:34-43)
                    const int ix = (int)(s4.w * (float)filmW), iy = (int)(c4.w * (float)filmH);
                    if (ix >= 0 && iy >= 0 && ix < filmW && iy < filmH) {
                        float4* p = film + (size_t)iy * filmW + ix;
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(c4.x), "f"(c4.y), "f"(c4.z), "f"(0.0f) : "memory");
                    }
                }
            } else {
                const float3 np = rayPoint(r, dist);
                nxt[slot] = make_float4(np.x, np.y, np.z, __int_as_float(closest));
                W.tdist[slot] = dist;
                if (closest == -1 || closest - S.objPrimCount >= 0 || lastBounce) key = (MODE == 0) ? kWfBins : -1;
                else key = wfMaterialBinOfTriangle(S, closest);
            }
        }
        const unsigned part = __ballot_sync(FULL, key >= 0);
        if (key >= 0) {
            const unsigned peers = __match_any_sync(part, key);
            const int leader = __ffs(peers) - 1;
            int* counter = (key == kWfBins) ? (cnt + kCntT) : (cnt + kWfCntStride + kCntIn + key);
            int* q = (key == kWfBins) ? qT : W.qIn[key];
            int off = 0;
            if (lane == leader) off = atomicAdd(counter, __popc(peers));
            off = __shfl_sync(peers, off, leader);
            q[off + __popc(peers & ((1u << lane) - 1u))] = slot;
        }
    }
}

// Queue traversal with ROUND-BASED LANE REFILL (LOOP 4).
//
// Why: in wfTraceSimpleKernel a warp walks until the longest of its 32 rays ends; on secondary bounces that is
// twice the steps of the average ray (profiles/r1_trace_sweep.md: box-step lane utilisation 0.51-0.55) and the
// kernel is issue-bound, so half of the issued box tests are for lanes whose ray has already ended.  The first
// regenerating kernel (wfTraceKernel) fixed the lanes but paid for it with warp votes in every step, parked
// triangle tests and 62 registers (32 warps/SM).  This one keeps the plain per-lane step and votes once per
// ROUND: every lane that holds a ray takes up to `stepsPerRound` steps of the branch-free walk with no
// warp-synchronous instruction in between (diverged groups still overlap), then the warp retires the finished
// rays together (coalesced routing with match.any), hands new queue items to the idle lanes when at least
// `refillAt` are idle, and starts the next round.  Per-lane state between rounds is {o, d, 1/d, k, dist, closest,
// slot, face}; rays that are not "pure" (axis-parallel or a near-zero component: measure zero) are listed in W.keyTmp
// (free once the queues are sorted) and traced by wfTraceSimpleKernel<ODD> right after.  Each ray's own sequence of box tests, triangle tests and distance
// updates is that of traversePure, so the results are bit-identical.
template <int BLOCK, int MINB, int MODE>
__global__ void __launch_bounds__(BLOCK, MINB) wfTraceRefillKernel(const DScene S, const WfState W, const int b, const int lastBounce,
                                                                   const float shadowEps, float4* __restrict__ film, const int filmW, const int filmH,
                                                                   const int stepsPerRound, const int refillAt) {
    int* const cnt = W.cnt + kWfCntStride * b;
    const int nS = cnt[kCntS], nE = cnt[kCntE], total = nS + nE;
    int* const work = cnt + kCntWork;
    int* const qT = W.qT + wfEndedBase(W, b);
    const WfField<float4> cur = W.hit[b & 1];
    const WfField<float4> nxt = W.hit[(b + 1) & 1];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned ltMask = (1u << lane) - 1u;
    const int n = S.bvhSize;
    const float4* __restrict__ triPos = S.triPos;

    // lane state between rounds, kept small (the kernel must fit the 40-48 registers of 48-40 warps per SM):
    //   tag   : -1 = no ray, else slot | (shadow ray ? 1 << 30 : 0)           o, dInv : origin and 1/direction
    //   k     : face-relative threaded index (n = ended), off = face * n        dist, closest : bvhHit's running result
    // The direction itself is only needed by triangle tests and by the hit point: re-read from the path record there.
    constexpr int kShadowBit = 1 << 30;
    int tag = -1, k = n, closest = -1, off = 0;
    float dist = 0.0f;
    float3 o = f3(0.0f), dInv = f3(1.0f);
    int chunkNext = 0, chunkEnd = 0;           // warp-uniform: claimed, not yet handed out items
    bool exhausted = (total == 0 || n == 0);

    while (true) {
        // ---- retire the rays that ended in the last round ----
        const bool fin = tag >= 0 && k == n;
        if (__any_sync(FULL, fin)) {
            int key = -1;
            const int slot = tag & (kShadowBit - 1);
            if (fin) {
                if (tag & kShadowBit) {
                    if (MODE == 0) {
                        if (closest >= 0) reinterpret_cast<int*>(W.shc + slot)[3] = 0;
                    } else if (closest < 0) {
                        const float4 s4 = W.sh[slot], c4 = W.shc[slot];         // accumulateFilm (light_path_integ.glsl:34-43)
                        const int ix = (int)(s4.w * (float)filmW), iy = (int)(c4.w * (float)filmH);
                        if (ix >= 0 && iy >= 0 && ix < filmW && iy < filmH) {
                            float4* p = film + (size_t)iy * filmW + ix;
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(c4.x), "f"(c4.y), "f"(c4.z), "f"(0.0f) : "memory");
                        }
                    }
                } else {
                    const float3 d = f3(W.dir[slot]);
                    const float3 np = o + d * dist;                              // rayPoint(ray, dist)
                    nxt[slot] = make_float4(np.x, np.y, np.z, __int_as_float(closest));
                    W.tdist[slot] = dist;
                    if (closest == -1 || closest - S.objPrimCount >= 0 || lastBounce) key = (MODE == 0) ? kWfBins : -1;
                    else key = wfMaterialBinOfTriangle(S, closest);
                }
                tag = -1;
            }
            const unsigned part = __ballot_sync(FULL, key >= 0);
            if (key >= 0) {
                const unsigned peers = __match_any_sync(part, key);
                const int leader = __ffs(peers) - 1;
                int* counter = (key == kWfBins) ? (cnt + kCntT) : (cnt + kWfCntStride + kCntIn + key);
                int* q = (key == kWfBins) ? qT : W.qIn[key];
                int at = 0;
                if (lane == leader) at = atomicAdd(counter, __popc(peers));
                at = __shfl_sync(peers, at, leader);
                q[at + __popc(peers & ltMask)] = slot;
            }
        }
        // ---- hand queue items to the idle lanes ----
        unsigned idle = __ballot_sync(FULL, tag < 0);
        while (!exhausted && (__popc(idle) >= refillAt)) {
            if (chunkNext >= chunkEnd) {
                int c0 = 0;
                if (lane == 0) c0 = atomicAdd(work, kWfChunk);
                c0 = __shfl_sync(FULL, c0, 0);
                chunkNext = min(c0, total);
                chunkEnd = min(c0 + kWfChunk, total);
            }
            const int take = min(__popc(idle), chunkEnd - chunkNext);
            const int rank = __popc(idle & ltMask);
            if (tag < 0 && rank < take) {
                const int i = chunkNext + rank;
                const bool shadow = i < nS;
                const int slot = shadow ? W.qS[i] : W.qE[i - nS];
                const float3 pos = f3(cur[slot]);
                Ray r;
                if (shadow) {
                    const float4 s4 = W.sh[slot];
                    if (MODE == 0) { r = makeRay(pos + f3(s4) * shadowEps, f3(s4)); dist = s4.w; }
                    else { const float4 o4 = W.sho[slot]; r = makeRay(f3(o4), f3(s4)); dist = o4.w; }
                } else {
                    const float3 dd = f3(W.dir[slot]);
                    r = (b == 0) ? makeRay(pos, dd) : rayOffseted(pos, dd);       // b = 0: camera rays start at the lens
                    dist = 1e8f;
                }
                const RayPrep rp = prepareRay(r);
                o = r.ori; dInv = rp.dInv;
                off = cubemapFace(-r.dir) * n;
                closest = -1; k = 0;
                tag = slot | (shadow ? kShadowBit : 0);
                if (!rp.pure) {             // axis-parallel or near-zero component (measure zero): boxHit's other branches, left to wfTraceSimpleKernel<ODD>
                    W.keyTmp[atomicAdd(cnt + kCntOdd, 1)] = i;
                    tag = -1;
                }
            }
            chunkNext += take;
            if (chunkNext >= total) exhausted = true;
            idle = __ballot_sync(FULL, tag < 0);
        }
        if (idle == FULL) {
            if (exhausted) break;
            continue;
        }
        // ---- one round: up to stepsPerRound steps of traversePure's walk per lane, no warp-synchronous instruction inside ----
        if (tag >= 0 && k != n) {
            int budget = stepsPerRound;
            do {
                float4 lo, hi;
                loadNode(S.nodes, k + off, lo, hi);
                const float ax = (lo.x - o.x) * dInv.x, ay = (lo.y - o.y) * dInv.y, az = (lo.z - o.z) * dInv.z;
                const float bx = (hi.x - o.x) * dInv.x, by = (hi.y - o.y) * dInv.y, bz = (hi.z - o.z) * dInv.z;
                const float nx = fminf(ax, bx), ny = fminf(ay, by), nz = fminf(az, bz);
                const float fx = fmaxf(ax, bx), fy = fmaxf(ay, by), fz = fmaxf(az, bz);
                const float dx = fx - nx, dy = fy - ny, dz = fz - nz;
                const float tyz = fz - ny, tzx = fx - nz, txy = fy - nx;
                const float tMin = fmaxf(fmaxf(nx, ny), nz), tMax = fminf(fminf(fx, fy), fz);
                const bool hit = (dy + dz > tyz) & (dz + dx > tzx) & (dx + dy > txy) & (tMax >= 0.0f) & (tMax >= tMin) & !(tMin > dist);
                const int prim = __float_as_int(lo.w);
                k = hit ? k + 1 : __float_as_int(hi.w);
                if (hit & (prim >= 0)) {
                    const float4* __restrict__ tp = triPos + 3 * (size_t)prim;
                    const float4 ta = __ldg(tp), tb = __ldg(tp + 1), tc = __ldg(tp + 2);
                    const bool shadow = (tag & kShadowBit) != 0;
                    const float3 d = f3((shadow ? W.sh : W.dir)[tag & (kShadowBit - 1)]);
                    float t;
                    if (intersectTriangle(f3(ta), f3(tb), f3(tc), o, d, t) && t < dist) {
                        closest = prim;
                        if (shadow) k = n;          // any hit ends the walk through the loop condition
                        else dist = t;
                    }
                }
            } while (k != n && --budget != 0);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Ray sorting.  A warp-step of the trace kernel costs the latency of its SLOWEST lane, and with
// incoherent lanes nearly every step has one lane that misses L2 (profiles/r1_ncu_wfTraceKernel_v2.csv:
// ~1500 cycles per step).  Secondary rays are therefore re-ordered so that neighbouring lanes start
// in the same region and walk the same threaded ordering: counting sort on
//   key = (face * 4 + signs of the two minor direction components) << 15 | 15-bit Morton code of the origin
// (three small kernels: histogram, scan, scatter).  Only the processing order changes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wfSortCountKernel(const DScene S, const WfState W, const int b, const int explicitShadowOrigin) {
    const int* cnt = W.cnt + kWfCntStride * b;
    const int nS = cnt[kCntS], total = nS + cnt[kCntE];
    float3 lo, scale;
    wfSortGrid(S, W, lo, scale);
    const WfField<float4> cur = W.hit[b & 1];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const bool sh = i < nS;
        const int slot = sh ? W.qS[i] : W.qE[i - nS];
        const float3 pos = f3((sh && explicitShadowOrigin) ? W.sho[slot] : cur[slot]);
        const float3 d = f3(sh ? W.sh[slot] : W.dir[slot]);
        wfSortRecordKey(W, sh, sh ? i : i - nS, wfSortKey(lo, scale, pos, d, W.sortMode, W.sortBits));
    }
}
// Exclusive scan of the two histograms, in place.  Block j scans 8192 consecutive bins (8 per thread)
// and records its total; the last block to finish turns the per-block totals of each histogram
// into block bases.  Afterwards offset(key) = base[key / 8192] + hist[key].  Grid = 2 * sortBins / 8192 blocks.
static constexpr int kWfScanTile = 8192;
__host__ __device__ constexpr int wfScanBlocks(int bits) { return 2 * wfSortBins(bits) / kWfScanTile; }    // 192 at 5 bits
// exclusive scan of one value per thread over the 1024-thread block; returns the block total through `total`
ZL_DEV int wfBlockExclusiveScan(int v, int* warpSum, int& total) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int x = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += x; }
    __syncthreads();                                                          // warpSum may still be read from the previous call
    if (lane == 31) warpSum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = warpSum[lane], winc = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int x = __shfl_up_sync(0xffffffffu, winc, off); if (lane >= off) winc += x; }
        warpSum[lane] = winc - w;                                             // exclusive warp bases
        if (lane == 31) warpSum[32] = winc;
    }
    __syncthreads();
    total = warpSum[32];
    return warpSum[warp] + inc - v;
}
__global__ void __launch_bounds__(1024) wfSortScanKernel(const WfState W) {
    __shared__ int warpSum[33];
    __shared__ bool last;
    int* h = W.hist + (size_t)blockIdx.x * kWfScanTile;
    int* base = W.hist + 2 * (size_t)W.sortBins;                              // gridDim.x block bases, then the ticket
    const int t = threadIdx.x;
    int4 a = reinterpret_cast<int4*>(h)[2 * t], c = reinterpret_cast<int4*>(h)[2 * t + 1];
    const int v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    int ex[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { ex[j] = sum; sum += v[j]; }
    int total;
    const int off0 = wfBlockExclusiveScan(sum, warpSum, total);
    if (t == 0) base[blockIdx.x] = total;
    reinterpret_cast<int4*>(h)[2 * t] = make_int4(off0 + ex[0], off0 + ex[1], off0 + ex[2], off0 + ex[3]);
    reinterpret_cast<int4*>(h)[2 * t + 1] = make_int4(off0 + ex[4], off0 + ex[5], off0 + ex[6], off0 + ex[7]);
    __threadfence();
    if (t == 0) last = atomicAdd(base + gridDim.x, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last) {                                                               // block totals -> block bases, per histogram (S, then E)
        __threadfence();
        const int half = (int)gridDim.x / 2;
        for (int hgram = 0; hgram < 2; hgram++) {
            volatile int* bs = base + hgram * half;
            int run = 0;
            for (int c0 = 0; c0 < half; c0 += 1024) {
                const int x = (c0 + t < half) ? bs[c0 + t] : 0;
                int chunk;
                const int e = wfBlockExclusiveScan(x, warpSum, chunk);
                if (c0 + t < half) bs[c0 + t] = run + e;
                run += chunk;
            }
        }
        if (t == 0) base[gridDim.x] = 0;
    }
}
__global__ void __launch_bounds__(256) wfSortScatterKernel(const WfState W, const int b) {
    const int* cnt = W.cnt + kWfCntStride * b;
    const int nS = cnt[kCntS], total = nS + cnt[kCntE];
    const int lane = threadIdx.x & 31;
    const int stride = gridDim.x * blockDim.x;
    // Four items per thread and iteration: the chain of an item is queue load -> L2 atomic on its bin -> scattered store, all latency
    // (issue slots 9 % busy with one item in flight per thread, profiles/r2_pass_full_rungholt.csv); the four chains overlap.  A warp's
    // k-th items are 32 consecutive queue entries, as before, so the vote-aggregated atomics see the same neighbours.
    constexpr int U = 4;
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += U * stride) {
        int slot[U], bin[U], off[U];
        unsigned peers[U];
        bool ok[U], sh[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = i0 + u * stride;
            ok[u] = i < total;
            sh[u] = i < nS;
            slot[u] = ok[u] ? (sh[u] ? W.qS[i] : W.qE[i - nS]) : 0;
            bin[u] = ok[u] ? (sh[u] ? 0 : W.sortBins) + W.keyTmp[sh[u] ? i : W.capacity + (i - nS)] : -1;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            off[u] = 0;
            peers[u] = 0;
            const unsigned act = __ballot_sync(__activemask(), ok[u]);
            if (ok[u]) {
                peers[u] = __match_any_sync(act, bin[u]);
                if (lane == __ffs(peers[u]) - 1) off[u] = atomicAdd(W.hist + bin[u], __popc(peers[u]));
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (ok[u]) {
                const int o = __shfl_sync(peers[u], off[u], __ffs(peers[u]) - 1);
                const int dst = W.hist[2 * (size_t)W.sortBins + bin[u] / kWfScanTile] + o + __popc(peers[u] & ((1u << lane) - 1u));
                (sh[u] ? W.qSs : W.qEs)[dst] = slot[u];
            }
        }
    }
}

// paths that end at bounce b (path_integ_naive.glsl:102-125 + the final film write)
ZL_DEV void wfResolveBody(const DScene& S, const ZlRenderParams& Uin, const WfState& W, float4* __restrict__ film, const int b) {
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
            const float4 d4 = W.dir[slot];
            const float3 wi = f3(d4), throughput = f3(t4);
            const float bsdfPdf = d4.w;
            const bool deltaBsdf = (flags & 1) != 0;
            if (nextId == -1) {
                float3 radiance = envLe(S, U, wi);
                float weight = 1.0f;
                if (U.sampleLight && !deltaBsdf) {
                    float envPdf = envPdfLi(S, U, wi) * pdfSelectEnv(S, U);
                    weight = (envPdf <= 0.0f) ? 0.0f : biHeuristic(bsdfPdf, envPdf);
                }
                result += radiance * throughput * weight;
            } else if (lightId >= 0) {
                const float3 pos = f3(W.hit[b & 1][slot]), nextPos = f3(hn);
                float3 radiance = lightLe(S, lightId, nextPos, -wi);
                float weight = 1.0f;
                if (U.sampleLight && !deltaBsdf) {
                    float lightPdf = lightPdfLi(S, lightId, pos, nextPos) * pdfSelectLight(S, U, lightId);
                    weight = (lightPdf <= 0.0f) ? 0.0f : biHeuristic(bsdfPdf, lightPdf);
                }
                result += radiance * throughput * weight;
            }
        }
        wfFilmAdd(W, U, film, slot, result);
    }
}
__global__ void __launch_bounds__(128, ZL_WF_STAGE_MINB) wfResolveKernel(const DScene S, const ZlRenderParams Uin, const WfState W, float4* __restrict__ film, const int b) {
    wfResolveBody(S, Uin, W, film, b);
}
__global__ void __launch_bounds__(128, ZL_WF_STAGE_MINB_DENSE) wfResolveDenseKernel(const DScene S, const ZlRenderParams Uin, const WfState W, float4* __restrict__ film, const int b) {
    wfResolveBody(S, Uin, W, film, b);
}

}  // namespace zl
#include "zl_wavefront_light.cuh"
#include "zl_wavefront_triple.cuh"
#endif  // !ZL_INSTRUMENT
