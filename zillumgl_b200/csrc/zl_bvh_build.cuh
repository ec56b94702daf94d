// zl_bvh_build.cuh — BVH::build / quickBuild (src/accelerator/BVH.cpp:116-144, 217-296) on the device.
//
// The reference builds its tree top-down with an explicit stack: 16-bucket binned SAH on the centroid extent along
// the widest axis of the CENTROID box, one triangle per leaf, nodes in pre-order (left child at offset + 1, right child
// at offset + 2 * nLeft), two-primitive nodes ordered by centroid, and a partition that keeps the left part in order and
// fills the right part from the back (partition<16>, BVH.cpp:97-114).  None of this depends on the ORDER in which the
// floating-point work is done: boxes grow by min / max, bucket counts are integers, the 15 SAH costs are evaluated from
// the finished bucket boxes with the reference's expression.  So the same tree can be built level by level:
//
//   bin      (per primitive)  bucket of the primitive in its segment; bucket count / box by atomics (ordered-int keys)
//   sah      (per segment)    prefix / suffix boxes, the 15 costs, split bucket, node record, child segments
//   flag + scan + scatter     stable-left / reversed-right partition as one exclusive scan of the "goes left" flags over
//                             the whole primitive array and a scatter into the second buffer; child centroid boxes by atomics
//   small    (per segment)    children with one or two primitives become nodes right away (leaf / ordered pair)
//
// A segment = the primitive range [l, r] of one node of the current level.  ~log2(T) + a few levels; each level touches
// only primitives whose node is not finished.  Output: bounds[2T-1] (6 floats) and sizeIndices[2T-1], the arrays
// BVH::buildHitTable / threadMtbvhKernel consume.  Bit-identical to the host build except that a bound that is a zero may
// carry the other sign (min / max of +0 and -0 depends on the visiting order on the host; the keys order -0 < +0).
#pragma once
#include <climits>
#include <cub/device/device_scan.cuh>
#include "zl_math.cuh"

namespace zl {
namespace bvhb {

static constexpr int kLeafMask = (int)0x80000000u;
static constexpr int kBuckets = 16;

// floats <-> unsigned keys with the same order (atomicMin / atomicMax on the keys)
ZL_DEV unsigned keyOf(float f) { const unsigned b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
ZL_DEV float floatOf(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

struct Seg {
    int offset, l, r;           // node index (pre-order), primitive range
    unsigned cmin[3], cmax[3];  // centroid box (rec.nodeExtent), as keys
};
struct SegOut {                 // what `sah` decides for a segment
    int splitBucket, nLeft;
    int child[2];               // index of the child segment in the next level's list, or -1 (child has <= 2 primitives)
};
struct Bins { int count[kBuckets]; unsigned bmin[kBuckets][3], bmax[kBuckets][3]; };

ZL_DEV int maxExtent(float3 lo, float3 hi) {                            // AABB::maxExtent (AABB.cpp)
    const float3 v = hi - lo;
    if (v.x > v.y) return v.x > v.z ? 0 : 2;
    return v.y > v.z ? 1 : 2;
}
ZL_DEV float comp(float3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
ZL_DEV float surfaceArea(float3 lo, float3 hi) { const float3 v = hi - lo; return 2.0f * (v.x * v.y + v.y * v.z + v.z * v.x); }
// int b = 16 * (c - axisMin) / (axisMax - axisMin), clamped (BVH.cpp:106-107,256-257); NaN / out of range converts like
// x86 cvttss2si (INT_MIN) and clamps to bucket 0 — the host restatement (host/BVH.cpp bucketIndex) does the same
ZL_DEV int bucketIndex(float c, float axisMin, float axisMax) {
    const float f = 16.0f * (c - axisMin) / (axisMax - axisMin);
    const int b = (f == f && f < 2147483648.0f && f >= -2147483648.0f) ? (int)f : INT_MIN;
    return max(min(b, 15), 0);
}
ZL_DEV float3 centroidOf(float4 lo, float4 hi) { return (f3(lo) + f3(hi)) * 0.5f; }                  // AABB::centroid
ZL_DEV void segAxis(const Seg& s, int& dim, float& axisMin, float& axisMax) {
    const float3 lo = f3(floatOf(s.cmin[0]), floatOf(s.cmin[1]), floatOf(s.cmin[2]));
    const float3 hi = f3(floatOf(s.cmax[0]), floatOf(s.cmax[1]), floatOf(s.cmax[2]));
    dim = maxExtent(lo, hi);
    axisMin = comp(lo, dim); axisMax = comp(hi, dim);
}
ZL_DEV void segInitBox(Seg& s) {
    const unsigned kMin = keyOf(1e8f), kMax = keyOf(-1e8f);                                          // AABB() (AABB.h:12)
    for (int a = 0; a < 3; a++) { s.cmin[a] = kMin; s.cmax[a] = kMax; }
}

// primitive records: lo = {pMin.xyz, bits(triangle index)}, hi = {pMax.xyz, -}
__global__ void initKernel(const float4* __restrict__ triPos, const int T, float4* __restrict__ lo, float4* __restrict__ hi,
                           int* __restrict__ segOf, Seg* __restrict__ segs) {
    // (same-address global atomics serialise in L2: every per-primitive reduction below goes through shared memory first
    //  whenever the whole block works on one segment, which is the case exactly where the contention would be)
    __shared__ unsigned smn[3], smx[3];
    if (threadIdx.x < 3) { smn[threadIdx.x] = keyOf(1e8f); smx[threadIdx.x] = keyOf(-1e8f); }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T) {
        const float3 a = f3(__ldg(triPos + 3 * (size_t)i)), b = f3(__ldg(triPos + 3 * (size_t)i + 1)), c = f3(__ldg(triPos + 3 * (size_t)i + 2));
        const float3 mn = gmin(gmin(a, b), c), mx = gmax(gmax(a, b), c);
        lo[i] = make_float4(mn.x, mn.y, mn.z, __int_as_float(i));
        hi[i] = make_float4(mx.x, mx.y, mx.z, 0.0f);
        segOf[i] = 0;
        const float3 cen = (mn + mx) * 0.5f;
        atomicMin(&smn[0], keyOf(cen.x)); atomicMin(&smn[1], keyOf(cen.y)); atomicMin(&smn[2], keyOf(cen.z));
        atomicMax(&smx[0], keyOf(cen.x)); atomicMax(&smx[1], keyOf(cen.y)); atomicMax(&smx[2], keyOf(cen.z));
    }
    __syncthreads();
    if (threadIdx.x < 3) { atomicMin(&segs[0].cmin[threadIdx.x], smn[threadIdx.x]); atomicMax(&segs[0].cmax[threadIdx.x], smx[threadIdx.x]); }
}
__global__ void rootKernel(Seg* segs, const int T) {
    Seg s; s.offset = 0; s.l = 0; s.r = T - 1; segInitBox(s);
    segs[0] = s;
}
__global__ void binInitKernel(Bins* __restrict__ bins, const int numSegs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numSegs * kBuckets) return;
    Bins& b = bins[i / kBuckets];
    const int k = i % kBuckets;
    b.count[k] = 0;
    const unsigned kMin = keyOf(1e8f), kMax = keyOf(-1e8f);
    for (int a = 0; a < 3; a++) { b.bmin[k][a] = kMin; b.bmax[k][a] = kMax; }
}
__global__ void binKernel(const float4* __restrict__ lo, const float4* __restrict__ hi, const int* __restrict__ segOf, const int T,
                          const Seg* __restrict__ segs, Bins* __restrict__ bins, unsigned char* __restrict__ bucketOf) {
    __shared__ Bins sb;
    __shared__ int uniformSeg;
    const int first = blockIdx.x * blockDim.x, last = min(first + (int)blockDim.x, T) - 1;
    if (threadIdx.x == 0) { const int a = segOf[first]; uniformSeg = (a >= 0 && a == segOf[last]) ? a : -1; }   // a segment is a contiguous range
    if (threadIdx.x < kBuckets) {
        sb.count[threadIdx.x] = 0;
        for (int a = 0; a < 3; a++) { sb.bmin[threadIdx.x][a] = keyOf(1e8f); sb.bmax[threadIdx.x][a] = keyOf(-1e8f); }
    }
    __syncthreads();
    const int i = first + threadIdx.x;
    const int s = i < T ? segOf[i] : -1;
    if (s >= 0) {
        int dim; float axisMin, axisMax;
        segAxis(segs[s], dim, axisMin, axisMax);
        const float4 l4 = lo[i], h4 = hi[i];
        const int b = bucketIndex(comp(centroidOf(l4, h4), dim), axisMin, axisMax);
        bucketOf[i] = (unsigned char)b;
        Bins& B = (uniformSeg >= 0) ? sb : bins[s];
        atomicAdd(&B.count[b], 1);
        atomicMin(&B.bmin[b][0], keyOf(l4.x)); atomicMin(&B.bmin[b][1], keyOf(l4.y)); atomicMin(&B.bmin[b][2], keyOf(l4.z));
        atomicMax(&B.bmax[b][0], keyOf(h4.x)); atomicMax(&B.bmax[b][1], keyOf(h4.y)); atomicMax(&B.bmax[b][2], keyOf(h4.z));
    }
    __syncthreads();
    if (uniformSeg >= 0 && threadIdx.x < kBuckets && sb.count[threadIdx.x] > 0) {
        Bins& B = bins[uniformSeg];
        const int b = threadIdx.x;
        atomicAdd(&B.count[b], sb.count[b]);
        for (int a = 0; a < 3; a++) { atomicMin(&B.bmin[b][a], sb.bmin[b][a]); atomicMax(&B.bmax[b][a], sb.bmax[b][a]); }
    }
}
// one thread per segment (>= 3 primitives): BVH.cpp:262-296
__global__ void sahKernel(const Seg* __restrict__ segs, const int numSegs, const Bins* __restrict__ bins, SegOut* __restrict__ out,
                          Seg* __restrict__ nextSegs, int* __restrict__ nextCount, float* __restrict__ bounds, int* __restrict__ sizeIndices) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= numSegs) return;
    const Seg seg = segs[s];
    const Bins& B = bins[s];
    const int nBoxes = seg.r - seg.l + 1;
    int count[kBuckets];
    float3 bmin[kBuckets], bmax[kBuckets];
    for (int i = 0; i < kBuckets; i++) {
        count[i] = B.count[i];
        bmin[i] = f3(floatOf(B.bmin[i][0]), floatOf(B.bmin[i][1]), floatOf(B.bmin[i][2]));
        bmax[i] = f3(floatOf(B.bmax[i][0]), floatOf(B.bmax[i][1]), floatOf(B.bmax[i][2]));
    }
    // suffix boxes / counts first (kept), then the prefix side on the fly
    int sufCount[kBuckets];
    float3 sufMin[kBuckets], sufMax[kBuckets];
    sufCount[15] = count[15]; sufMin[15] = bmin[15]; sufMax[15] = bmax[15];
    for (int i = 14; i >= 0; i--) { sufCount[i] = sufCount[i + 1] + count[i]; sufMin[i] = gmin(sufMin[i + 1], bmin[i]); sufMax[i] = gmax(sufMax[i + 1], bmax[i]); }
    int preCount = count[0];
    float3 preMin = bmin[0], preMax = bmax[0];
    int splitBucket = 0, nLeft = preCount;
    float minCost = preCount * surfaceArea(preMin, preMax) + sufCount[1] * surfaceArea(sufMin[1], sufMax[1]);
    for (int i = 1; i < 15; i++) {
        preCount += count[i]; preMin = gmin(preMin, bmin[i]); preMax = gmax(preMax, bmax[i]);
        const float cost = preCount * surfaceArea(preMin, preMax) + sufCount[i + 1] * surfaceArea(sufMin[i + 1], sufMax[i + 1]);
        if (cost < minCost) { minCost = cost; splitBucket = i; nLeft = preCount; }
    }
    // node record: bounds[offset] = preBox[15] = union of all buckets (= sufBox[0])
    float* nb = bounds + 6 * (size_t)seg.offset;
    nb[0] = sufMin[0].x; nb[1] = sufMin[0].y; nb[2] = sufMin[0].z; nb[3] = sufMax[0].x; nb[4] = sufMax[0].y; nb[5] = sufMax[0].z;
    sizeIndices[seg.offset] = nBoxes * 2 - 1;
    if (nLeft == nBoxes) nLeft = nBoxes - 1;                                                         // `if (pr == nBoxes) pr--` (BVH.cpp:112)
    SegOut o;
    o.splitBucket = splitBucket; o.nLeft = nLeft;
    const int sizes[2] = {nLeft, nBoxes - nLeft};
    const int offs[2] = {seg.offset + 1, seg.offset + 2 * nLeft};
    const int ls[2] = {seg.l, seg.l + nLeft};
    for (int c = 0; c < 2; c++) {
        o.child[c] = -1;
        if (sizes[c] >= 3) {
            const int idx = atomicAdd(nextCount, 1);
            Seg ns; ns.offset = offs[c]; ns.l = ls[c]; ns.r = ls[c] + sizes[c] - 1; segInitBox(ns);
            nextSegs[idx] = ns;
            o.child[c] = idx;
        }
    }
    out[s] = o;
}
__global__ void flagKernel(const int* __restrict__ segOf, const unsigned char* __restrict__ bucketOf, const SegOut* __restrict__ out, const int T,
                           int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const int s = segOf[i];
    flag[i] = (s >= 0 && (int)bucketOf[i] <= out[s].splitBucket) ? 1 : 0;
}
// stable on the left, reversed on the right (partition<16>, BVH.cpp:97-114)
__global__ void scatterKernel(const float4* __restrict__ lo, const float4* __restrict__ hi, const int* __restrict__ segOf, const int T,
                              const Seg* __restrict__ segs, const SegOut* __restrict__ out, const int* __restrict__ flag, const int* __restrict__ scan,
                              float4* __restrict__ lo2, float4* __restrict__ hi2, int* __restrict__ segOf2, Seg* __restrict__ nextSegs) {
    __shared__ unsigned smn[2][3], smx[2][3];
    __shared__ int uniformSeg;
    const int first = blockIdx.x * blockDim.x, last = min(first + (int)blockDim.x, T) - 1;
    if (threadIdx.x == 0) { const int a = segOf[first]; uniformSeg = (a >= 0 && a == segOf[last]) ? a : -1; }
    if (threadIdx.x < 6) { smn[threadIdx.x / 3][threadIdx.x % 3] = keyOf(1e8f); smx[threadIdx.x / 3][threadIdx.x % 3] = keyOf(-1e8f); }
    __syncthreads();
    const int i = first + threadIdx.x;
    const int s = i < T ? segOf[i] : -1;
    if (s >= 0) {
        const int l = segs[s].l, r = segs[s].r;
        const int leftsBefore = scan[i] - scan[l];
        const int pos = flag[i] ? l + leftsBefore : r - ((i - l) - leftsBefore);
        const float4 l4 = lo[i], h4 = hi[i];
        lo2[pos] = l4; hi2[pos] = h4;
        const SegOut o = out[s];
        const int side = pos < l + o.nLeft ? 0 : 1;
        const int child = o.child[side];
        segOf2[pos] = child;
        if (child >= 0) {
            const float3 cen = centroidOf(l4, h4);
            unsigned* mn = (uniformSeg >= 0) ? smn[side] : nextSegs[child].cmin;
            unsigned* mx = (uniformSeg >= 0) ? smx[side] : nextSegs[child].cmax;
            atomicMin(&mn[0], keyOf(cen.x)); atomicMin(&mn[1], keyOf(cen.y)); atomicMin(&mn[2], keyOf(cen.z));
            atomicMax(&mx[0], keyOf(cen.x)); atomicMax(&mx[1], keyOf(cen.y)); atomicMax(&mx[2], keyOf(cen.z));
        }
    }
    __syncthreads();
    if (uniformSeg >= 0 && threadIdx.x < 6) {
        const int side = threadIdx.x / 3, a = threadIdx.x % 3;
        const int child = out[uniformSeg].child[side];
        if (child >= 0) { atomicMin(&nextSegs[child].cmin[a], smn[side][a]); atomicMax(&nextSegs[child].cmax[a], smx[side][a]); }
    }
}
// a node with one (leaf) or two primitives (BVH.cpp:231-247), primitives read from the partitioned buffer
ZL_DEV void emitSmall(const float4* __restrict__ lo, const float4* __restrict__ hi, int offset, int l, int size, float* __restrict__ bounds, int* __restrict__ sizeIndices) {
    auto leaf = [&](int off, float4 l4, float4 h4) {
        float* nb = bounds + 6 * (size_t)off;
        nb[0] = l4.x; nb[1] = l4.y; nb[2] = l4.z; nb[3] = h4.x; nb[4] = h4.y; nb[5] = h4.z;
        sizeIndices[off] = __float_as_int(l4.w) | kLeafMask;
    };
    if (size == 1) { leaf(offset, lo[l], hi[l]); return; }
    float4 l0 = lo[l], h0 = hi[l], l1 = lo[l + 1], h1 = hi[l + 1];
    float* nb = bounds + 6 * (size_t)offset;
    const float3 mn = gmin(f3(l0), f3(l1)), mx = gmax(f3(h0), f3(h1));
    nb[0] = mn.x; nb[1] = mn.y; nb[2] = mn.z; nb[3] = mx.x; nb[4] = mx.y; nb[5] = mx.z;
    sizeIndices[offset] = 3;
    const float3 c0 = centroidOf(l0, h0), c1 = centroidOf(l1, h1);
    // rec.splitDim of this node = maxExtent of the centroid box of its two primitives (AABB() grown by both)
    const float3 cmn = gmin(gmin(f3(1e8f), c0), c1), cmx = gmax(gmax(f3(-1e8f), c0), c1);
    const int dim = maxExtent(cmn, cmx);
    if (comp(c0, dim) > comp(c1, dim)) { leaf(offset + 1, l1, h1); leaf(offset + 2, l0, h0); }
    else { leaf(offset + 1, l0, h0); leaf(offset + 2, l1, h1); }
}
__global__ void smallKernel(const float4* __restrict__ lo2, const float4* __restrict__ hi2, const Seg* __restrict__ segs, const SegOut* __restrict__ out,
                            const int numSegs, float* __restrict__ bounds, int* __restrict__ sizeIndices) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= numSegs) return;
    const Seg seg = segs[s];
    const SegOut o = out[s];
    const int nBoxes = seg.r - seg.l + 1;
    if (o.nLeft <= 2) emitSmall(lo2, hi2, seg.offset + 1, seg.l, o.nLeft, bounds, sizeIndices);
    if (nBoxes - o.nLeft <= 2) emitSmall(lo2, hi2, seg.offset + 2 * o.nLeft, seg.l + o.nLeft, nBoxes - o.nLeft, bounds, sizeIndices);
}
__global__ void rootSmallKernel(const float4* __restrict__ lo, const float4* __restrict__ hi, const int T, float* __restrict__ bounds, int* __restrict__ sizeIndices) {
    emitSmall(lo, hi, 0, 0, T, bounds, sizeIndices);
}

struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) cudaFree(p); }
    template <typename T> cudaError_t get(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = (T*)p;
        return e;
    }
};

}  // namespace bvhb

// triPos: device, 3 float4 per triangle (xyz = vertex).  bounds (6 * (2T-1) floats) and sizeIndices (2T-1 ints): device, written here.
// levelsOut (optional): number of levels run.  Synchronous (reads the segment count back once per level).
static cudaError_t buildBvhOnDevice(const float4* triPos, int T, float* bounds, int* sizeIndices, int* levelsOut = nullptr) {
    using namespace bvhb;
#define ZLB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)
    Scratch sc;
    float4 *lo[2], *hi[2];
    int* segOf[2];
    Seg* segs[2];
    SegOut* out; Bins* bins; unsigned char* bucketOf; int *flag, *scan, *nextCount;
    const size_t maxSegs = (size_t)T / 3 + 2;
    for (int k = 0; k < 2; k++) { ZLB(sc.get(&lo[k], T)); ZLB(sc.get(&hi[k], T)); ZLB(sc.get(&segOf[k], T)); ZLB(sc.get(&segs[k], maxSegs)); }
    ZLB(sc.get(&out, maxSegs)); ZLB(sc.get(&bins, maxSegs)); ZLB(sc.get(&bucketOf, T)); ZLB(sc.get(&flag, T)); ZLB(sc.get(&scan, T)); ZLB(sc.get(&nextCount, 1));
    size_t tempBytes = 0;
    ZLB(cub::DeviceScan::ExclusiveSum(nullptr, tempBytes, flag, scan, T));
    char* temp; ZLB(sc.get(&temp, tempBytes));
    const int B = 256, gridT = (T + B - 1) / B;
    rootKernel<<<1, 1>>>(segs[0], T);
    initKernel<<<gridT, B>>>(triPos, T, lo[0], hi[0], segOf[0], segs[0]);
    int levels = 0;
    if (T <= 2) {
        rootSmallKernel<<<1, 1>>>(lo[0], hi[0], T, bounds, sizeIndices);
    } else {
        int numSegs = 1, cur = 0;
        while (numSegs > 0) {
            const int gridS = (numSegs + B - 1) / B;
            ZLB(cudaMemsetAsync(nextCount, 0, sizeof(int)));
            binInitKernel<<<(numSegs * kBuckets + B - 1) / B, B>>>(bins, numSegs);
            binKernel<<<gridT, B>>>(lo[cur], hi[cur], segOf[cur], T, segs[cur], bins, bucketOf);
            sahKernel<<<gridS, B>>>(segs[cur], numSegs, bins, out, segs[cur ^ 1], nextCount, bounds, sizeIndices);
            flagKernel<<<gridT, B>>>(segOf[cur], bucketOf, out, T, flag);
            ZLB(cub::DeviceScan::ExclusiveSum(temp, tempBytes, flag, scan, T));
            ZLB(cudaMemsetAsync(segOf[cur ^ 1], 0xff, (size_t)T * sizeof(int)));                     // -1: finished
            scatterKernel<<<gridT, B>>>(lo[cur], hi[cur], segOf[cur], T, segs[cur], out, flag, scan, lo[cur ^ 1], hi[cur ^ 1], segOf[cur ^ 1], segs[cur ^ 1]);
            smallKernel<<<gridS, B>>>(lo[cur ^ 1], hi[cur ^ 1], segs[cur], out, numSegs, bounds, sizeIndices);
            ZLB(cudaGetLastError());
            ZLB(cudaMemcpy(&numSegs, nextCount, sizeof(int), cudaMemcpyDeviceToHost));
            cur ^= 1;
            levels++;
            if (levels > 4 * 1024 * 1024) return cudaErrorUnknown;                                   // cannot happen: every level splits every segment
        }
    }
    ZLB(cudaGetLastError());
    ZLB(cudaDeviceSynchronize());
    if (levelsOut) *levelsOut = levels;
#undef ZLB
    return cudaSuccess;
}


// ---- BVH2 records (zl_traverse.cuh traverseBvh2): one 64-byte record per interior node of the builder's pre-order tree ----
namespace bvh2b {
__global__ void leafFlagKernel(const int* __restrict__ sizeIndices, const int n, int* __restrict__ flag) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) flag[p] = (sizeIndices[p] & bvhb::kLeafMask) ? 1 : 0;
}
// depth of every node by descending from the root (like threadMtbvhKernel); max over nodes into *maxDepth
__global__ void depthKernel(const int* __restrict__ sizeIndices, const int n, int* __restrict__ maxDepth) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int cur = 0, depth = 0;
    while (cur != p) {
        const int l = cur + 1;
        const int sl = sizeIndices[l];
        const int sizeL = (sl & bvhb::kLeafMask) ? 1 : sl;
        cur = (p < l + sizeL) ? l : l + sizeL;
        depth++;
    }
    if (sizeIndices[p] & bvhb::kLeafMask) atomicMax(maxDepth, depth);
}
__global__ void recordKernel(const float* __restrict__ bounds, const int* __restrict__ sizeIndices, const int* __restrict__ leavesBefore, const int n,
                             float4* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int sp = sizeIndices[p];
    if (sp & bvhb::kLeafMask) return;
    const int L = p + 1;
    const int sl = sizeIndices[L];
    const int R = L + ((sl & bvhb::kLeafMask) ? 1 : sl);
    const int sr = sizeIndices[R];
    const int refL = (sl & bvhb::kLeafMask) ? ~(sl & 0x7fffffff) : L - leavesBefore[L];
    const int refR = (sr & bvhb::kLeafMask) ? ~(sr & 0x7fffffff) : R - leavesBefore[R];
    const float* bl = bounds + 6 * (size_t)L;
    const float* br = bounds + 6 * (size_t)R;
    // BVH::buildHitTable (BVH.cpp:300-308, 339-340): child order of face i by the strict comparison of the children's centroids
    const float3 cl = (f3(bl[0], bl[1], bl[2]) + f3(bl[3], bl[4], bl[5])) * 0.5f, cr = (f3(br[0], br[1], br[2]) + f3(br[3], br[4], br[5])) * 0.5f;   // AABB::centroid
    const int bits = (cl.x > cr.x ? 1 : 0) | (cl.x < cr.x ? 2 : 0) | (cl.y > cr.y ? 4 : 0) | (cl.y < cr.y ? 8 : 0) | (cl.z > cr.z ? 16 : 0) | (cl.z < cr.z ? 32 : 0);
    float4* o = out + 4 * (size_t)(p - leavesBefore[p]);
    o[0] = make_float4(bl[0], bl[1], bl[3], bl[4]);
    o[1] = make_float4(bl[2], bl[5], __int_as_float(refL), __int_as_float(refR));
    o[2] = make_float4(br[0], br[1], br[3], br[4]);
    o[3] = make_float4(br[2], br[5], __int_as_float(bits), 0.0f);
}
}  // namespace bvh2b

// out: (n + 1) / 2 - 1 ... = T - 1 records of 4 float4 (n = 2T - 1 nodes); maxDepthOut: depth of the deepest leaf
static cudaError_t buildBvh2OnDevice(const float* bounds, const int* sizeIndices, const int n, float4* out, int* maxDepthOut) {
#define ZLB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)
    int *flag = nullptr, *scan = nullptr, *dDepth = nullptr; char* temp = nullptr;
    size_t tempBytes = 0;
    ZLB(cudaMalloc((void**)&flag, (size_t)n * sizeof(int)));
    ZLB(cudaMalloc((void**)&scan, (size_t)n * sizeof(int)));
    ZLB(cudaMalloc((void**)&dDepth, sizeof(int)));
    ZLB(cudaMemset(dDepth, 0, sizeof(int)));
    ZLB(cub::DeviceScan::ExclusiveSum(nullptr, tempBytes, flag, scan, n));
    ZLB(cudaMalloc((void**)&temp, tempBytes));
    const int B = 256, grid = (n + B - 1) / B;
    bvh2b::leafFlagKernel<<<grid, B>>>(sizeIndices, n, flag);
    ZLB(cub::DeviceScan::ExclusiveSum(temp, tempBytes, flag, scan, n));
    bvh2b::recordKernel<<<grid, B>>>(bounds, sizeIndices, scan, n, out);
    bvh2b::depthKernel<<<grid, B>>>(sizeIndices, n, dDepth);
    ZLB(cudaGetLastError());
    ZLB(cudaMemcpy(maxDepthOut, dDepth, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(flag); cudaFree(scan); cudaFree(dDepth); cudaFree(temp);
#undef ZLB
    return cudaSuccess;
}

}  // namespace zl
