#include "BVH.h"
#include <algorithm>
#include <chrono>
#include <climits>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace zillum {

static inline int bucketIndex(float c, float axisMin, float axisMax) {
    // int b = 16 * (c - axisMin) / (axisMax - axisMin), clamped (BVH.cpp:106-107,256-257).
    // NaN (axisMax == axisMin) converts like x86 cvttss2si -> INT_MIN -> clamps to bucket 0.
    float f = 16.0f * (c - axisMin) / (axisMax - axisMin);
    int b = (f == f && f < 2147483648.0f && f >= -2147483648.0f) ? (int)f : INT_MIN;
    return std::max(std::min(b, 15), 0);
}

PackedBVH BVH::build(bool threadOnHost) {
    using clk = std::chrono::steady_clock;
    size_t nPrims = indices.size() / 3;
    primInfo.resize(nPrims);
    scratch.resize(nPrims);
    bucketId.resize(nPrims);
    treeSize = nPrims * 2 - 1;
    bounds.resize(treeSize);
    sizeIndices.resize(treeSize);

    AABB rootCentExtent;
    for (size_t i = 0; i < nPrims; i++) {
        PrimInfo h;
        h.bound = AABB(vertices[indices[i * 3 + 0]], vertices[indices[i * 3 + 1]], vertices[indices[i * 3 + 2]]);
        h.centroid = h.bound.centroid();
        h.index = (int)i;
        rootCentExtent.expand(h.centroid);
        primInfo[i] = h;
    }
    auto t0 = clk::now();
    quickBuild(rootCentExtent);
    auto t1 = clk::now();
    if (threadOnHost) buildHitTable();
    auto t2 = clk::now();
    buildSeconds = std::chrono::duration<double>(t1 - t0).count();
    flattenSeconds = std::chrono::duration<double>(t2 - t1).count();
    PackedBVH out{std::move(bounds), std::move(hitTable), std::move(sizeIndices)};
    primInfo.clear(); primInfo.shrink_to_fit();
    scratch.clear(); scratch.shrink_to_fit();
    return out;
}

// Processes `rec` and everything below it depth-first.  Records covering fewer than
// `deferBelow` primitives are pushed to `deferred` instead (they become parallel tasks).
void BVH::buildRange(BuildRec root, std::vector<BuildRec>* deferred, int deferBelow) {
    std::vector<BuildRec> stack;
    stack.push_back(root);
    while (!stack.empty()) {
        BuildRec rec = stack.back();
        stack.pop_back();
        const int offset = rec.offset, splitDim = rec.splitDim, l = rec.l, r = rec.r;
        const int nBoxes = r - l + 1;
        if (deferred && nBoxes < deferBelow && !(rec.offset == root.offset)) { deferred->push_back(rec); continue; }
        sizeIndices[offset] = (nBoxes == 1) ? (primInfo[l].index | BVH_LEAF_MASK) : (nBoxes * 2 - 1);
        if (nBoxes == 1) { bounds[offset] = primInfo[l].bound; continue; }
        if (nBoxes == 2) {
            bounds[offset] = AABB(primInfo[l].bound, primInfo[r].bound);
            if (primInfo[l].centroid[splitDim] > primInfo[r].centroid[splitDim]) std::swap(primInfo[l], primInfo[r]);
            stack.push_back({offset + 2, AABB(primInfo[r].centroid), 0, r, r});
            stack.push_back({offset + 1, AABB(primInfo[l].centroid), 0, l, l});
            continue;
        }
        const float axisMin = rec.nodeExtent.pMin[splitDim], axisMax = rec.nodeExtent.pMax[splitDim];
        int count[16] = {0};
        AABB box[16];
        for (int i = l; i <= r; i++) {
            int b = bucketIndex(primInfo[i].centroid[splitDim], axisMin, axisMax);
            bucketId[i] = (uint8_t)b;
            count[b]++;
            box[b].expand(primInfo[i].bound);
        }
        int preCount[16], sufCount[16];
        AABB preBox[16], sufBox[16];
        preCount[0] = count[0]; preBox[0] = box[0];
        sufCount[15] = count[15]; sufBox[15] = box[15];
        for (int i = 1; i < 16; i++) {
            preCount[i] = preCount[i - 1] + count[i];
            preBox[i] = AABB(preBox[i - 1], box[i]);
            sufCount[15 - i] = sufCount[16 - i] + count[15 - i];
            sufBox[15 - i] = AABB(sufBox[16 - i], box[15 - i]);
        }
        bounds[offset] = preBox[15];
        int splitBucket = 0;
        float minCost = preCount[0] * preBox[0].surfaceArea() + sufCount[1] * sufBox[1].surfaceArea();
        for (int i = 1; i < 15; i++) {
            float cost = preCount[i] * preBox[i].surfaceArea() + sufCount[i + 1] * sufBox[i + 1].surfaceArea();
            if (cost < minCost) { minCost = cost; splitBucket = i; }
        }
        // stable on the left, reversed on the right (partition<16>, BVH.cpp:97-114)
        std::copy(primInfo.begin() + l, primInfo.begin() + r + 1, scratch.begin() + l);
        int pl = l, pr = r + 1;
        for (int i = l; i <= r; i++) {
            if ((int)bucketId[i] <= splitBucket) primInfo[pl++] = scratch[i];
            else primInfo[--pr] = scratch[i];
        }
        if (pr == r + 1) pr--;
        const int splitPoint = pr - 1;
        AABB lchCentBox, rchCentBox;
        for (int i = l; i <= splitPoint; i++) lchCentBox.expand(primInfo[i].centroid);
        for (int i = splitPoint + 1; i <= r; i++) rchCentBox.expand(primInfo[i].centroid);
        stack.push_back({offset + 2 * (splitPoint - l) + 2, rchCentBox, rchCentBox.maxExtent(), splitPoint + 1, r});
        stack.push_back({offset + 1, lchCentBox, lchCentBox.maxExtent(), l, splitPoint});
    }
}

void BVH::quickBuild(const AABB& rootExtent) {
    BuildRec root{0, rootExtent, rootExtent.maxExtent(), 0, (int)primInfo.size() - 1};
    int n = (int)primInfo.size();
    int threads = 1;
#ifdef _OPENMP
    threads = omp_get_max_threads();
#endif
    if (n < 65536 || threads == 1) { buildRange(root, nullptr, 0); return; }
    // top of the tree sequentially, then the deferred subtrees in parallel
    std::vector<BuildRec> tasks;
    buildRange(root, &tasks, std::max(n / (threads * 16), 4096));
    std::sort(tasks.begin(), tasks.end(), [](const BuildRec& a, const BuildRec& b) { return (a.r - a.l) > (b.r - b.l); });
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < (int)tasks.size(); i++) buildRange(tasks[i], nullptr, 0);
}

void BVH::buildHitTable() {
    hitTable.resize(treeSize * 18);
#pragma omp parallel for schedule(static, 1)
    for (int face = 0; face < 6; face++) {
        std::vector<int> stack;
        stack.reserve(256);
        int* table = hitTable.data() + treeSize * 3 * face;
        const int axis = face >> 1;
        const bool plus = (face & 1) == 0;       // X+ Y+ Z+ compare with '>', X- Y- Z- with '<'
        int index = 0;
        stack.push_back(0);
        while (!stack.empty()) {
            int k = stack.back();
            stack.pop_back();
            bool isLeaf = (sizeIndices[k] & BVH_LEAF_MASK) != 0;
            int nodeSize = isLeaf ? 1 : sizeIndices[k];
            table[index * 3 + 0] = k;
            table[index * 3 + 1] = isLeaf ? (sizeIndices[k] ^ BVH_LEAF_MASK) : -1;
            table[index * 3 + 2] = index + nodeSize;
            index++;
            if (isLeaf) continue;
            int lSize = sizeIndices[k + 1];
            if (lSize & BVH_LEAF_MASK) lSize = 1;
            int lch = k + 1, rch = k + 1 + lSize;
            float a = bounds[lch].centroid()[axis], b = bounds[rch].centroid()[axis];
            if (!(plus ? (a > b) : (a < b))) std::swap(lch, rch);
            stack.push_back(rch);
            stack.push_back(lch);
        }
    }
}

}  // namespace zillum
