// Host BVH build + six-direction multiple-threaded BVH (MTBVH) flattening.
// Same interface as the reference's accelerator (src/accelerator/BVH.h:13-47):
//   BVH(vertices, indices).build() -> PackedBVH{bounds, hitTable}
// and the same tree: 16-bucket binned SAH on the centroid extent, one triangle per leaf,
// nodes in DFS pre-order, six hit/miss-link orderings (src/accelerator/BVH.cpp:116-144,
// 217-296, 298-346).  The implementation is ours: one scratch buffer instead of an
// allocation per partition, bucket ids cached per primitive, subtrees built in parallel
// (each build record only touches its own primitive range and node range, so the result
// does not depend on scheduling), and the six orderings flattened concurrently.
#pragma once
#include <cstdint>
#include <vector>
#include "Math.h"

namespace zillum {

class AABB {
public:
    AABB() : pMin(1e8f), pMax(-1e8f) {}
    AABB(const Vec3f& p) : pMin(p), pMax(p) {}
    AABB(const Vec3f& lo, const Vec3f& hi) : pMin(lo), pMax(hi) {}
    AABB(const Vec3f& va, const Vec3f& vb, const Vec3f& vc) : pMin(vmin(vmin(va, vb), vc)), pMax(vmax(vmax(va, vb), vc)) {}
    AABB(const AABB& a, const AABB& b) : pMin(vmin(a.pMin, b.pMin)), pMax(vmax(a.pMax, b.pMax)) {}
    void expand(const AABB& rhs) { pMin = vmin(pMin, rhs.pMin); pMax = vmax(pMax, rhs.pMax); }
    Vec3f centroid() const { return (pMin + pMax) * 0.5f; }
    float surfaceArea() const { Vec3f v = pMax - pMin; return 2.0f * (v.x * v.y + v.y * v.z + v.z * v.x); }
    int maxExtent() const {
        Vec3f v = pMax - pMin;
        if (v.x > v.y) return v.x > v.z ? 0 : 2;
        return v.y > v.z ? 1 : 2;
    }
    Vec3f pMin, pMax;
};

const int BVH_LEAF_MASK = (int)0x80000000u;

struct PackedBVH {
    std::vector<AABB> bounds;       // 2T-1 nodes, pre-order
    std::vector<int> hitTable;      // 6 * (2T-1) * (nodeIndex, primIndex|-1, missIndex); empty when built with threadOnHost = false
    std::vector<int> sizeIndices;   // the pre-order tree itself: subtree node count, or primIndex | BVH_LEAF_MASK (BVH.h:39)
};

class BVH {
public:
    BVH(const std::vector<Vec3f>& vertices, const std::vector<uint32_t>& indices) : vertices(vertices), indices(indices) {}
    // threadOnHost = false skips buildHitTable(): the six orderings are then threaded on the device from
    // bounds + sizeIndices (ZlSceneDesc::sizeIndices, threadMtbvhKernel)
    PackedBVH build(bool threadOnHost = true);
    double buildSeconds = 0.0, flattenSeconds = 0.0;

private:
    struct PrimInfo { AABB bound; Vec3f centroid; int index; };
    struct BuildRec { int offset; AABB nodeExtent; int splitDim; int l, r; };
    void quickBuild(const AABB& rootExtent);
    void buildRange(BuildRec rec, std::vector<BuildRec>* deferred, int deferBelow);
    void buildHitTable();

    const std::vector<Vec3f>& vertices;
    const std::vector<uint32_t>& indices;
    std::vector<PrimInfo> primInfo, scratch;
    std::vector<uint8_t> bucketId;
    std::vector<AABB> bounds;
    std::vector<int> sizeIndices;
    std::vector<int> hitTable;
    size_t treeSize = 0;
};

}  // namespace zillum
