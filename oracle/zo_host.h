// ORACLE — test infrastructure only (see zo_vec.h header).  Pinned to oracle/_ref by tests/test_ref_parity.py.
// zo_host.h — restatement of the host-side preparation that feeds the shaders:
//   src/accelerator/{AABB,BVH}.cpp (binned-SAH quickBuild + six-direction hit table),
//   src/math/AliasTable.h, src/core/EnvironmentMap.cpp (two-level alias tables),
//   src/core/Sampler.cpp (sobolSample), src/core/Camera.cpp (update),
//   src/core/Scene.cpp:200-243 (light power table), Model.cpp:62-72 (model matrix).
#pragma once
#include <algorithm>
#include <climits>
#include <utility>
#include <vector>
#include "zo_vec.h"
#include "../include/zillum_cuda.h"

namespace zo {

// ---- AABB.h:10-38, AABB.cpp:3-37 ----
struct AABB {
    vec3 pMin, pMax;
    AABB() : pMin(1e8f), pMax(-1e8f) {}
    AABB(vec3 p) : pMin(p), pMax(p) {}
    AABB(vec3 a, vec3 b) : pMin(a), pMax(b) {}
    AABB(vec3 va, vec3 vb, vec3 vc) { pMin = gmin(gmin(va, vb), vc); pMax = gmax(gmax(va, vb), vc); }
    static AABB join(const AABB& a, const AABB& b) { return AABB(gmin(a.pMin, b.pMin), gmax(a.pMax, b.pMax)); }
    void expand(const AABB& r) { pMin = gmin(pMin, r.pMin); pMax = gmax(pMax, r.pMax); }
    vec3 centroid() const { return (pMin + pMax) * 0.5f; }
    float surfaceArea() const { vec3 v = pMax - pMin; return 2.0f * (v.x * v.y + v.y * v.z + v.z * v.x); }
    int maxExtent() const {
        vec3 v = pMax - pMin;
        if (v.x > v.y) return v.x > v.z ? 0 : 2;
        return v.y > v.z ? 1 : 2;
    }
};

struct PackedBVH { std::vector<float> bounds; std::vector<int32_t> hitTable; int treeSize = 0; };

// float -> int as the x86 cvttss2si the reference binary executes: truncation, and the
// "integer indefinite" INT_MIN for NaN / out-of-range (reached when axisMax == axisMin).
inline int truncToInt(float f) {
    if (!(f == f) || f >= 2147483648.0f || f < -2147483648.0f) return INT_MIN;
    return (int)f;
}

// ---- BVH.cpp:116-144, 217-296, 298-346 ----
inline PackedBVH buildBVH(const float* vertices, const uint32_t* indices, int numTriangles) {
    struct PrimInfo { AABB bound; vec3 centroid; int index; };
    struct BuildRec { int offset; AABB nodeExtent; int splitDim; int l, r; };
    const int LeafMask = (int)0x80000000u;
    const int NumBuckets = 16;
    PackedBVH out;
    if (numTriangles <= 0) return out;

    auto V = [&](uint32_t i) { return vec3(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]); };
    std::vector<PrimInfo> prim(numTriangles);
    int treeSize = numTriangles * 2 - 1;
    std::vector<AABB> bounds(treeSize);
    std::vector<int> sizeIndices(treeSize);

    AABB rootCentExtent;
    for (int i = 0; i < numTriangles; i++) {                                   // BVH.cpp:126-135
        PrimInfo h;
        h.bound = AABB(V(indices[i * 3 + 0]), V(indices[i * 3 + 1]), V(indices[i * 3 + 2]));
        h.centroid = h.bound.centroid();
        h.index = i;
        rootCentExtent.expand(AABB(h.centroid));
        prim[i] = h;
    }

    // quickBuild (BVH.cpp:217-296)
    std::vector<BuildRec> stack;
    stack.push_back({0, rootCentExtent, rootCentExtent.maxExtent(), 0, numTriangles - 1});
    std::vector<PrimInfo> tmp;
    while (!stack.empty()) {
        BuildRec rec = stack.back();
        stack.pop_back();
        int offset = rec.offset, splitDim = rec.splitDim, l = rec.l, r = rec.r;
        int size = (r - l) * 2 + 1;
        sizeIndices[offset] = (size == 1) ? (prim[l].index | LeafMask) : size;
        if (l == r) { bounds[offset] = prim[l].bound; continue; }
        int nBoxes = r - l + 1;
        if (nBoxes == 2) {
            bounds[offset] = AABB::join(prim[l].bound, prim[r].bound);
            if (prim[l].centroid[splitDim] > prim[r].centroid[splitDim]) std::swap(prim[l], prim[r]);
            stack.push_back({offset + 2, AABB(prim[r].centroid), 0, r, r});
            stack.push_back({offset + 1, AABB(prim[l].centroid), 0, l, l});
            continue;
        }
        float axisMin = rec.nodeExtent.pMin[splitDim];
        float axisMax = rec.nodeExtent.pMax[splitDim];
        struct Bucket { int count = 0; AABB box; };
        Bucket buckets[NumBuckets], prefix[NumBuckets], suffix[NumBuckets];
        auto bucketOf = [&](const PrimInfo& p) {
            int b = truncToInt((float)NumBuckets * (p.centroid[splitDim] - axisMin) / (axisMax - axisMin));
            b = std::max(std::min(b, NumBuckets - 1), 0);
            return b;
        };
        for (int i = l; i <= r; i++) {
            int b = bucketOf(prim[i]);
            buckets[b].count++;
            buckets[b].box.expand(prim[i].bound);
        }
        prefix[0] = buckets[0];
        suffix[NumBuckets - 1] = buckets[NumBuckets - 1];
        for (int i = 1; i < NumBuckets; i++) {
            prefix[i].count = prefix[i - 1].count + buckets[i].count;
            prefix[i].box = AABB::join(prefix[i - 1].box, buckets[i].box);
            suffix[NumBuckets - 1 - i].count = suffix[NumBuckets - i].count + buckets[NumBuckets - i - 1].count;
            suffix[NumBuckets - 1 - i].box = AABB::join(suffix[NumBuckets - i].box, buckets[NumBuckets - i - 1].box);
        }
        bounds[offset] = prefix[NumBuckets - 1].box;

        int splitPoint = 0;
        float minCost = (float)prefix[0].count * prefix[0].box.surfaceArea() +
                        (float)suffix[1].count * suffix[1].box.surfaceArea();
        for (int i = 1; i < NumBuckets - 1; i++) {
            float cost = (float)prefix[i].count * prefix[i].box.surfaceArea() +
                         (float)suffix[i + 1].count * suffix[i + 1].box.surfaceArea();
            if (cost < minCost) { minCost = cost; splitPoint = i; }
        }
        // partition<16> (BVH.cpp:97-114): left part keeps order, right part is filled from the back
        tmp.assign(prim.begin() + l, prim.begin() + l + nBoxes);
        int pl = 0, pr = nBoxes;
        for (int i = 0; i < nBoxes; i++) {
            int b = bucketOf(tmp[i]);
            if (b <= splitPoint) prim[l + pl++] = tmp[i]; else prim[l + --pr] = tmp[i];
        }
        if (pr == nBoxes) pr--;
        splitPoint = l + pr - 1;

        AABB lchCentBox, rchCentBox;
        for (int i = l; i <= splitPoint; i++) lchCentBox.expand(AABB(prim[i].centroid));
        for (int i = splitPoint + 1; i <= r; i++) rchCentBox.expand(AABB(prim[i].centroid));
        stack.push_back({offset + 2 * (splitPoint - l) + 2, rchCentBox, rchCentBox.maxExtent(), splitPoint + 1, r});
        stack.push_back({offset + 1, lchCentBox, lchCentBox.maxExtent(), l, splitPoint});
    }

    // buildHitTable (BVH.cpp:298-346)
    out.treeSize = treeSize;
    out.hitTable.resize((size_t)treeSize * 18);
    std::vector<int> st(treeSize);
    for (int face = 0; face < 6; face++) {
        size_t tableOffset = (size_t)treeSize * 3 * face;
        int top = 0, index = 0;
        st[top++] = 0;
        int axis = face / 2;
        bool greater = (face % 2) == 0;     // X+,Y+,Z+ use '>', X-,Y-,Z- use '<'
        while (top) {
            int k = st[--top];
            bool isLeaf = (sizeIndices[k] & LeafMask) != 0;
            int nodeSize = isLeaf ? 1 : sizeIndices[k];
            out.hitTable[tableOffset + (size_t)index * 3 + 0] = k;
            out.hitTable[tableOffset + (size_t)index * 3 + 1] = isLeaf ? (sizeIndices[k] ^ LeafMask) : -1;
            out.hitTable[tableOffset + (size_t)index * 3 + 2] = index + nodeSize;
            index++;
            if (isLeaf) continue;
            int lSize = sizeIndices[k + 1];
            if (lSize & LeafMask) lSize = 1;
            int lch = k + 1, rch = k + 1 + lSize;
            float a = bounds[lch].centroid()[axis], b = bounds[rch].centroid()[axis];
            bool keep = greater ? (a > b) : (a < b);
            if (!keep) std::swap(lch, rch);
            st[top++] = rch;
            st[top++] = lch;
        }
    }
    out.bounds.resize((size_t)treeSize * 6);
    for (int i = 0; i < treeSize; i++) {
        out.bounds[6 * (size_t)i + 0] = bounds[i].pMin.x; out.bounds[6 * (size_t)i + 1] = bounds[i].pMin.y; out.bounds[6 * (size_t)i + 2] = bounds[i].pMin.z;
        out.bounds[6 * (size_t)i + 3] = bounds[i].pMax.x; out.bounds[6 * (size_t)i + 4] = bounds[i].pMax.y; out.bounds[6 * (size_t)i + 5] = bounds[i].pMax.z;
    }
    return out;
}

// ---- AliasTable.h:12-56 ----
inline void buildAliasTable(const float* pdf, int n, int32_t* alias, float* prob) {
    typedef std::pair<int, float> Element;
    for (int i = 0; i < n; i++) prob[i] = pdf[i];
    float sumPdf = 0.0f;
    for (int i = 0; i < n; i++) sumPdf += prob[i];
    float sumInv = (float)n / sumPdf;
    std::vector<Element> greater(n * 2 + 1), lesser(n * 2 + 1);
    int gTop = 0, lTop = 0;
    for (int i = 0; i < n; i++) {
        prob[i] *= sumInv;
        if (prob[i] < 1.0f) lesser[lTop++] = Element(i, prob[i]); else greater[gTop++] = Element(i, prob[i]);
    }
    while (gTop != 0 && lTop != 0) {
        Element le = lesser[--lTop], ge = greater[--gTop];
        int l = le.first, g = ge.first; float pl = le.second, pg = ge.second;
        alias[l] = g; prob[l] = pl;
        pg += pl - 1.0f;
        if (pg < 1.0f) lesser[lTop++] = Element(g, pg); else greater[gTop++] = Element(g, pg);
    }
    while (gTop != 0) { Element ge = greater[--gTop]; alias[ge.first] = ge.first; prob[ge.first] = ge.second; }
    while (lTop != 0) { Element le = lesser[--lTop]; alias[le.first] = le.first; prob[le.first] = le.second; }
}

// ---- EnvironmentMap.cpp:66-114 (strided variant; residual entries get prob 1) ----
inline float setupAliasTableStrided(int32_t* alias, float* pdf, int n, int stride) {
    typedef std::pair<int, float> Element;
    float sum = 0.0f;
    for (int i = 0, off = 0; i < n; i++, off += stride) sum += pdf[off];
    float sumInv = n / sum;
    std::vector<Element> greater(n * 2 + 1), lesser(n * 2 + 1);
    int gTop = 0, lTop = 0;
    for (int i = 0, off = 0; i < n; i++, off += stride) {
        pdf[off] *= sumInv;
        if (pdf[off] < 1.0f) lesser[lTop++] = Element(i, pdf[off]); else greater[gTop++] = Element(i, pdf[off]);
    }
    while (gTop != 0 && lTop != 0) {
        Element le = lesser[--lTop], ge = greater[--gTop];
        int l = le.first, g = ge.first; float pl = le.second, pg = ge.second;
        alias[l * stride] = g; pdf[l * stride] = pl;
        pg += pl - 1.0f;
        if (pg < 1.0f) lesser[lTop++] = Element(g, pg); else greater[gTop++] = Element(g, pg);
    }
    while (gTop != 0) { int g = greater[--gTop].first; alias[g * stride] = g; pdf[g * stride] = 1.0f; }
    while (lTop != 0) { int l = lesser[--lTop].first; alias[l * stride] = l; pdf[l * stride] = 1.0f; }
    return sum;
}

// ---- EnvironmentMap.cpp:8-59: (W+1)xH tables; returns mSumPdf (float, before the int truncation) ----
inline float buildEnvTables(const float* rgb, int width, int height, int32_t* alias, float* pdf) {
    auto offset = [width](int i, int j) { return (size_t)i * (width + 1) + j; };
    for (int i = 0; i < height; i++)
        for (int j = 0; j < width; j++) {
            const float* p = rgb + 3 * ((size_t)i * width + j);
            float lum = 0.2126f * p[0] + 0.7152f * p[1] + 0.0722f * p[2];
            pdf[offset(i, j)] = lum * std::sin((float)(i + 0.5f) / height * 3.141592653589793f);
        }
    for (int i = 0; i < height; i++)
        pdf[offset(i, width)] = setupAliasTableStrided(alias + offset(i, 0), pdf + offset(i, 0), width, 1);
    return setupAliasTableStrided(alias + width, pdf + width, height, width + 1);
}

// ---- Sampler.cpp:19-28 ----
inline uint32_t sobolSample(const uint32_t* matrices, uint32_t index, int dim, uint32_t scramble = 0) {
    uint32_t r = scramble;
    for (int i = dim * 32; index != 0; index >>= 1, i++)
        if (index & 1u) r ^= matrices[i];
    return r;
}

// ---- Camera.cpp:149-162 + the uniforms of NaivePath.cpp:49-59 ----
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
inline void cameraUpdate(const float pos[3], const float angleDeg[3], float fovDeg, float aspect,
                         float lensRadius, float focalDist, ZlCamera* out) {
    float x = std::sin(radians(angleDeg[0])) * std::cos(radians(angleDeg[1]));
    float y = std::cos(radians(angleDeg[0])) * std::cos(radians(angleDeg[1]));
    float z = std::sin(radians(angleDeg[1]));
    vec3 front = normalize(vec3(x, y, z));
    vec3 right = normalize(cross(front, vec3(0.0f, 0.0f, 1.0f)));
    // glm::rotate(mat4(1), angle.z /* radians, not converted */, front) applied to `right`
    float a = angleDeg[2];
    float c = std::cos(a), s = std::sin(a);
    vec3 axis = normalize(front);
    vec3 t = axis * (1.0f - c);
    vec3 r0(c + t.x * axis.x, t.x * axis.y + s * axis.z, t.x * axis.z - s * axis.y);
    vec3 r1(t.y * axis.x - s * axis.z, c + t.y * axis.y, t.y * axis.z + s * axis.x);
    vec3 r2(t.z * axis.x + s * axis.y, t.z * axis.y - s * axis.x, c + t.z * axis.z);
    right = normalize(r0 * right.x + r1 * right.y + r2 * right.z);
    vec3 up = normalize(cross(right, front));
    out->F[0] = front.x; out->F[1] = front.y; out->F[2] = front.z;
    out->R[0] = right.x; out->R[1] = right.y; out->R[2] = right.z;
    out->U[0] = up.x; out->U[1] = up.y; out->U[2] = up.z;
    mat3 inv = inverse(mat3(right, up, front));
    out->matInv[0] = inv.c0.x; out->matInv[1] = inv.c0.y; out->matInv[2] = inv.c0.z;
    out->matInv[3] = inv.c1.x; out->matInv[4] = inv.c1.y; out->matInv[5] = inv.c1.z;
    out->matInv[6] = inv.c2.x; out->matInv[7] = inv.c2.y; out->matInv[8] = inv.c2.z;
    out->pos[0] = pos[0]; out->pos[1] = pos[1]; out->pos[2] = pos[2];
    out->tanFOV = std::tan(radians(fovDeg * 0.5f));
    out->asp = aspect;
    out->lensRadius = lensRadius;
    out->focalDist = focalDist;
}

// ---- Scene.cpp:200-243: per-light-mesh power split by area; returns lightSumPdf ----
// meshFirstTri/meshNumTris: triangle ranges (global ids) of each light mesh; meshPower: 3 floats each.
inline float buildLightTable(const float* vertices, const uint32_t* indices, int numMeshes,
                             const int* meshFirstTri, const int* meshNumTris, const float* meshPower,
                             float* lightPowerOut /*3 per light tri*/, float* pdfOut) {
    auto V = [&](uint32_t i) { return vec3(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]); };
    float lightSumPdf = 0.0f;
    int o = 0;
    for (int m = 0; m < numMeshes; m++) {
        vec3 sumPower(meshPower[3 * m], meshPower[3 * m + 1], meshPower[3 * m + 2]);
        float sumArea = 0.0f;
        for (int i = 0; i < meshNumTris[m]; i++) {
            int t = meshFirstTri[m] + i;
            vec3 va = V(indices[3 * t]), vb = V(indices[3 * t + 1]), vc = V(indices[3 * t + 2]);
            sumArea += length(cross(vc - va, vb - va));
        }
        for (int i = 0; i < meshNumTris[m]; i++) {
            int t = meshFirstTri[m] + i;
            vec3 va = V(indices[3 * t]), vb = V(indices[3 * t + 1]), vc = V(indices[3 * t + 2]);
            float area = length(cross(vc - va, vb - va));
            vec3 power = sumPower * area / sumArea;
            lightPowerOut[3 * o] = power.x; lightPowerOut[3 * o + 1] = power.y; lightPowerOut[3 * o + 2] = power.z;
            float lum = dot(power, vec3(0.299f, 0.587f, 0.114f));
            pdfOut[o] = lum;
            lightSumPdf += lum;
            o++;
        }
    }
    return lightSumPdf;
}

}  // namespace zo
