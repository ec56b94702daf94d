// zl_traverse.cuh — stackless six-direction MTBVH traversal: bvhHit (closest hit) and bvhTest
// (any hit), restating intersection.glsl:226-329 (boxHit), :63-121 (intersectTriangle),
// :367-427 (bvhTest / bvhHit) with bit-exact results.
//
// What changed for B200 is only where the bytes come from: one 32-byte threaded node record
// (two LDG.128 to the same sector) replaces three hit-table texel fetches + two bounds
// fetches, and three LDG.128 of pre-gathered vertices replace three index + three vertex
// fetches.  Per-ray invariants (reciprocal direction, the axis-parallel / near-zero
// classification of boxHit) are hoisted out of the loop; they are the same IEEE values the
// GLSL recomputes at every node.
#pragma once
#include "zl_scene.cuh"

namespace zl {

struct Ray { float3 ori, dir; };
ZL_DEV Ray makeRay(float3 o, float3 d) { Ray r; r.ori = o; r.dir = d; return r; }
ZL_DEV float3 rayPoint(Ray r, float t) { return r.ori + r.dir * t; }                     // intersection.glsl:26-29
ZL_DEV Ray rayOffseted(float3 ori, float3 dir) { return makeRay(ori + dir * 1e-4f, dir); }  // :31-37
ZL_DEV Ray rayOffseted(Ray r) { return rayOffseted(r.ori, r.dir); }                      // :39-42

// per-ray constants of boxHit
struct RayPrep {
    float3 o, d, dInv;
    int mode;   // 0..2: |d.axis| > 1-eps (axis-parallel fast path); 3: general
    bool smallX, smallY, smallZ;
    bool pure;  // general mode with no near-zero component: the common case, no NaN can arise in the slab test
};
ZL_DEV RayPrep prepareRay(Ray ray) {
    const float eps = 1e-6f;
    RayPrep p;
    p.o = ray.ori; p.d = ray.dir;
    p.dInv = f3(1.0f / ray.dir.x, 1.0f / ray.dir.y, 1.0f / ray.dir.z);
    p.mode = (fabsf(ray.dir.x) > 1.0f - eps) ? 0 : (fabsf(ray.dir.y) > 1.0f - eps) ? 1 : (fabsf(ray.dir.z) > 1.0f - eps) ? 2 : 3;
    p.smallX = fabsf(ray.dir.x) < eps; p.smallY = fabsf(ray.dir.y) < eps; p.smallZ = fabsf(ray.dir.z) < eps;
    p.pure = (p.mode == 3) && !p.smallX && !p.smallY && !p.smallZ;
    return p;
}

// Conservative rejection for the "near-zero component" branches of boxHit.  The reference
// IGNORES the slab of an axis whose |d| < 1e-6 (intersection.glsl:291-319), so such a ray
// enters every node that overlaps it in the other two axes — up to 3e5 node visits per ray
// in the Rungholt-class scene, all of whose leaf triangles then miss.  While the ray is
// inside the node (t <= tMax) it cannot drift further than |d.a| * tMax along the ignored
// axis, so a node whose slab is further away than that (plus a guard 20x larger than the
// acceptance slop of the triangle test) cannot contain a hit and is skipped.  Hit ids and
// distances are unchanged (tests/test_gpu_traversal.py::test_near_zero_direction_rays); only
// the number of visited nodes drops.  CULL=false keeps the reference's exact visit sequence
// (used by the counting build and by zl_trace_rays' step counters).
ZL_DEV bool outsideIgnoredSlab(float o, float d, float lo, float hi, float tMax) {
    float reach = fabsf(d) * tMax + 2e-5f * tMax + 1e-5f * (fabsf(o) + 1.0f);
    return (o - reach > hi) || (o + reach < lo);
}

// The general branch of boxHit (intersection.glsl:278-289, 321-328) for rays whose three direction
// components all have 1e-6 <= |d| <= 1 - 1e-6.  Then 1/d is finite, no product is NaN, and GLSL
// min/max coincide with fminf/fmaxf (one FMNMX instead of a compare + select; the sign of a zero
// result may differ but is never observed: the values are only compared).  Same sub, mul and
// compare sequence, so the accepted set and tMin are bit-identical.
ZL_DEV bool boxHitPure(float3 pMin, float3 pMax, const RayPrep& r, float& tMin) {
    float3 vta = (pMin - r.o) * r.dInv, vtb = (pMax - r.o) * r.dInv;
    float3 vtMin = f3(fminf(vta.x, vtb.x), fminf(vta.y, vtb.y), fminf(vta.z, vtb.z));
    float3 vtMax = f3(fmaxf(vta.x, vtb.x), fmaxf(vta.y, vtb.y), fmaxf(vta.z, vtb.z));
    float3 dt = vtMax - vtMin;
    float tyz = vtMax.z - vtMin.y, tzx = vtMax.x - vtMin.z, txy = vtMax.y - vtMin.x;
    if (dt.y + dt.z > tyz && dt.z + dt.x > tzx && dt.x + dt.y > txy) {
        tMin = fmaxf(fmaxf(vtMin.x, vtMin.y), vtMin.z);
        float tMax = fminf(fminf(vtMax.x, vtMax.y), vtMax.z);
        return tMax >= 0.0f && tMax >= tMin;
    }
    return false;
}

// intersection.glsl:226-329.  Branch order and comparison strictness are the reference's.
template <bool CULL>
ZL_DEV bool boxHit(float3 pMin, float3 pMax, const RayPrep& r, float& tMin) {
    float tMax;
    const float3 o = r.o;
    if (r.mode != 3) {
        if (r.mode == 0) {
            if (o.y > pMin.y && o.y < pMax.y && o.z > pMin.z && o.z < pMax.z) {
                float ta = (pMin.x - o.x) * r.dInv.x, tb = (pMax.x - o.x) * r.dInv.x;
                tMin = gmin(ta, tb); tMax = gmax(ta, tb);
                return tMax >= 0.0f && tMax >= tMin;
            }
            return false;
        }
        if (r.mode == 1) {
            if (o.x > pMin.x && o.x < pMax.x && o.z > pMin.z && o.z < pMax.z) {
                float ta = (pMin.y - o.y) * r.dInv.y, tb = (pMax.y - o.y) * r.dInv.y;
                tMin = gmin(ta, tb); tMax = gmax(ta, tb);
                return tMax >= 0.0f && tMax >= tMin;
            }
            return false;
        }
        if (o.x > pMin.x && o.x < pMax.x && o.y > pMin.y && o.y < pMax.y) {
            float ta = (pMin.z - o.z) * r.dInv.z, tb = (pMax.z - o.z) * r.dInv.z;
            tMin = gmin(ta, tb); tMax = gmax(ta, tb);
            return tMax >= 0.0f && tMax >= tMin;
        }
        return false;
    }
    float3 vta = (pMin - o) * r.dInv, vtb = (pMax - o) * r.dInv;
    float3 vtMin = f3(gmin(vta.x, vtb.x), gmin(vta.y, vtb.y), gmin(vta.z, vtb.z));
    float3 vtMax = f3(gmax(vta.x, vtb.x), gmax(vta.y, vtb.y), gmax(vta.z, vtb.z));
    float3 dt = vtMax - vtMin;
    float tyz = vtMax.z - vtMin.y, tzx = vtMax.x - vtMin.z, txy = vtMax.y - vtMin.x;
    if (r.smallX) {
        if (dt.y + dt.z > tyz) {
            tMin = gmax(vtMin.y, vtMin.z); tMax = gmin(vtMax.y, vtMax.z);
            if (CULL && outsideIgnoredSlab(o.x, r.d.x, pMin.x, pMax.x, tMax)) return false;
            return tMax >= 0.0f && tMax >= tMin;
        }
    }
    if (r.smallY) {
        if (dt.z + dt.x > tzx) {
            tMin = gmax(vtMin.z, vtMin.x); tMax = gmin(vtMax.z, vtMax.x);
            if (CULL && outsideIgnoredSlab(o.y, r.d.y, pMin.y, pMax.y, tMax)) return false;
            return tMax >= 0.0f && tMax >= tMin;
        }
    }
    if (r.smallZ) {
        if (dt.x + dt.y > txy) {
            tMin = gmax(vtMin.x, vtMin.y); tMax = gmin(vtMax.x, vtMax.y);
            if (CULL && outsideIgnoredSlab(o.z, r.d.z, pMin.z, pMax.z, tMax)) return false;
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

// intersection.glsl:63-109 (two-sided Moeller-Trumbore on the un-normalised determinant)
ZL_DEV bool intersectTriangle(float3 a, float3 b, float3 c, float3 o, float3 d, float& dist) {
    const float eps = 1e-6f;
    float3 ab = b - a, ac = c - a;
    float3 p = cross(d, ac);
    float det = dot(ab, p);
    if (fabsf(det) < eps) return false;
    float3 ao = o - a;
    if (det < 0) { ao = -ao; det = -det; }
    float u = dot(ao, p);
    if (u < 0.0f || u > det) return false;
    float3 q = cross(ao, ab);
    float v = dot(d, q);
    if (v < 0.0f || u + v > det) return false;
    float t = dot(ac, q) / det;
    dist = t;
    return t > 0.0f;
}

// One threaded node record = 32 bytes = one DRAM/L2 sector, fetched with a single 256-bit
// load (LDG.E.256, new on sm_100) through the read-only path.  Memory order (packNodeRecord, zl_scene.cuh):
// {pMin.x, pMin.y, pMax.x, pMax.y, pMin.z, pMax.z, prim, miss}; lo / hi is the {pMin, prim}{pMax, miss} view.
ZL_DEV void loadNode(const float4* __restrict__ nodes, int k, float4& lo, float4& hi) {
    const float4* p = nodes + 2 * (size_t)k;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(hi.x), "=f"(hi.y), "=f"(lo.z), "=f"(hi.z), "=f"(lo.w), "=f"(hi.w)
                 : "l"(p));
}

// same load with the L2 asked to bring in the whole 128-byte line (4 consecutive threaded records:
// the hit link k+1 is the next record, so 3 of 4 "inner" steps then find their node in L2)
ZL_DEV void loadNodeL2Line(const float4* __restrict__ nodes, int k, float4& lo, float4& hi) {
    const float4* p = nodes + 2 * (size_t)k;
    asm volatile("ld.global.nc.L2::128B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(hi.x), "=f"(hi.y), "=f"(lo.z), "=f"(hi.z), "=f"(lo.w), "=f"(hi.w)
                 : "l"(p));
}
ZL_DEV void prefetchNode(const float4* __restrict__ nodes, int k) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(nodes + 2 * (size_t)k));
}

struct TraceCounters { int nodes, tris; };   // bvhDebug-style visit counters (intersection.glsl:331-365)

// Packed FP32 pairs (sm_100: add / mul.rn.f32x2 -> FADD2 / FMUL2, one issue slot for two results, each half rounded
// exactly like the scalar instruction, so the bits stay those of the reference's arithmetic).
typedef unsigned long long f32x2_t;
ZL_DEV f32x2_t pack2(float a, float b) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
ZL_DEV void unpack2(f32x2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
ZL_DEV f32x2_t add2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
ZL_DEV f32x2_t sub2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
ZL_DEV f32x2_t mul2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// the record as four register pairs: {pMin.x, pMin.y}, {pMax.x, pMax.y}, {pMin.z, pMax.z}, {prim, miss}
// pol (warp-uniform, DScene::nodePolicy): 0 = default priorities; 1 = evict_last in L1 and L2 (SASS LDG.E.EL.ELL2.256): node records
// are the data that is re-used across rays, while the path-state records, queues and film that stream through the same caches are
// not — an eviction-priority split of the 126 MB L2 between the two (A/B switch ZL_NODE_POLICY, profiles/r2_trace_sweep.md).
ZL_DEV void loadNodePairs(unsigned long long faceBase, int k, f32x2_t& pLoXY, f32x2_t& pHiXY, f32x2_t& pZ, int& prim, int& miss, const int pol = 0) {
    f32x2_t links;
    const unsigned long long a = faceBase + 32ull * (unsigned long long)(long long)k;
    if (pol) asm volatile("ld.global.nc.L1::evict_last.L2::evict_last.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(pLoXY), "=l"(pHiXY), "=l"(pZ), "=l"(links) : "l"(a));
    else asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(pLoXY), "=l"(pHiXY), "=l"(pZ), "=l"(links) : "l"(a));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(prim), "=r"(miss) : "l"(links));
}

// The walk for "pure" rays (every |d| component in [1e-6, 1 - 1e-6]: all but a measure-zero set), written
// for the issue slots it costs: ncu shows the traversal kernels issue-bound at 48 warps/SM (and the ALU pipe, which
// runs FMNMX / FSETP at half rate, close behind).  The step is branch-free up to the leaf test: the accept decision
// of boxHitPure and the `boxDist > dist` cull are evaluated as predicates (same operations, same order of evaluation
// per operand, so the same bits), the successor is one select between k + 1 and the miss link, and the record
// address is one IMAD.WIDE on a face base pointer held in registers.
//   * the twelve `(p - o) * (1/d)` operations issue as three FADD2 + three FMUL2 on the record's register pairs
//     (`p + (-o)` is `p - o` bit for bit), `dt.xy` and `dt.xy + dt.z` as two more packed adds;
//   * OCT = the ray's direction octant (bit a set: d.a < 0), or -1 for "not known at compile time".  With lo <= hi
//     and rounding monotone, vtMin.a = min(vta.a, vtb.a) IS vta.a for d.a > 0 and vtb.a for d.a < 0 (equal values
//     differ at most in the sign of a zero, which only compares ever see), so an octant-specialised walk needs no
//     min / max per axis at all: six half-rate FMNMX gone.  Sorted queues are octant-uniform per warp (the sort key
//     leads with face and quadrant), so the callers pick the specialised walk per warp (traverseWarp below).
// Box step: 28 instructions for OCT >= 0 with d.x, d.y of one sign, 29 otherwise, 36 for OCT = -1 (the compiler's
// rendering of the general loop had 60, the scalar branch-free one 46).
template <bool ANYHIT, bool COUNT, int OCT>
ZL_DEV int traversePure(const float4* __restrict__ faceNodes, const float4* __restrict__ triPos, const int n, const RayPrep& rp, float& dist, TraceCounters* cnt,
                        const int pol = 0, int k = 0, int closest = -1) {
    if (n == 0 || k == n) return ANYHIT ? (closest >= 0 ? 1 : 0) : closest;
    unsigned long long base = (unsigned long long)faceNodes;
    asm volatile("" : "+l"(base));      // keep the face base as one 64-bit register value (not re-derived from the kernel parameter every step)
    const f32x2_t nOxy = pack2(-rp.o.x, -rp.o.y), nOzz = pack2(-rp.o.z, -rp.o.z);
    const f32x2_t iXy = pack2(rp.dInv.x, rp.dInv.y), iZz = pack2(rp.dInv.z, rp.dInv.z);
    do {
        f32x2_t pLo, pHi, pZ;
        int prim, miss;
        loadNodePairs(base, k, pLo, pHi, pZ, prim, miss, pol);
        if (COUNT) cnt->nodes++;
        const f32x2_t A = mul2(add2(pLo, nOxy), iXy);     // vta.xy
        const f32x2_t B = mul2(add2(pHi, nOxy), iXy);     // vtb.xy
        const f32x2_t C = mul2(add2(pZ, nOzz), iZz);      // {vta.z, vtb.z}
        float ax, ay, az, bx, by, bz;
        unpack2(A, ax, ay); unpack2(B, bx, by); unpack2(C, az, bz);
        float nx, ny, nz, fx, fy, fz, dx, dy;
        if (OCT < 0) {
            nx = fminf(ax, bx); ny = fminf(ay, by); nz = fminf(az, bz);
            fx = fmaxf(ax, bx); fy = fmaxf(ay, by); fz = fmaxf(az, bz);
        } else {
            nx = (OCT & 1) ? bx : ax; fx = (OCT & 1) ? ax : bx;
            ny = (OCT & 2) ? by : ay; fy = (OCT & 2) ? ay : by;
            nz = (OCT & 4) ? bz : az; fz = (OCT & 4) ? az : bz;
        }
        if (OCT >= 0 && (OCT & 3) == 0) unpack2(sub2(B, A), dx, dy);              // vtMax.xy - vtMin.xy
        else if (OCT >= 0 && (OCT & 3) == 3) unpack2(sub2(A, B), dx, dy);
        else if (OCT < 0) unpack2(sub2(pack2(fx, fy), pack2(nx, ny)), dx, dy);
        else { dx = fx - nx; dy = fy - ny; }
        const float dz = fz - nz;
        const float tyz = fz - ny, tzx = fx - nz, txy = fy - nx;
        float szx, syz;
        unpack2(add2(pack2(dx, dy), pack2(dz, dz)), szx, syz);                                       // dz + dx == dt.z + dt.x, dy + dz == dt.y + dt.z
        const float sxy = dx + dy;
        const float tMin = fmaxf(fmaxf(nx, ny), nz), tMax = fminf(fminf(fx, fy), fz);
        const bool hit = (syz > tyz) & (szx > tzx) & (sxy > txy) & (tMax >= 0.0f) & (tMax >= tMin) & !(tMin > dist);
        k = hit ? k + 1 : miss;
        if (hit & (prim >= 0)) {
            const float4* __restrict__ tp = triPos + 3 * (size_t)prim;
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            if (COUNT) cnt->tris++;
            float t;
            if (intersectTriangle(f3(a), f3(b), f3(c), rp.o, rp.d, t) && t < dist) {
                closest = prim;
                if (ANYHIT) k = n;      // any hit ends the walk through the loop condition (no exit from inside the divergent region)
                else dist = t;
            }
        }
    } while (k != n);
    return ANYHIT ? (closest >= 0 ? 1 : 0) : closest;
}
// The scalar form of the same branch-free step (round 1, 46 instructions): plain FADD / FMUL / FMNMX instead of the packed pairs.  Its
// dependent chain per step is a few cycles shorter than the packed one (no pack / unpack moves, lower-latency scalar pipe), which shows
// when the node records are L2-resident and a step is ~250 cycles of memory latency + that chain (Sponza-class: DScene::octantWalk == 2,
// chosen by scene size in zl_scene_create; A/B switch ZL_OCTANT_WALK=2).  Same operations on the same operands: same bits.
template <bool ANYHIT>
ZL_DEV int traversePureScalar(const float4* __restrict__ faceNodes, const float4* __restrict__ triPos, const int n, const RayPrep& rp, float& dist) {
    int closest = -1;
    int k = 0;
    if (n == 0) return ANYHIT ? 0 : closest;
    unsigned long long base = (unsigned long long)faceNodes;
    asm volatile("" : "+l"(base));
    do {
        float4 lo, hi;
        loadNode(reinterpret_cast<const float4*>(base), k, lo, hi);
        const float ax = (lo.x - rp.o.x) * rp.dInv.x, ay = (lo.y - rp.o.y) * rp.dInv.y, az = (lo.z - rp.o.z) * rp.dInv.z;
        const float bx = (hi.x - rp.o.x) * rp.dInv.x, by = (hi.y - rp.o.y) * rp.dInv.y, bz = (hi.z - rp.o.z) * rp.dInv.z;
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
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            float t;
            if (intersectTriangle(f3(a), f3(b), f3(c), rp.o, rp.d, t) && t < dist) {
                closest = prim;
                if (ANYHIT) k = n;
                else dist = t;
            }
        }
    } while (k != n);
    return ANYHIT ? (closest >= 0 ? 1 : 0) : closest;
}
ZL_DEV int rayOctant(float3 d) { return (d.x < 0.0f ? 1 : 0) | (d.y < 0.0f ? 2 : 0) | (d.z < 0.0f ? 4 : 0); }

// ANYHIT = false: bvhHit  -> returns closest primitive id or -1, dist = hit distance or 1e8
// ANYHIT = true : bvhTest -> returns 1 if anything is hit closer than `dist`, else 0
template <bool ANYHIT, bool COUNT>
ZL_DEV int traversePrepared(const float4* __restrict__ nodes, const float4* __restrict__ triPos, const int n, const RayPrep& rp, float& dist, TraceCounters* cnt) {
    if (rp.pure) return traversePure<ANYHIT, COUNT, -1>(nodes, triPos, n, rp, dist, cnt);
    // axis-parallel rays and rays with a near-zero component: the reference's branch order (boxHit)
    int closest = -1;
    int k = 0;
    while (k != n) {
        float4 lo, hi;
        loadNode(nodes, k, lo, hi);
        if (COUNT) cnt->nodes++;
        float boxDist;
        const bool bHit = boxHit<!COUNT>(f3(lo), f3(hi), rp, boxDist);
        if (!bHit || boxDist > dist) { k = __float_as_int(hi.w); continue; }
        const int prim = __float_as_int(lo.w);
        if (prim >= 0) {
            const float4* __restrict__ tp = triPos + 3 * (size_t)prim;
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            if (COUNT) cnt->tris++;
            float t;
            if (intersectTriangle(f3(a), f3(b), f3(c), rp.o, rp.d, t) && t < dist) {
                if (ANYHIT) return 1;
                dist = t;
                closest = prim;
            }
        }
        k++;
    }
    return ANYHIT ? 0 : closest;
}
template <bool ANYHIT, bool COUNT>
ZL_DEV int traverseCore(const float4* __restrict__ allNodes, const float4* __restrict__ triPos, const int n, Ray ray, float& dist, TraceCounters* cnt) {
    const RayPrep rp = prepareRay(ray);
    const float4* __restrict__ nodes = allNodes + (size_t)cubemapFace(-ray.dir) * (size_t)n * 2;
    if (!ANYHIT) dist = 1e8f;
    return traversePrepared<ANYHIT, COUNT>(nodes, triPos, n, rp, dist, cnt);
}

// ---------------------------------------------------------------------------------------------------------------------
// The same walk over "child boxes in the parent" records (BVH2 layout) with a short per-lane stack (ZL_BVH2_WALK, DScene::bvh2).
//
// The threaded walk fetches one 32-byte record per VISITED node — about 48 per ray in the Rungholt-class scene, half of which are
// visited only to be rejected — and every fetch depends on the one before: the kernel waits on that chain (profiles/r2_trace_sweep.md).
// The threaded order of face f is nothing but a depth-first order whose child order at every interior node is fixed by comparing the
// children's centroids along the face's axis (BVH::buildHitTable, BVH.cpp:298-346).  So the reference's visit sequence can be
// produced from ONE copy of the builder's tree: record r of interior node N = {box(L), box(R), ref(L), ref(R), 6 order bits},
// 64 bytes (two sectors), L / R the builder's left / right child, ref >= 0 an interior record, ref < 0 the leaf ~ref, bit f set
// iff L comes first in face f (the strict comparison of BVH.cpp:339-340: ties put R first).  At an accepted interior node the walk
// tests BOTH child boxes from the one record (boxHit is a function of ray and box only), remembers the second child with its tMin
// on a per-lane stack if its box is hit, and goes on with the first child exactly as the reference does: enter it iff its box is hit
// and !(tMin > dist), with the dist of that moment; the remembered child is re-checked against the dist of ITS moment when popped —
// the same two conditions `!bHit || boxDist > dist` the reference evaluates on arrival (intersection.glsl:409).  Leaves have no
// record: their box sits in the parent and their triangle is fetched only when that test passes.  Same nodes, same order, same
// tests, same strict `t < dist` updates => same ids and distances, bit for bit (tests: every traversal test runs both walks).
// Per ray this fetches one record per ENTERED interior node (about 20) instead of one per visited node, from 0.4 GB instead of
// 2.4 GB of records (one copy instead of six orderings).
static constexpr int kBvh2Stack = 64;           // entries per lane; scenes whose tree is deeper keep the threaded walk (DScene::bvh2 == nullptr)
struct Bvh2Child { f32x2_t lo, hi, z; int ref; };
// ray-invariant part of boxHitPure for one child: returns hit (without the dist cull) and tMin
template <int OCT>
ZL_DEV bool slabTestPure(const f32x2_t pLo, const f32x2_t pHi, const f32x2_t pZ, const f32x2_t nOxy, const f32x2_t nOzz, const f32x2_t iXy, const f32x2_t iZz, float& tMinOut) {
    const f32x2_t A = mul2(add2(pLo, nOxy), iXy), B = mul2(add2(pHi, nOxy), iXy), C = mul2(add2(pZ, nOzz), iZz);
    float ax, ay, az, bx, by, bz;
    unpack2(A, ax, ay); unpack2(B, bx, by); unpack2(C, az, bz);
    float nx, ny, nz, fx, fy, fz, dx, dy;
    if (OCT < 0) {
        nx = fminf(ax, bx); ny = fminf(ay, by); nz = fminf(az, bz);
        fx = fmaxf(ax, bx); fy = fmaxf(ay, by); fz = fmaxf(az, bz);
    } else {
        nx = (OCT & 1) ? bx : ax; fx = (OCT & 1) ? ax : bx;
        ny = (OCT & 2) ? by : ay; fy = (OCT & 2) ? ay : by;
        nz = (OCT & 4) ? bz : az; fz = (OCT & 4) ? az : bz;
    }
    if (OCT >= 0 && (OCT & 3) == 0) unpack2(sub2(B, A), dx, dy);
    else if (OCT >= 0 && (OCT & 3) == 3) unpack2(sub2(A, B), dx, dy);
    else if (OCT < 0) unpack2(sub2(pack2(fx, fy), pack2(nx, ny)), dx, dy);
    else { dx = fx - nx; dy = fy - ny; }
    const float dz = fz - nz;
    const float tyz = fz - ny, tzx = fx - nz, txy = fy - nx;
    float szx, syz;
    unpack2(add2(pack2(dx, dy), pack2(dz, dz)), szx, syz);
    const float sxy = dx + dy;
    const float tMin = fmaxf(fmaxf(nx, ny), nz), tMax = fminf(fminf(fx, fy), fz);
    tMinOut = tMin;
    return (syz > tyz) & (szx > tzx) & (sxy > txy) & (tMax >= 0.0f) & (tMax >= tMin);
}
template <bool ANYHIT, int OCT>
ZL_DEV int traverseBvh2(const DScene& S, const RayPrep& rp, const int face, float& dist) {
    const f32x2_t nOxy = pack2(-rp.o.x, -rp.o.y), nOzz = pack2(-rp.o.z, -rp.o.z);
    const f32x2_t iXy = pack2(rp.dInv.x, rp.dInv.y), iZz = pack2(rp.dInv.z, rp.dInv.z);
    int closest = -1;
    {   // the root's own box (the reference's first step)
        float t0;
        const bool h = slabTestPure<OCT>(pack2(S.rootLo.x, S.rootLo.y), pack2(S.rootHi.x, S.rootHi.y), pack2(S.rootLo.z, S.rootHi.z), nOxy, nOzz, iXy, iZz, t0);
        if (!h || t0 > dist) return ANYHIT ? 0 : -1;
    }
    int2 stack[kBvh2Stack];
    int sp = 0;
    int cur = 0;                                   // >= 0: interior record to open; < 0: the leaf ~cur whose box test passed; 0x7fffffff: pop
    bool done = false;
    unsigned long long base = (unsigned long long)S.bvh2;
    asm volatile("" : "+l"(base));
    const float4* __restrict__ triPos = S.triPos;
    // (two restructurings of this loop were measured and were slower: one iteration = "[pop one] [open one] [test one triangle]" with
    //  every part predicated, 7.09 against 6.08 ms on the Rungholt-class pass, and a `while (true)` form of this same logic, 8.45 ms —
    //  profiles/r2_trace_sweep.md)
    while (!done) {
        if (cur >= 0) {
            // open interior record `cur`: both child boxes in one 64-byte fetch
            f32x2_t aLo, aHi, aZ, bLo, bHi, bZ, l0, l1;
            const unsigned long long addr = base + 64ull * (unsigned long long)cur;
            asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(aLo), "=l"(aHi), "=l"(aZ), "=l"(l0) : "l"(addr));
            asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(bLo), "=l"(bHi), "=l"(bZ), "=l"(l1) : "l"(addr + 32ull));
            int refA, refB, bits, pad;
            asm("mov.b64 {%0, %1}, %2;" : "=r"(refA), "=r"(refB) : "l"(l0));
            asm("mov.b64 {%0, %1}, %2;" : "=r"(bits), "=r"(pad) : "l"(l1));
            float tA, tB;
            const bool hA = slabTestPure<OCT>(aLo, aHi, aZ, nOxy, nOzz, iXy, iZz, tA);
            const bool hB = slabTestPure<OCT>(bLo, bHi, bZ, nOxy, nOzz, iXy, iZz, tB);
            const bool aFirst = (bits >> face) & 1;
            const int ref1 = aFirst ? refA : refB, ref2 = aFirst ? refB : refA;
            const float t1 = aFirst ? tA : tB, t2 = aFirst ? tB : tA;
            const bool h1 = aFirst ? hA : hB, h2 = aFirst ? hB : hA;
            if (h2) { stack[sp] = make_int2(ref2, __float_as_int(t2)); sp++; }
            if (h1 & !(t1 > dist)) { cur = ref1; if (cur >= 0) continue; }
            else cur = 0x7fffffff;                 // nothing to enter: pop
        }
        if (cur < 0) {                             // a leaf whose box test passed: the triangle test of the reference's leaf step
            const int prim = ~cur;
            const float4* __restrict__ tp = triPos + 3 * (size_t)prim;
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            float t;
            if (intersectTriangle(f3(a), f3(b), f3(c), rp.o, rp.d, t) && t < dist) {
                closest = prim;
                if (ANYHIT) return 1;
                dist = t;
            }
        }
        // pop the next remembered child that survives the cull (intersection.glsl:409) with the dist of this moment
        cur = 0x7fffffff;
        while (sp > 0) {
            sp--;
            const int2 e = stack[sp];
            if (!(__int_as_float(e.y) > dist)) { cur = e.x; break; }
        }
        if (cur == 0x7fffffff) done = true;
    }
    return ANYHIT ? 0 : closest;
}

// The queue / ray-set kernels' entry: when every converged lane of the warp holds a pure ray of ONE direction octant
// (sorted queues, camera tiles: nearly always), the warp takes that octant's specialised walk (no per-axis min / max);
// a mixed warp takes the general walk, which keeps its lanes in lock step whatever their octants.  Same results either way.
// LEAN > 0: the instantiations the default configuration runs — the A/B walks that are off by default (BVH2 records + stack, eviction
// priorities on the node loads) are compiled out and the scene's choice of step (1 = packed octant walks, 2 = scalar walk) is fixed,
// so the other walks' stack frames, code and register pressure do not ride along in the kernel (704 -> 104 bytes of stack frame).
template <bool ANYHIT, int LEAN = 0>
ZL_DEV int traverseWarp(const DScene& S, Ray ray, float& dist) {
    const RayPrep rp = prepareRay(ray);
    const int n = S.bvhSize;
    const float4* __restrict__ nodes = S.nodes + (size_t)cubemapFace(-ray.dir) * (size_t)n * 2;
    if (!ANYHIT) dist = 1e8f;
    const float4* const bvh2 = LEAN ? nullptr : S.bvh2;
    const int nodePolicy = LEAN ? 0 : S.nodePolicy;
    const int octantWalk = LEAN ? LEAN : S.octantWalk;
    if (octantWalk == 2 && rp.pure && bvh2 == nullptr) return traversePureScalar<ANYHIT>(nodes, S.triPos, n, rp, dist);
    const int oct = rp.pure ? rayOctant(ray.dir) : 8;
    int uniform = 0;
    if (octantWalk) __match_all_sync(__activemask(), oct, &uniform);
    if (!LEAN && bvh2 != nullptr && rp.pure) {            // child-boxes-in-the-parent records + short stack: same visit sequence, fewer dependent fetches
        const int face = cubemapFace(-ray.dir);
        if (uniform && oct < 8) {
            switch (oct) {
            case 0: return traverseBvh2<ANYHIT, 0>(S, rp, face, dist);
            case 1: return traverseBvh2<ANYHIT, 1>(S, rp, face, dist);
            case 2: return traverseBvh2<ANYHIT, 2>(S, rp, face, dist);
            case 3: return traverseBvh2<ANYHIT, 3>(S, rp, face, dist);
            case 4: return traverseBvh2<ANYHIT, 4>(S, rp, face, dist);
            case 5: return traverseBvh2<ANYHIT, 5>(S, rp, face, dist);
            case 6: return traverseBvh2<ANYHIT, 6>(S, rp, face, dist);
            default: return traverseBvh2<ANYHIT, 7>(S, rp, face, dist);
            }
        }
        return traverseBvh2<ANYHIT, -1>(S, rp, face, dist);
    }
    if (LEAN != 2 && uniform && oct < 8) {
        switch (oct) {
        case 0: return traversePure<ANYHIT, false, 0>(nodes, S.triPos, n, rp, dist, nullptr, nodePolicy);
        case 1: return traversePure<ANYHIT, false, 1>(nodes, S.triPos, n, rp, dist, nullptr, nodePolicy);
        case 2: return traversePure<ANYHIT, false, 2>(nodes, S.triPos, n, rp, dist, nullptr, nodePolicy);
        case 3: return traversePure<ANYHIT, false, 3>(nodes, S.triPos, n, rp, dist, nullptr, nodePolicy);
        case 4: return traversePure<ANYHIT, false, 4>(nodes, S.triPos, n, rp, dist, nullptr, nodePolicy);
        case 5: return traversePure<ANYHIT, false, 5>(nodes, S.triPos, n, rp, dist, nullptr, nodePolicy);
        case 6: return traversePure<ANYHIT, false, 6>(nodes, S.triPos, n, rp, dist, nullptr, nodePolicy);
        default: return traversePure<ANYHIT, false, 7>(nodes, S.triPos, n, rp, dist, nullptr, nodePolicy);
        }
    }
    if (rp.pure) return traversePure<ANYHIT, false, -1>(nodes, S.triPos, n, rp, dist, nullptr, nodePolicy);
    return traversePrepared<ANYHIT, false>(nodes, S.triPos, n, rp, dist, nullptr);
}

// ---------------------------------------------------------------------------------------------------------------------
// Walk with INTRA-WARP RAY COMPACTION (north_star "ballot/shuffle ray compaction"; A/B switch ZL_WF_TRACE_LOOP=7).
//
// ncu on the plain walk: the L1 data stage is the busiest unit (55 % of peak on the Rungholt-class pass, 74 % Sponza-class,
// profiles/r2_ncu_trace_*.csv) while 12-14 of 32 lanes are live — a 256-bit warp load is processed a quarter-warp at a time and
// costs its wavefronts per quarter that has ANY live lane.  Rays finish at different steps, so the live lanes end up scattered
// over all four quarters.  Here the warp votes every kCompactEvery steps and, when the live rays would fit into fewer
// quarter-warps than they occupy, shuffles them down into the lowest lanes (14 registers per ray: origin, direction,
// reciprocal direction, dist, k, closest, face, home lane).  No new rays are brought in — the warp's rays stay the coherent
// set the sort made them — so, unlike the regenerating / refill kernels of round 1, lanes do not start touching unrelated
// sectors.  A finished ray leaves its result in a 32-entry per-warp shared-memory table indexed by its home lane; every lane
// picks its own up at the end.  Per ray the sequence of box tests, triangle tests and dist updates is traversePure's.
static constexpr int kCompactEvery = 4;
template <bool ANYHIT, int OCT>
ZL_DEV int traversePureCompact(const float4* __restrict__ allNodes, const float4* __restrict__ triPos, const int n, RayPrep rp, int face, float& dist,
                               int2* __restrict__ warpRes /* 32 entries of this warp, shared memory */, const int pol) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int closest = -1, k = 0, home = lane;
    bool alive = n > 0;
    if (!alive) warpRes[lane] = make_int2(-1, __float_as_int(dist));
    f32x2_t nOxy = pack2(-rp.o.x, -rp.o.y), nOzz = pack2(-rp.o.z, -rp.o.z);
    f32x2_t iXy = pack2(rp.dInv.x, rp.dInv.y), iZz = pack2(rp.dInv.z, rp.dInv.z);
    unsigned long long base = (unsigned long long)(allNodes + (size_t)face * (size_t)n * 2);
    for (int iter = 0;; iter++) {
        if ((iter & (kCompactEvery - 1)) == 0) {
            const unsigned live = __ballot_sync(FULL, alive);
            if (!live) break;
            const int nLive = __popc(live);
            const int quartersUsed = ((live & 0xffu) != 0) + ((live & 0xff00u) != 0) + ((live & 0xff0000u) != 0) + ((live & 0xff000000u) != 0);
            if (((nLive + 7) >> 3) < quartersUsed) {
                // lane d < nLive takes over the d-th live ray (in lane order); the others fall idle
                const int src = (lane < nLive) ? (int)__fns(live, 0, lane + 1) : lane;
                rp.o.x = __shfl_sync(FULL, rp.o.x, src); rp.o.y = __shfl_sync(FULL, rp.o.y, src); rp.o.z = __shfl_sync(FULL, rp.o.z, src);
                rp.d.x = __shfl_sync(FULL, rp.d.x, src); rp.d.y = __shfl_sync(FULL, rp.d.y, src); rp.d.z = __shfl_sync(FULL, rp.d.z, src);
                rp.dInv.x = __shfl_sync(FULL, rp.dInv.x, src); rp.dInv.y = __shfl_sync(FULL, rp.dInv.y, src); rp.dInv.z = __shfl_sync(FULL, rp.dInv.z, src);
                dist = __shfl_sync(FULL, dist, src); k = __shfl_sync(FULL, k, src); closest = __shfl_sync(FULL, closest, src);
                face = __shfl_sync(FULL, face, src); home = __shfl_sync(FULL, home, src);
                alive = lane < nLive;
                nOxy = pack2(-rp.o.x, -rp.o.y); nOzz = pack2(-rp.o.z, -rp.o.z);
                iXy = pack2(rp.dInv.x, rp.dInv.y); iZz = pack2(rp.dInv.z, rp.dInv.z);
                base = (unsigned long long)(allNodes + (size_t)face * (size_t)n * 2);
            }
        }
        if (alive) {
            f32x2_t pLo, pHi, pZ;
            int prim, miss;
            loadNodePairs(base, k, pLo, pHi, pZ, prim, miss, pol);
            const f32x2_t A = mul2(add2(pLo, nOxy), iXy), B = mul2(add2(pHi, nOxy), iXy), C = mul2(add2(pZ, nOzz), iZz);
            float ax, ay, az, bx, by, bz;
            unpack2(A, ax, ay); unpack2(B, bx, by); unpack2(C, az, bz);
            float nx, ny, nz, fx, fy, fz, dx, dy;
            if (OCT < 0) {
                nx = fminf(ax, bx); ny = fminf(ay, by); nz = fminf(az, bz);
                fx = fmaxf(ax, bx); fy = fmaxf(ay, by); fz = fmaxf(az, bz);
            } else {
                nx = (OCT & 1) ? bx : ax; fx = (OCT & 1) ? ax : bx;
                ny = (OCT & 2) ? by : ay; fy = (OCT & 2) ? ay : by;
                nz = (OCT & 4) ? bz : az; fz = (OCT & 4) ? az : bz;
            }
            if (OCT >= 0 && (OCT & 3) == 0) unpack2(sub2(B, A), dx, dy);
            else if (OCT >= 0 && (OCT & 3) == 3) unpack2(sub2(A, B), dx, dy);
            else if (OCT < 0) unpack2(sub2(pack2(fx, fy), pack2(nx, ny)), dx, dy);
            else { dx = fx - nx; dy = fy - ny; }
            const float dz = fz - nz;
            const float tyz = fz - ny, tzx = fx - nz, txy = fy - nx;
            float szx, syz;
            unpack2(add2(pack2(dx, dy), pack2(dz, dz)), szx, syz);
            const float sxy = dx + dy;
            const float tMin = fmaxf(fmaxf(nx, ny), nz), tMax = fminf(fminf(fx, fy), fz);
            const bool hit = (syz > tyz) & (szx > tzx) & (sxy > txy) & (tMax >= 0.0f) & (tMax >= tMin) & !(tMin > dist);
            k = hit ? k + 1 : miss;
            if (hit & (prim >= 0)) {
                const float4* __restrict__ tp = triPos + 3 * (size_t)prim;
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                float t;
                if (intersectTriangle(f3(a), f3(b), f3(c), rp.o, rp.d, t) && t < dist) {
                    closest = prim;
                    if (ANYHIT) k = n;
                    else dist = t;
                }
            }
            if (k == n) { alive = false; warpRes[home] = make_int2(closest, __float_as_int(dist)); }
        }
    }
    __syncwarp(FULL);
    const int2 r = warpRes[lane];
    __syncwarp(FULL);
    dist = __int_as_float(r.y);
    return ANYHIT ? (r.x >= 0 ? 1 : 0) : r.x;
}
// Entry for a FULLY converged warp whose 32 lanes all hold a pure ray of one octant and one kind (any-hit / closest-hit); the caller
// falls back to traverseWarp otherwise.
template <bool ANYHIT>
ZL_DEV int traverseWarpCompact(const DScene& S, Ray ray, float& dist, const int oct, int2* __restrict__ warpRes) {
    const RayPrep rp = prepareRay(ray);
    const int face = cubemapFace(-ray.dir);
    if (!ANYHIT) dist = 1e8f;
    switch (oct) {
    case 0: return traversePureCompact<ANYHIT, 0>(S.nodes, S.triPos, S.bvhSize, rp, face, dist, warpRes, S.nodePolicy);
    case 1: return traversePureCompact<ANYHIT, 1>(S.nodes, S.triPos, S.bvhSize, rp, face, dist, warpRes, S.nodePolicy);
    case 2: return traversePureCompact<ANYHIT, 2>(S.nodes, S.triPos, S.bvhSize, rp, face, dist, warpRes, S.nodePolicy);
    case 3: return traversePureCompact<ANYHIT, 3>(S.nodes, S.triPos, S.bvhSize, rp, face, dist, warpRes, S.nodePolicy);
    case 4: return traversePureCompact<ANYHIT, 4>(S.nodes, S.triPos, S.bvhSize, rp, face, dist, warpRes, S.nodePolicy);
    case 5: return traversePureCompact<ANYHIT, 5>(S.nodes, S.triPos, S.bvhSize, rp, face, dist, warpRes, S.nodePolicy);
    case 6: return traversePureCompact<ANYHIT, 6>(S.nodes, S.triPos, S.bvhSize, rp, face, dist, warpRes, S.nodePolicy);
    default: return traversePureCompact<ANYHIT, 7>(S.nodes, S.triPos, S.bvhSize, rp, face, dist, warpRes, S.nodePolicy);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Shared-memory staging of the top BVH levels (north_star; A/B switch ZL_WF_TRACE_LOOP=6, wfTraceStagedKernel).
//
// The threaded orderings are pre-order, so the top D levels are not a prefix of the node array.  buildStagedTopKernel
// copies, per face, the nodes of depth < D into a compact array IN THEIR OWN PRE-ORDER, 48 bytes each:
//   {pMin.x, pMin.y, pMax.x, pMax.y} {pMin.z, pMax.z, bits(prim | -1), bits(missRef)} {bits(hitRef), -, -, -}
// with explicit links: ref >= 0 is an index into the face's global records (bvhSize = end of the walk), ref < 0 is staged entry
// -ref - 1.  The miss link of a node leads to a node of the same depth or higher up, so staged miss links stay inside the staged
// array; only the hit link of a depth D-1 node leaves it.  A walk starts in the staged copy (entry 0 = the root), follows it until a
// hit link hands it a global index, and finishes in traversePure from there.  Links of the GLOBAL records are not rewritten, so
// later returns to top-level nodes (miss links out of deep subtrees) read the global records as before.  The visit sequence, and
// with it every result, is the reference's.  The staged array is brought into shared memory by one TMA bulk copy per CTA.
static constexpr int kTopDepth = 7;                          // levels staged per face
static constexpr int kTopNodes = (1 << kTopDepth) - 1;       // entries reserved per face (a face may use fewer: leaves above depth D)
static constexpr int kTopVecs = 3;                           // float4 per staged entry
static constexpr int kTopBytes = 6 * kTopNodes * kTopVecs * 16;

__global__ void buildStagedTopKernel(const float4* __restrict__ nodes, const int n, float4* __restrict__ top) {
    const int face = blockIdx.x;
    if (threadIdx.x != 0) return;
    const float4* fn = nodes + (size_t)face * (size_t)n * 2;
    float4* out = top + (size_t)face * kTopNodes * kTopVecs;
    int gIdx[kTopNodes], gDepth[kTopNodes];
    int stackG[2 * kTopDepth + 2], stackD[2 * kTopDepth + 2];
    int count = 0, sp = 0;
    if (n > 0) { stackG[sp] = 0; stackD[sp] = 0; sp++; }
    while (sp) {                                             // pre-order over the nodes of depth < kTopDepth
        sp--;
        const int g = stackG[sp], d = stackD[sp];
        gIdx[count] = g; gDepth[count] = d; count++;
        const float4 r1 = fn[2 * (size_t)g + 1];
        const int prim = __float_as_int(r1.z);
        if (prim < 0 && d + 1 < kTopDepth) {
            const int first = g + 1;
            const int second = __float_as_int(fn[2 * (size_t)first + 1].w);       // miss link of the first child = the second child
            stackG[sp] = second; stackD[sp] = d + 1; sp++;
            stackG[sp] = first; stackD[sp] = d + 1; sp++;
        }
    }
    for (int s = 0; s < kTopNodes; s++) {
        float4 r0 = make_float4(0, 0, 0, 0), r1 = r0, r2 = r0;
        if (s < count) {
            const int g = gIdx[s];
            r0 = fn[2 * (size_t)g]; r1 = fn[2 * (size_t)g + 1];
            const int prim = __float_as_int(r1.z), miss = __float_as_int(r1.w);
            int missRef = miss;                              // == n: end of the walk
            if (miss != n) {
                missRef = miss;                              // (a target below the staged levels cannot occur; keep the global index if it did)
                for (int t = s + 1; t < count; t++) if (gIdx[t] == miss) { missRef = -t - 1; break; }
            }
            int hitRef;
            if (prim >= 0) hitRef = missRef;                 // a leaf: k + 1 is its miss target
            else if (gDepth[s] + 1 < kTopDepth) hitRef = -(s + 1) - 1;            // first child = next staged entry
            else hitRef = g + 1;                             // leaves the staged levels
            r1.w = __int_as_float(missRef);
            r2.x = __int_as_float(hitRef);
        }
        out[s * kTopVecs] = r0; out[s * kTopVecs + 1] = r1; out[s * kTopVecs + 2] = r2;
    }
}

// The staged part of a pure ray's walk: scalar restatement of the box step (same IEEE operations as traversePure, so the same
// decisions).  Returns the global record index to continue from (n = finished); closest / dist carry over.
template <bool ANYHIT>
ZL_DEV int traverseStagedTop(const float4* __restrict__ topFace /* shared */, const float4* __restrict__ triPos, const int n, const RayPrep& rp, float& dist, int& closest) {
    int ref = n > 0 ? -1 : n;
    while (ref < 0) {
        const float4* e = topFace + (-ref - 1) * kTopVecs;
        const float4 r0 = e[0], r1 = e[1];
        const int prim = __float_as_int(r1.z);
        float tMin;
        const bool bHit = boxHitPure(f3(r0.x, r0.y, r1.x), f3(r0.z, r0.w, r1.y), rp, tMin);
        if (!bHit || tMin > dist) { ref = __float_as_int(r1.w); continue; }
        ref = __float_as_int(e[2].x);
        if (prim >= 0) {
            const float4* __restrict__ tp = triPos + 3 * (size_t)prim;
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            float t;
            if (intersectTriangle(f3(a), f3(b), f3(c), rp.o, rp.d, t) && t < dist) {
                closest = prim;
                if (ANYHIT) return n;
                dist = t;
            }
        }
    }
    return ref;
}
template <bool ANYHIT>
ZL_DEV int traverseWarpStaged(const DScene& S, const float4* __restrict__ topShared, Ray ray, float& dist) {
    const RayPrep rp = prepareRay(ray);
    const int n = S.bvhSize;
    const int face = cubemapFace(-ray.dir);
    const float4* __restrict__ nodes = S.nodes + (size_t)face * (size_t)n * 2;
    if (!ANYHIT) dist = 1e8f;
    if (!rp.pure) return traversePrepared<ANYHIT, false>(nodes, S.triPos, n, rp, dist, nullptr);
    int closest = -1;
    const int k = traverseStagedTop<ANYHIT>(topShared + face * kTopNodes * kTopVecs, S.triPos, n, rp, dist, closest);
    return traversePure<ANYHIT, false, -1>(nodes, S.triPos, n, rp, dist, nullptr, S.nodePolicy, k, closest);
}

// Two rays per lane (wfTraceDualKernel).  The queue trace kernel is bound by the latency of one node load per warp-step
// with one load in flight per warp; here every lane walks TWO pure rays in the same loop, both node records
// requested before either is tested, so a warp has two loads in flight.  Per ray the sequence of box tests, triangle tests and
// distance updates is exactly traversePure<..., OCT>'s (same operations on the same operands), so results are bit-identical.
// k and end are ABSOLUTE record indices (face offset included: the two rays of a lane may use different faces); the miss
// links stored in the records are relative to the face.
struct WalkRay {
    float3 o, d, dInv;
    float dist;         // in: search limit (1e8 for closest-hit); out: hit distance (closest-hit)
    int k, end;         // next record / one past the face's last record; k == end: finished (or not taking part)
    int closest;        // -1, or the last accepted primitive
    bool anyhit;
};
template <int OCT>
ZL_DEV void walkStep(WalkRay& r, const bool active, const int n, const f32x2_t pLo, const f32x2_t pHi, const f32x2_t pZ, const int prim, const int miss, const float4* __restrict__ triPos) {
    const f32x2_t nOxy = pack2(-r.o.x, -r.o.y), nOzz = pack2(-r.o.z, -r.o.z);
    const f32x2_t iXy = pack2(r.dInv.x, r.dInv.y), iZz = pack2(r.dInv.z, r.dInv.z);
    const f32x2_t A = mul2(add2(pLo, nOxy), iXy), B = mul2(add2(pHi, nOxy), iXy), C = mul2(add2(pZ, nOzz), iZz);
    float ax, ay, az, bx, by, bz;
    unpack2(A, ax, ay); unpack2(B, bx, by); unpack2(C, az, bz);
    float nx, ny, nz, fx, fy, fz, dx, dy;
    if (OCT < 0) {
        nx = fminf(ax, bx); ny = fminf(ay, by); nz = fminf(az, bz);
        fx = fmaxf(ax, bx); fy = fmaxf(ay, by); fz = fmaxf(az, bz);
    } else {
        nx = (OCT & 1) ? bx : ax; fx = (OCT & 1) ? ax : bx;
        ny = (OCT & 2) ? by : ay; fy = (OCT & 2) ? ay : by;
        nz = (OCT & 4) ? bz : az; fz = (OCT & 4) ? az : bz;
    }
    if (OCT >= 0 && (OCT & 3) == 0) unpack2(sub2(B, A), dx, dy);
    else if (OCT >= 0 && (OCT & 3) == 3) unpack2(sub2(A, B), dx, dy);
    else if (OCT < 0) unpack2(sub2(pack2(fx, fy), pack2(nx, ny)), dx, dy);
    else { dx = fx - nx; dy = fy - ny; }
    const float dz = fz - nz;
    const float tyz = fz - ny, tzx = fx - nz, txy = fy - nx;
    float szx, syz;
    unpack2(add2(pack2(dx, dy), pack2(dz, dz)), szx, syz);
    const float sxy = dx + dy;
    const float tMin = fmaxf(fmaxf(nx, ny), nz), tMax = fminf(fminf(fx, fy), fz);
    const bool hit = active & (syz > tyz) & (szx > tzx) & (sxy > txy) & (tMax >= 0.0f) & (tMax >= tMin) & !(tMin > r.dist);
    r.k = hit ? r.k + 1 : (active ? miss + (r.end - n) : r.k);
    if (hit & (prim >= 0)) {
        const float4* __restrict__ tp = triPos + 3 * (size_t)prim;
        const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
        float t;
        if (intersectTriangle(f3(a), f3(b), f3(c), r.o, r.d, t) && t < r.dist) {
            r.closest = prim;
            if (r.anyhit) r.k = r.end;
            else r.dist = t;
        }
    }
}
template <int OCT>
ZL_DEV void traverseDual(const float4* __restrict__ allNodes, const float4* __restrict__ triPos, const int n, WalkRay& ray0, WalkRay& ray1) {
    WalkRay r0 = ray0, r1 = ray1;       // working copies in registers
    unsigned long long base = (unsigned long long)allNodes;
    asm volatile("" : "+l"(base));
    while ((r0.k != r0.end) | (r1.k != r1.end)) {
        // both records are requested before either is tested; a finished ray re-reads the record at its end index
        // (the next face's first record, or the pad record behind the last face) and its step is a no-op
        f32x2_t lo0, hi0, z0, lo1, hi1, z1;
        int prim0, miss0, prim1, miss1;
        loadNodePairs(base, r0.k, lo0, hi0, z0, prim0, miss0);
        loadNodePairs(base, r1.k, lo1, hi1, z1, prim1, miss1);
        walkStep<OCT>(r0, r0.k != r0.end, n, lo0, hi0, z0, prim0, miss0, triPos);
        walkStep<OCT>(r1, r1.k != r1.end, n, lo1, hi1, z1, prim1, miss1, triPos);
    }
    ray0.dist = r0.dist; ray0.closest = r0.closest;
    ray1.dist = r1.dist; ray1.closest = r1.closest;
}
// oct: 0..7 when every ray that takes part has that octant (the caller voted), -1 otherwise
ZL_DEV void traverseDualDispatch(const float4* __restrict__ allNodes, const float4* __restrict__ triPos, const int n, WalkRay& r0, WalkRay& r1, const int oct) {
    if (oct == 0) traverseDual<0>(allNodes, triPos, n, r0, r1);
    else if (oct == 1) traverseDual<1>(allNodes, triPos, n, r0, r1);
    else if (oct == 2) traverseDual<2>(allNodes, triPos, n, r0, r1);
    else if (oct == 3) traverseDual<3>(allNodes, triPos, n, r0, r1);
    else if (oct == 4) traverseDual<4>(allNodes, triPos, n, r0, r1);
    else if (oct == 5) traverseDual<5>(allNodes, triPos, n, r0, r1);
    else if (oct == 6) traverseDual<6>(allNodes, triPos, n, r0, r1);
    else if (oct == 7) traverseDual<7>(allNodes, triPos, n, r0, r1);
    else traverseDual<-1>(allNodes, triPos, n, r0, r1);
}

// Same walk with the hit link requested one step ahead.  The hit link of threaded entry k is the next
// record in memory, so its address is known before record k has been tested: both loads are in
// flight together and a "hit" step (about half of all steps) finds its record already in registers;
// only miss links cost a dependent load.  The visit sequence, and with it every result, is unchanged;
// a record fetched for a step that turns out to be a miss is dropped (it shares the 128-byte line of
// its predecessor three times out of four).  The node array carries one pad record so that k + 1 == n
// may be read.
template <bool ANYHIT>
ZL_DEV int traverseSpec(const float4* __restrict__ allNodes, const float4* __restrict__ triPos, const int n, Ray ray, float& dist) {
    const RayPrep rp = prepareRay(ray);
    const float4* __restrict__ nodes = allNodes + (size_t)cubemapFace(-ray.dir) * (size_t)n * 2;
    if (!ANYHIT) dist = 1e8f;
    int closest = -1;
    if (n == 0) return ANYHIT ? 0 : closest;
    int k = 0;
    float4 lo, hi, nlo, nhi;
    loadNode(nodes, 0, lo, hi);
    while (true) {
        loadNode(nodes, k + 1, nlo, nhi);
        float boxDist;
        const bool bHit = rp.pure ? boxHitPure(f3(lo), f3(hi), rp, boxDist) : boxHit<true>(f3(lo), f3(hi), rp, boxDist);
        if (!bHit || boxDist > dist) {
            k = __float_as_int(hi.w);
            if (k == n) break;
            loadNode(nodes, k, lo, hi);
            continue;
        }
        const int prim = __float_as_int(lo.w);
        if (prim >= 0) {
            const float4* __restrict__ tp = triPos + 3 * (size_t)prim;
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            float t;
            if (intersectTriangle(f3(a), f3(b), f3(c), rp.o, rp.d, t) && t < dist) {
                if (ANYHIT) return 1;
                dist = t;
                closest = prim;
            }
        }
        k++;
        if (k == n) break;
        lo = nlo; hi = nhi;
    }
    return ANYHIT ? 0 : closest;
}

// inlined form (the dedicated traversal kernels)
template <bool ANYHIT, bool COUNT>
ZL_DEV int traverse(const DScene& S, Ray ray, float& dist, TraceCounters* cnt) {
    return traverseCore<ANYHIT, COUNT>(S.nodes, S.triPos, S.bvhSize, ray, dist, cnt);
}
// out-of-line form (the integrator kernels call it from several places): x = id, y = bits(dist)
struct HitResult { int id; float dist; };
template <bool COUNT>
ZL_CALL HitResult bvhHitCall(const float4* __restrict__ nodes, const float4* __restrict__ triPos, int n, float3 o, float3 d, TraceCounters* cnt) {
    HitResult r;
    r.id = traverseCore<false, COUNT>(nodes, triPos, n, makeRay(o, d), r.dist, cnt);
    return r;
}
template <bool COUNT>
ZL_CALL int bvhTestCall(const float4* __restrict__ nodes, const float4* __restrict__ triPos, int n, float3 o, float3 d, float dist, TraceCounters* cnt) {
    return traverseCore<true, COUNT>(nodes, triPos, n, makeRay(o, d), dist, cnt);
}

#ifndef ZL_INSTRUMENT
ZL_DEV int bvhHit(const DScene& S, Ray ray, float& dist) {                              // :395-427
    HitResult r = bvhHitCall<false>(S.nodes, S.triPos, S.bvhSize, ray.ori, ray.dir, nullptr);
    dist = r.dist;
    return r.id;
}
ZL_DEV bool bvhTest(const DScene& S, Ray ray, float dist) {                             // :367-393
    return bvhTestCall<false>(S.nodes, S.triPos, S.bvhSize, ray.ori, ray.dir, dist, nullptr) != 0;
}
ZL_DEV void countEvent(const DScene&, int) {}
#else
// Instrumented build (zl_instrumented.cu): same code, plus the bvhDebug-style visit counters of
// intersection.glsl:331-365 summed into S.counters = {rays, nodes, tris, shades, splats, paths}.
ZL_DEV void countEvent(const DScene& S, int slot) { if (S.counters) atomicAdd(S.counters + slot, 1ull); }
ZL_DEV void countRay(const DScene& S, const TraceCounters& c) {
    if (!S.counters) return;
    atomicAdd(S.counters + 0, 1ull);
    atomicAdd(S.counters + 1, (unsigned long long)c.nodes);
    atomicAdd(S.counters + 2, (unsigned long long)c.tris);
}
ZL_DEV int bvhHit(const DScene& S, Ray ray, float& dist) {
    TraceCounters c{0, 0};
    HitResult r = bvhHitCall<true>(S.nodes, S.triPos, S.bvhSize, ray.ori, ray.dir, &c);
    dist = r.dist;
    countRay(S, c);
    return r.id;
}
ZL_DEV bool bvhTest(const DScene& S, Ray ray, float dist) {
    TraceCounters c{0, 0};
    bool hit = bvhTestCall<true>(S.nodes, S.triPos, S.bvhSize, ray.ori, ray.dir, dist, &c) != 0;
    countRay(S, c);
    return hit;
}
#endif
ZL_DEV bool visible(const DScene& S, float3 x, float3 y) {                               // :429-434
    float dist = distance(x, y) - 2e-5f;
    float3 wi = normalize(y - x);
    return !bvhTest(S, makeRay(x + wi * 1e-5f, wi), dist);
}

}  // namespace zl
