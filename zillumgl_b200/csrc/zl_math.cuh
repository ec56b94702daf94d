// zl_math.cuh — device vector math with the numeric contract of the hot path pinned:
// IEEE-754 binary32, round-to-nearest, NO fused multiply-add (the library is compiled with
// --fmad=false; the traversal additionally never relies on reassociation), IEEE division
// and square root (nvcc defaults -prec-div=true -prec-sqrt=true).  Expression order follows
// the GLSL sources left to right (math.glsl and friends) so that results are reproducible
// against the CPU oracle bit for bit: the traversal uses +,-,*,/ and comparisons only, and
// shading takes sin/cos/pow/log/atan2 from include/zl_libm.h, the same code on both sides.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/zl_libm.h"   // sin/cos/atan2/log/pow: one implementation for device, oracle and oracle/_ref

#define ZL_DEV __device__ __forceinline__
// Out-of-line building blocks: the integrator kernels were 240 KB of SASS when everything was
// inlined (39 % of stall samples were instruction-cache misses, profiles/r1); each of these
// exists once per kernel instead of once per call site.
#define ZL_CALL __device__ __noinline__

namespace zl {

static constexpr float Pi = 3.14159265358979323846f;    // math.glsl:4
static constexpr float PiInv = 1.0f / Pi;               // math.glsl:5

ZL_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
ZL_DEV float3 f3(float a) { return make_float3(a, a, a); }
ZL_DEV float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
ZL_DEV float2 f2(float x, float y) { return make_float2(x, y); }

ZL_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
ZL_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
ZL_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
ZL_DEV float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
ZL_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
ZL_DEV float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
ZL_DEV float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
ZL_DEV float3 operator/(float s, float3 a) { return f3(s / a.x, s / a.y, s / a.z); }
ZL_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
ZL_DEV float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
ZL_DEV float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
ZL_DEV float3& operator*=(float3& a, float s) { a = a * s; return a; }
ZL_DEV float3& operator/=(float3& a, float s) { a = a / s; return a; }

ZL_DEV float2 operator+(float2 a, float2 b) { return f2(a.x + b.x, a.y + b.y); }
ZL_DEV float2 operator-(float2 a, float2 b) { return f2(a.x - b.x, a.y - b.y); }
ZL_DEV float2 operator*(float2 a, float2 b) { return f2(a.x * b.x, a.y * b.y); }
ZL_DEV float2 operator/(float2 a, float2 b) { return f2(a.x / b.x, a.y / b.y); }
ZL_DEV float2 operator*(float2 a, float s) { return f2(a.x * s, a.y * s); }
ZL_DEV float2 operator-(float2 a, float s) { return f2(a.x - s, a.y - s); }
ZL_DEV float2 operator+(float2 a, float s) { return f2(a.x + s, a.y + s); }

// GLSL min/max: min(x,y) = (y < x) ? y : x ; max(x,y) = (x < y) ? y : x  (NaN-order sensitive)
ZL_DEV float gmin(float x, float y) { return (y < x) ? y : x; }
ZL_DEV float gmax(float x, float y) { return (x < y) ? y : x; }
ZL_DEV float3 gmax(float3 a, float3 b) { return f3(gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z)); }
ZL_DEV float3 gmin(float3 a, float3 b) { return f3(gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z)); }
ZL_DEV float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }

ZL_DEV float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
ZL_DEV float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
ZL_DEV float3 cross(float3 a, float3 b) { return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
ZL_DEV float length(float3 a) { return sqrtf(dot(a, a)); }
ZL_DEV float length(float2 a) { return sqrtf(dot(a, a)); }
ZL_DEV float distance(float3 a, float3 b) { return length(a - b); }
// normalize(v) = v * (1 / sqrt(dot(v,v))): one IEEE reciprocal + three multiplies
ZL_DEV float3 normalize(float3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
ZL_DEV float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
ZL_DEV float3 mix(float3 a, float3 b, float t) { return a * (1.0f - t) + b * t; }
ZL_DEV float3 mix(float3 a, float3 b, float3 t) { return a * (f3(1.0f) - t) + b * t; }
ZL_DEV float fract(float x) { return x - floorf(x); }
ZL_DEV float3 reflect(float3 I, float3 N) { return I - N * (2.0f * dot(N, I)); }

ZL_DEV float square(float x) { return x * x; }                                           // math.glsl:8-11
ZL_DEV float biHeuristic(float f, float g) { return f * f / (f * f + g * g); }           // math.glsl:20-23
ZL_DEV float satDot(float3 a, float3 b) { return gmax(dot(a, b), 0.0f); }                // math.glsl:43-46
ZL_DEV float absDot(float3 a, float3 b) { return fabsf(dot(a, b)); }                     // math.glsl:48-51
ZL_DEV float distSquare(float3 x, float3 y) { return dot(x - y, x - y); }                // math.glsl:53-56
ZL_DEV float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }                     // math.glsl:187-191
ZL_DEV float luminance(float3 c) { return dot(c, f3(0.299f, 0.587f, 0.114f)); }          // math.glsl:193-196
ZL_DEV bool isBlack(float3 c) { return luminance(c) < 1e-5f; }                           // math.glsl:198-201
ZL_DEV bool hasNan(float3 c) { return isnan(c.x) || isnan(c.y) || isnan(c.z); }          // math.glsl:203-206
ZL_DEV float maxComponent(float3 v) { return gmax(v.x, gmax(v.y, v.z)); }                // math.glsl:120-123
ZL_DEV bool sameHemisphere(float3 n, float3 a, float3 b) { return dot(n, a) * dot(n, b) > 0; }  // math.glsl:107-110

ZL_DEV int maxExtent(float3 v) {                                                         // math.glsl:112-118
    if (v.x > v.y) return v.x > v.z ? 0 : 2;
    return v.y > v.z ? 1 : 2;
}
ZL_DEV int cubemapFace(float3 dir) {                                                     // math.glsl:125-131
    int maxDim = maxExtent(f3(fabsf(dir.x), fabsf(dir.y), fabsf(dir.z)));
    if (maxDim == 0) return dir.x > 0 ? 0 : 1;
    if (maxDim == 1) return dir.y > 0 ? 2 : 3;
    return dir.z > 0 ? 4 : 5;
}

ZL_CALL float2 toConcentricDisk(float2 v) {                                               // math.glsl:25-41
    if (v.x == 0.0f && v.y == 0.0f) return f2(0.0f, 0.0f);
    v = v * 2.0f - 1.0f;
    float phi, r;
    if (v.x * v.x > v.y * v.y) { r = v.x; phi = Pi * v.y / v.x * 0.25f; }
    else { r = v.y; phi = Pi * 0.5f - Pi * v.x / v.y * 0.25f; }
    float s, c;
    zl_sincosf(phi, &s, &c);
    return f2(r * c, r * s);
}
ZL_CALL float2 sphereToPlane(float3 uv) {                                                 // math.glsl:58-64
    float theta = zl_atan2f(uv.y, uv.x);
    if (theta < 0.0f) theta += Pi * 2.0f;
    float phi = zl_atan2f(length(f2(uv.x, uv.y)), uv.z);
    return f2(theta * PiInv * 0.5f, phi * PiInv);
}
ZL_CALL float3 planeToSphere(float2 uv) {                                                 // math.glsl:66-71
    float theta = uv.x * Pi * 2.0f, phi = uv.y * Pi;
    float st, ct, sp, cp;
    zl_sincosf(theta, &st, &ct);
    zl_sincosf(phi, &sp, &cp);
    return f3(ct * sp, st * sp, cp);
}
struct Mat3 { float3 c0, c1, c2; };   // column-major, like GLSL mat3
ZL_DEV float3 operator*(const Mat3& m, float3 v) { return m.c0 * v.x + m.c1 * v.y + m.c2 * v.z; }
ZL_DEV Mat3 inverse(const Mat3& m) {  // cofactor form of GLSL inverse(mat3)
    float a00 = m.c0.x, a01 = m.c0.y, a02 = m.c0.z;
    float a10 = m.c1.x, a11 = m.c1.y, a12 = m.c1.z;
    float a20 = m.c2.x, a21 = m.c2.y, a22 = m.c2.z;
    float k00 = a11 * a22 - a21 * a12;
    float k10 = a01 * a22 - a21 * a02;
    float k20 = a01 * a12 - a11 * a02;
    float det = a00 * k00 - a10 * k10 + a20 * k20;
    float inv = 1.0f / det;
    Mat3 r;
    r.c0 = f3(k00 * inv, -k10 * inv, k20 * inv);
    r.c1 = f3(-(a10 * a22 - a20 * a12) * inv, (a00 * a22 - a20 * a02) * inv, -(a00 * a12 - a10 * a02) * inv);
    r.c2 = f3((a10 * a21 - a20 * a11) * inv, -(a00 * a21 - a20 * a01) * inv, (a00 * a11 - a10 * a01) * inv);
    return r;
}
ZL_DEV float3 getTangent(float3 n) { return (fabsf(n.z) > 0.999f) ? f3(0, 1, 0) : f3(0, 0, 1); }   // math.glsl:73-76
ZL_DEV Mat3 tbnMatrix(float3 n) {                                                        // math.glsl:78-84
    float3 t = getTangent(n);
    float3 b = normalize(cross(n, t));
    t = cross(b, n);
    return Mat3{t, b, n};
}
ZL_DEV float3 normalToWorld(float3 n, float3 v) { return normalize(tbnMatrix(n) * v); }  // math.glsl:86-89
ZL_CALL float4 sampleCosineWeighted(float3 n, float2 u) {                                 // math.glsl:99-105
    float2 uv = toConcentricDisk(u);
    float z = sqrtf(1.0f - dot(uv, uv));
    float3 v = normalToWorld(n, f3(uv.x, uv.y, z));
    return make_float4(v.x, v.y, v.z, PiInv * z);
}
ZL_DEV float3 sampleTriangleUniform(float3 va, float3 vb, float3 vc, float2 uv) {        // math.glsl:139-145
    float r = sqrtf(uv.y);
    float u = 1.0f - r;
    float v = uv.x * r;
    return va * (1.0f - u - v) + vb * u + vc * v;
}
ZL_DEV float triangleArea(float3 va, float3 vb, float3 vc) { return 0.5f * length(cross(vc - va, vb - va)); }   // math.glsl:147-150
ZL_CALL float3 rotateZ(float3 v, float angle) {                                           // math.glsl:180-185
    float s, c;
    zl_sincosf(angle, &s, &c);
    return f3(v.x * c - v.y * s, v.x * s + v.y * c, v.z);
}

// random.glsl:5-13 (Wang hash)
ZL_DEV uint32_t hash(uint32_t seed) {
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}

}  // namespace zl
