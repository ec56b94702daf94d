// ORACLE — test infrastructure only.  CPU restatement of the reference GLSL hot path.
// Nothing under oracle/ is part of the product; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build, link or call it.
// PINNED: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md §4,
// §8c), so this restatement is pinned against the reference ITSELF instead — oracle/_ref, the
// reference's own C++ host code and GLSL text compiled for the host (Makefile target `ref`):
// tests/test_ref_parity.py holds every function below to the reference's result bit for bit.
//
// zo_vec.h — GLSL vector semantics pinned to IEEE-754 binary32, round-to-nearest, no FMA
// contraction (build with -ffp-contract=off).  Evaluation order is left-to-right as written
// in the GLSL source; see SURVEY.md App. D for the meaning given to each built-in.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace zo {

struct vec2 { float x, y; vec2() : x(0), y(0) {} vec2(float a) : x(a), y(a) {} vec2(float a, float b) : x(a), y(b) {} };
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(float a) : x(a), y(a), z(a) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    vec3(vec2 v, float c) : x(v.x), y(v.y), z(c) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a) : x(a), y(a), z(a), w(a) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    vec3 xyz() const { return vec3(x, y, z); }
};

inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(vec2 a, vec2 b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator/(vec2 a, vec2 b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec2 operator*(vec2 a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, vec2 a) { return vec2(s * a.x, s * a.y); }
inline vec2 operator/(vec2 a, float s) { return vec2(a.x / s, a.y / s); }
inline vec2 operator/(float s, vec2 a) { return vec2(s / a.x, s / a.y); }
inline vec2 operator-(vec2 a, float s) { return vec2(a.x - s, a.y - s); }
inline vec2 operator+(vec2 a, float s) { return vec2(a.x + s, a.y + s); }

inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator/(float s, vec3 a) { return vec3(s / a.x, s / a.y, s / a.z); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator+(vec3 a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(vec3 a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }
inline vec3& operator/=(vec3& a, vec3 b) { a = a / b; return a; }

// GLSL min/max: min(x,y) = y < x ? y : x ; max(x,y) = x < y ? y : x (spec 8.3).
inline float gmin(float x, float y) { return (y < x) ? y : x; }
inline float gmax(float x, float y) { return (x < y) ? y : x; }
inline vec3 gmin(vec3 a, vec3 b) { return vec3(gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z)); }
inline vec3 gmax(vec3 a, vec3 b) { return vec3(gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z)); }
inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
inline vec3 gabs(vec3 a) { return vec3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }

inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) {
    return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline float length(vec2 a) { return std::sqrt(dot(a, a)); }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline float distance(vec3 a, vec3 b) { return length(a - b); }
// normalize(v) = v * (1/sqrt(dot(v,v))): one IEEE reciprocal, three multiplies (the GL
// driver uses v * inversesqrt(dot); precision is driver-defined, we pin this form).
inline vec3 normalize(vec3 a) { float inv = 1.0f / std::sqrt(dot(a, a)); return a * inv; }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(vec3 a, vec3 b, vec3 t) { return a * (vec3(1.0f) - t) + b * t; }
inline float fract(float x) { return x - std::floor(x); }
inline vec2 fract(vec2 v) { return vec2(fract(v.x), fract(v.y)); }
inline vec3 reflect(vec3 I, vec3 N) { return I - N * (2.0f * dot(N, I)); }

// column-major 3x3
struct mat3 {
    vec3 c0, c1, c2;
    mat3() {}
    mat3(vec3 a, vec3 b, vec3 c) : c0(a), c1(b), c2(c) {}
};
inline vec3 operator*(const mat3& m, vec3 v) { return m.c0 * v.x + m.c1 * v.y + m.c2 * v.z; }
// cofactor inverse (GLSL inverse(mat3); glm::inverse on the host)
inline mat3 inverse(const mat3& m) {
    float a00 = m.c0.x, a01 = m.c0.y, a02 = m.c0.z;
    float a10 = m.c1.x, a11 = m.c1.y, a12 = m.c1.z;
    float a20 = m.c2.x, a21 = m.c2.y, a22 = m.c2.z;
    float k00 = a11 * a22 - a21 * a12;
    float k10 = a01 * a22 - a21 * a02;
    float k20 = a01 * a12 - a11 * a02;
    float det = a00 * k00 - a10 * k10 + a20 * k20;
    float inv = 1.0f / det;
    mat3 r;
    r.c0 = vec3(k00 * inv, -k10 * inv, k20 * inv);
    r.c1 = vec3(-(a10 * a22 - a20 * a12) * inv, (a00 * a22 - a20 * a02) * inv, -(a00 * a12 - a10 * a02) * inv);
    r.c2 = vec3((a10 * a21 - a20 * a11) * inv, -(a00 * a21 - a20 * a01) * inv, (a00 * a11 - a10 * a01) * inv);
    return r;
}

inline uint32_t floatBits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float bitsFloat(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// IEEE binary16 round-trip (RGB16F storage of the environment map, EnvironmentMap.cpp:13).
inline float roundToHalf(float f) {
    uint32_t x = floatBits(f);
    uint32_t sign = x & 0x80000000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return f;                        // inf / nan unchanged
    if (ax >= 0x477ff000u) return bitsFloat(sign | 0x7f800000u); // rounds to >= 65520 -> inf
    if (ax < 0x33000001u) return bitsFloat(sign);            // below half of min subnormal -> 0
    if (ax < 0x38800000u) {                                  // half subnormal: quantum 2^-24
        float a = bitsFloat(ax);
        float q = a * 16777216.0f;                           // exact scaling
        float r = std::nearbyint(q);                         // round-to-nearest-even
        return bitsFloat(sign | floatBits(r * (1.0f / 16777216.0f)));
    }
    uint32_t rem = ax & 0x1fffu, base = ax & ~0x1fffu;       // keep 10 mantissa bits
    if (rem > 0x1000u || (rem == 0x1000u && (base & 0x2000u))) base += 0x2000u;
    return bitsFloat(sign | base);
}

}  // namespace zo
