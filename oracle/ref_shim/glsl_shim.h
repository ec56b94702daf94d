// ORACLE — test infrastructure only.  Part of the recipe that builds oracle/_ref.
//
// glsl_shim.h — the GLSL 4.50 types and built-ins that the reference's shader text uses, as
// C++, so that the text of /root/reference/src/shader/*.glsl (rewritten mechanically by
// oracle/ref_glsl2cpp.py: `out/inout T x` -> `T& x`, swizzles -> member calls, float literals
// -> binary32, @type/@include resolved like src/core/Shader.cpp:180-259) compiles with g++ and
// runs on the host.  Nothing here restates the reference's algorithms: this file is the "GL
// driver" under the reference's own code.  What a driver leaves implementation-defined is
// pinned to the same choices as the CUDA path and the CPU oracle (DESIGN.md "Numerics"):
//   * IEEE-754 binary32, round to nearest, no contraction (built -ffp-contract=off);
//   * normalize(v) = v * (1 / sqrt(dot(v,v))), length = sqrt(dot), reflect = I - N*(2 dot(N,I)),
//     mix = a(1-t) + bt, min/max/clamp as the spec's comparisons, inverse(mat3) by cofactors;
//   * sin / cos / atan / asin / acos / log / pow / exp from include/zl_libm.h;
//   * texelFetch out of range returns 0; texture() = LINEAR + REPEAT bilinear in binary32 on
//     the stored texels; image stores / atomics out of range are dropped (robust GL behaviour).
// Function-argument evaluation order: GLSL evaluates left to right; C++ leaves it unspecified,
// so the rewriting script turns the four constructor calls whose arguments have side effects
// (random.glsl:45,74,79,84) into braced initialisation, which C++ orders left to right.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include "../../include/zl_libm.h"

namespace glsl {

typedef unsigned int uint;

struct ivec2; struct uvec2; struct vec3; struct vec4;

struct vec2 {
    union { struct { float x, y; }; struct { float r, g; }; };
    vec2() : x(0), y(0) {}
    explicit vec2(float a) : x(a), y(a) {}
    vec2(float a, float b) : x(a), y(b) {}
    explicit vec2(const ivec2& v);
    vec2 xy() const { return *this; }
    vec2 rg() const { return *this; }
};
struct vec3 {
    union { struct { float x, y, z; }; struct { float r, g, b; }; };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    vec3(vec2 v, float c) : x(v.x), y(v.y), z(c) {}
    vec3(float a, vec2 v) : x(a), y(v.x), z(v.y) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
    vec2 xy() const { return vec2(x, y); }
    vec2 yz() const { return vec2(y, z); }
    vec2 rg() const { return vec2(x, y); }
    vec3 xyz() const { return *this; }
    vec3 rgb() const { return *this; }
};
struct vec4 {
    union { struct { float x, y, z, w; }; struct { float r, g, b, a; }; };
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float a_, float b_, float c_, float d_) : x(a_), y(b_), z(c_), w(d_) {}
    vec4(vec3 v, float d_) : x(v.x), y(v.y), z(v.z), w(d_) {}
    vec4(vec2 p, vec2 q) : x(p.x), y(p.y), z(q.x), w(q.y) {}
    vec2 xy() const { return vec2(x, y); }
    vec2 zw() const { return vec2(z, w); }
    vec2 yz() const { return vec2(y, z); }
    vec3 xyz() const { return vec3(x, y, z); }
    vec3 rgb() const { return vec3(x, y, z); }
};
struct ivec2 {
    union { struct { int x, y; }; struct { int r, g; }; };
    ivec2() : x(0), y(0) {}
    explicit ivec2(int a) : x(a), y(a) {}
    ivec2(int a, int b) : x(a), y(b) {}
    explicit ivec2(const vec2& v) : x((int)v.x), y((int)v.y) {}           // truncation toward zero
    explicit ivec2(const uvec2& v);
    ivec2 xy() const { return *this; }
};
struct ivec3 { int x, y, z; ivec3() : x(0), y(0), z(0) {} ivec3(int a, int b, int c) : x(a), y(b), z(c) {} ivec2 xy() const { return ivec2(x, y); } };
struct ivec4 {
    union { struct { int x, y, z, w; }; struct { int r, g, b, a; }; };
    ivec4() : x(0), y(0), z(0), w(0) {}
    ivec4(int a_, int b_, int c_, int d_) : x(a_), y(b_), z(c_), w(d_) {}
};
struct uvec2 { uint x, y; uvec2() : x(0), y(0) {} uvec2(uint a, uint b) : x(a), y(b) {} };
struct uvec3 {
    uint x, y, z;
    uvec3() : x(0), y(0), z(0) {}
    uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
    uvec2 xy() const { return uvec2(x, y); }
};
struct uvec4 {
    union { struct { uint x, y, z, w; }; struct { uint r, g, b, a; }; };
    uvec4() : x(0), y(0), z(0), w(0) {}
    uvec4(uint a_, uint b_, uint c_, uint d_) : x(a_), y(b_), z(c_), w(d_) {}
};
inline vec2::vec2(const ivec2& v) : x((float)v.x), y((float)v.y) {}
inline ivec2::ivec2(const uvec2& v) : x((int)v.x), y((int)v.y) {}

// ---- operators (component-wise, one IEEE operation per component, operands in source order) ----
#define GLSL_VEC_OPS(V, ...)                                                                      \
    inline V operator+(V a, V b) { return V(__VA_ARGS__(+)); }                                    \
    inline V operator-(V a, V b) { return V(__VA_ARGS__(-)); }                                    \
    inline V operator*(V a, V b) { return V(__VA_ARGS__(*)); }                                    \
    inline V operator/(V a, V b) { return V(__VA_ARGS__(/)); }
#define GLSL_C2(op) a.x op b.x, a.y op b.y
#define GLSL_C3(op) a.x op b.x, a.y op b.y, a.z op b.z
#define GLSL_C4(op) a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w
GLSL_VEC_OPS(vec2, GLSL_C2)
GLSL_VEC_OPS(vec3, GLSL_C3)
GLSL_VEC_OPS(vec4, GLSL_C4)
#undef GLSL_VEC_OPS
#define GLSL_SCALAR_OPS(V)                                                                        \
    inline V operator+(V a, float s) { return a + V(s); }                                         \
    inline V operator-(V a, float s) { return a - V(s); }                                         \
    inline V operator*(V a, float s) { return a * V(s); }                                         \
    inline V operator/(V a, float s) { return a / V(s); }                                         \
    inline V operator+(float s, V a) { return V(s) + a; }                                         \
    inline V operator-(float s, V a) { return V(s) - a; }                                         \
    inline V operator*(float s, V a) { return V(s) * a; }                                         \
    inline V operator/(float s, V a) { return V(s) / a; }                                         \
    inline V& operator+=(V& a, V b) { a = a + b; return a; }                                      \
    inline V& operator-=(V& a, V b) { a = a - b; return a; }                                      \
    inline V& operator*=(V& a, V b) { a = a * b; return a; }                                      \
    inline V& operator/=(V& a, V b) { a = a / b; return a; }                                      \
    inline V& operator+=(V& a, float s) { a = a + s; return a; }                                  \
    inline V& operator-=(V& a, float s) { a = a - s; return a; }                                  \
    inline V& operator*=(V& a, float s) { a = a * s; return a; }                                  \
    inline V& operator/=(V& a, float s) { a = a / s; return a; }
GLSL_SCALAR_OPS(vec2)
GLSL_SCALAR_OPS(vec3)
GLSL_SCALAR_OPS(vec4)
#undef GLSL_SCALAR_OPS
inline vec2 operator-(vec2 a) { return vec2(-a.x, -a.y); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec4 operator-(vec4 a) { return vec4(-a.x, -a.y, -a.z, -a.w); }
inline ivec2 operator/(ivec2 a, int s) { return ivec2(a.x / s, a.y / s); }
inline ivec2 operator*(ivec2 a, int s) { return ivec2(a.x * s, a.y * s); }
inline ivec2 operator+(ivec2 a, ivec2 b) { return ivec2(a.x + b.x, a.y + b.y); }
inline bool operator!=(ivec2 a, ivec2 b) { return a.x != b.x || a.y != b.y; }

// ---- common functions (GLSL 4.50 §8.3) ----
inline float abs(float x) { return zl_fabsf(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline vec2 abs(vec2 v) { return vec2(abs(v.x), abs(v.y)); }
inline vec3 abs(vec3 v) { return vec3(abs(v.x), abs(v.y), abs(v.z)); }
inline float min(float x, float y) { return (y < x) ? y : x; }
inline float max(float x, float y) { return (x < y) ? y : x; }
inline int min(int x, int y) { return (y < x) ? y : x; }
inline int max(int x, int y) { return (x < y) ? y : x; }
inline vec2 min(vec2 a, vec2 b) { return vec2(min(a.x, b.x), min(a.y, b.y)); }
inline vec2 max(vec2 a, vec2 b) { return vec2(max(a.x, b.x), max(a.y, b.y)); }
inline vec3 min(vec3 a, vec3 b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 max(vec3 a, vec3 b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline vec3 min(vec3 a, float b) { return min(a, vec3(b)); }
inline vec3 max(vec3 a, float b) { return max(a, vec3(b)); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline vec3 clamp(vec3 v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec2 mix(vec2 a, vec2 b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(vec3 a, vec3 b, vec3 t) { return a * (vec3(1.0f) - t) + b * t; }
inline float floor(float x) { return std::floor(x); }
inline float fract(float x) { return x - std::floor(x); }
inline vec2 fract(vec2 v) { return vec2(fract(v.x), fract(v.y)); }
inline vec3 fract(vec3 v) { return vec3(fract(v.x), fract(v.y), fract(v.z)); }
inline bool isnan(float x) { return x != x; }
inline bool isinf(float x) { return zl_fabsf(x) == zl_u2f(0x7f800000u); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }

// ---- angle / exponential functions (§8.1, §8.2): include/zl_libm.h ----
inline float sin(float x) { return zl_sinf(x); }
inline float cos(float x) { return zl_cosf(x); }
inline float tan(float x) { float s, c; zl_sincosf(x, &s, &c); return s / c; }
inline float asin(float x) { return zl_asinf(x); }
inline float acos(float x) { return zl_acosf(x); }
inline float atan(float y, float x) { return zl_atan2f(y, x); }
inline float atan(float x) { return zl_atanf(x); }
inline float pow(float x, float y) { return zl_powf(x, y); }
inline vec3 pow(vec3 x, vec3 y) { return vec3(pow(x.x, y.x), pow(x.y, y.y), pow(x.z, y.z)); }
inline float exp(float x) { return zl_expf(x); }
inline vec3 exp(vec3 v) { return vec3(exp(v.x), exp(v.y), exp(v.z)); }
inline float log(float x) { return zl_logf(x); }

// ---- geometric functions (§8.5) ----
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(vec2 a) { return std::sqrt(dot(a, a)); }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline float distance(vec3 a, vec3 b) { return length(a - b); }
inline vec3 normalize(vec3 a) { float inv = 1.0f / std::sqrt(dot(a, a)); return a * inv; }
inline vec3 reflect(vec3 I, vec3 N) { return I - N * (2.0f * dot(N, I)); }

// ---- mat3 (column-major) ----
struct mat3 {
    vec3 c0, c1, c2;
    mat3() {}
    mat3(vec3 a, vec3 b, vec3 c) : c0(a), c1(b), c2(c) {}
    vec3& operator[](int i) { return (&c0)[i]; }
};
inline vec3 operator*(const mat3& m, vec3 v) { return m.c0 * v.x + m.c1 * v.y + m.c2 * v.z; }
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

// ---- opaque types: buffer textures, 2-D textures, texture arrays, images ----
// A bound object is a host pointer + geometry + the element layout the GL format implies.
struct TexBinding {
    const void* data = nullptr;     // texel storage (float / int32 / uint32 / uint8 per `kind`)
    int comps = 0;                  // components per texel (1..4)
    int w = 0, h = 1, layers = 1;   // buffer textures: w = element count
    int kind = 0;                   // 0 float32, 1 int32 / uint32 bits, 2 sRGB8 (decode through lut)
    const float* lut = nullptr;     // sRGB decode table for kind 2
    uint32_t (*generator)(int) = nullptr;   // buffer textures whose texels are computed on demand (uSobolSeq)
};
struct samplerBuffer : TexBinding {};
struct isamplerBuffer : TexBinding {};
struct usamplerBuffer : TexBinding {};
struct sampler2D : TexBinding {};
struct isampler2D : TexBinding {};
struct sampler2DArray : TexBinding {};
struct image2D { float* data = nullptr; int w = 0, h = 0, comps = 0; };

// texelFetch(gsamplerBuffer, i): missing components read (0, 0, 0, 1); out of range reads 0 (robust buffer access)
inline vec4 texelFetch(const samplerBuffer& s, int i) {
    if (i < 0 || i >= s.w || !s.data) return vec4(0.0f, 0.0f, 0.0f, 0.0f);
    const float* p = (const float*)s.data + (size_t)i * s.comps;
    return vec4(p[0], s.comps > 1 ? p[1] : 0.0f, s.comps > 2 ? p[2] : 0.0f, s.comps > 3 ? p[3] : 1.0f);
}
inline ivec4 texelFetch(const isamplerBuffer& s, int i) {
    if (i < 0 || i >= s.w || !s.data) return ivec4(0, 0, 0, 0);
    const int32_t* p = (const int32_t*)s.data + (size_t)i * s.comps;
    return ivec4(p[0], s.comps > 1 ? p[1] : 0, s.comps > 2 ? p[2] : 0, s.comps > 3 ? p[3] : 1);
}
inline uvec4 texelFetch(const usamplerBuffer& s, int i) {
    if (s.generator) return uvec4(s.generator(i), 0u, 0u, 1u);
    if (i < 0 || i >= s.w || !s.data) return uvec4(0u, 0u, 0u, 0u);
    const uint32_t* p = (const uint32_t*)s.data + (size_t)i * s.comps;
    return uvec4(p[0], s.comps > 1 ? p[1] : 0u, s.comps > 2 ? p[2] : 0u, s.comps > 3 ? p[3] : 1u);
}
inline vec4 texelFetch(const sampler2D& s, ivec2 c, int /*lod*/) {
    if (c.x < 0 || c.y < 0 || c.x >= s.w || c.y >= s.h || !s.data) return vec4(0.0f, 0.0f, 0.0f, 0.0f);
    const float* p = (const float*)s.data + ((size_t)c.y * s.w + c.x) * s.comps;
    return vec4(p[0], s.comps > 1 ? p[1] : 0.0f, s.comps > 2 ? p[2] : 0.0f, s.comps > 3 ? p[3] : 1.0f);
}
inline ivec4 texelFetch(const isampler2D& s, ivec2 c, int /*lod*/) {
    if (c.x < 0 || c.y < 0 || c.x >= s.w || c.y >= s.h || !s.data) return ivec4(0, 0, 0, 0);
    const int32_t* p = (const int32_t*)s.data + ((size_t)c.y * s.w + c.x) * s.comps;
    return ivec4(p[0], s.comps > 1 ? p[1] : 0, s.comps > 2 ? p[2] : 0, s.comps > 3 ? p[3] : 1);
}
inline ivec2 textureSize(const sampler2D& s, int /*lod*/) { return ivec2(s.w, s.h); }

// GL_LINEAR + GL_REPEAT footprint: texel centres at (i + 0.5) / size
struct Bilerp { int i0, i1; float f; };
inline Bilerp bilerpRepeat(float u, int size) {
    float x = u * (float)size - 0.5f;
    float fl = std::floor(x);
    Bilerp b;
    b.f = x - fl;
    int i = (int)fl;
    int m = i % size; if (m < 0) m += size;
    b.i0 = m;
    b.i1 = (m + 1 == size) ? 0 : m + 1;
    return b;
}
inline vec4 texelAt(const TexBinding& s, int x, int y, int layer) {
    size_t idx = (((size_t)layer * s.h + y) * s.w + x) * s.comps;
    float c[4] = {0.0f, 0.0f, 0.0f, 1.0f};
    if (s.kind == 2) { const uint8_t* p = (const uint8_t*)s.data + idx; for (int k = 0; k < s.comps; k++) c[k] = s.lut[p[k]]; }
    else { const float* p = (const float*)s.data + idx; for (int k = 0; k < s.comps; k++) c[k] = p[k]; }
    return vec4(c[0], c[1], c[2], c[3]);
}
inline vec4 bilinear(const TexBinding& s, vec2 uv, int layer) {
    Bilerp bx = bilerpRepeat(uv.x, s.w), by = bilerpRepeat(uv.y, s.h);
    vec4 a = texelAt(s, bx.i0, by.i0, layer) * (1.0f - bx.f) + texelAt(s, bx.i1, by.i0, layer) * bx.f;
    vec4 b = texelAt(s, bx.i0, by.i1, layer) * (1.0f - bx.f) + texelAt(s, bx.i1, by.i1, layer) * bx.f;
    return a * (1.0f - by.f) + b * by.f;
}
inline vec4 texture(const sampler2D& s, vec2 uv) {
    if (!s.data) return vec4(0.0f, 0.0f, 0.0f, 1.0f);
    return bilinear(s, uv, 0);
}
// texture2DArray(sampler, vec3(uv, layer)) (GL_EXT_texture_array): layer = round(z) clamped; an unbound or empty array reads 0
inline vec4 texture2DArray(const sampler2DArray& s, vec3 p) {
    int layer = (int)std::floor(p.z + 0.5f);
    if (!s.data || layer < 0 || layer >= s.layers) return vec4(0.0f, 0.0f, 0.0f, 0.0f);
    return bilinear(s, vec2(p.x, p.y), layer);
}
inline vec4 imageLoad(const image2D& im, ivec2 c) {
    if (c.x < 0 || c.y < 0 || c.x >= im.w || c.y >= im.h) return vec4(0.0f, 0.0f, 0.0f, 0.0f);
    const float* p = im.data + ((size_t)c.y * im.w + c.x) * im.comps;
    return vec4(p[0], im.comps > 1 ? p[1] : 0.0f, im.comps > 2 ? p[2] : 0.0f, im.comps > 3 ? p[3] : 1.0f);
}
inline void imageStore(const image2D& im, ivec2 c, vec4 v) {
    if (c.x < 0 || c.y < 0 || c.x >= im.w || c.y >= im.h) return;
    float* p = im.data + ((size_t)c.y * im.w + c.x) * im.comps;
    const float s[4] = {v.x, v.y, v.z, v.w};
    for (int k = 0; k < im.comps; k++) p[k] = s[k];
}
inline float imageAtomicAdd(const image2D& im, ivec2 c, float v) {          // GL_NV_shader_atomic_float
    if (c.x < 0 || c.y < 0 || c.x >= im.w || c.y >= im.h) return 0.0f;
    float* p = im.data + ((size_t)c.y * im.w + c.x) * im.comps;
    float old;
#pragma omp atomic capture
    { old = *p; *p += v; }
    return old;
}

// ---- program registry (filled by the generated translation units) ----
enum UniformType { U_INT, U_UINT, U_FLOAT, U_BOOL, U_VEC2, U_VEC3, U_VEC4, U_IVEC2, U_MAT3,
                   U_SAMPLER_BUFFER, U_ISAMPLER_BUFFER, U_USAMPLER_BUFFER, U_SAMPLER_2D, U_ISAMPLER_2D,
                   U_SAMPLER_2D_ARRAY, U_IMAGE_2D };
struct UniformEntry { const char* name; UniformType type; void* ptr; int binding; const char* format; };
struct Program {
    const char* name;
    const UniformEntry* uniforms; int numUniforms;
    int localSize[3];
    void (*invoke)(uint gx, uint gy, uint gz);                      // sets gl_GlobalInvocationID, runs main()
    int (*kat)(int op, const float* in, int inStride, float* out, int outStride, size_t n);   // library known-answer hook
    int (*trace)(const float* rays, size_t n, int anyhit, const float* tMax, int32_t* ids, float* t, int32_t* steps);
    Program* next;
};
Program*& programList();
inline void registerProgram(Program* p) { p->next = programList(); programList() = p; }

}  // namespace glsl
