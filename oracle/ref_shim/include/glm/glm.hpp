// ORACLE — test infrastructure only.  Part of the recipe that builds oracle/_ref.
//
// glm/glm.hpp — the subset of OpenGL Mathematics (g-truc/glm; the reference does not vendor or pin
// it, restated here from the published 0.9.9.8 sources' algorithms) that the reference's host code
// uses, so that src/accelerator/{AABB,BVH}.cpp, src/core/{Camera,Model,Scene,Sampler,
// EnvironmentMap,MaterialLoader}.cpp and src/integrator/{NaivePath,LightPath,TriplePath}.cpp compile
// unmodified, where they lie, with g++.  Scalar (non-SIMD) code paths, binary32, same expression
// order as glm: dot = tmp.x + tmp.y + tmp.z over the component products, normalize = v *
// inversesqrt(dot(v, v)) with inversesqrt(x) = 1 / sqrt(x), min / max by the comparison, mat4 * vec4
// as (m0 v0 + m1 v1) + (m2 v2 + m3 v3), mat4 * mat4 column by column left to right, the cofactor
// inverses of func_matrix.inl, rotate / translate / scale / lookAt / perspective of
// ext/matrix_transform.inl and ext/matrix_clip_space.inl (right-handed, [-1, 1] depth).
#pragma once
#include <cmath>
#include <cstddef>

namespace glm {

template <typename T> struct tvec2;
template <typename T> struct tvec3;
template <typename T> struct tvec4;

template <typename T> struct tvec2 {
    union { T x, r, s; }; union { T y, g, t; };
    tvec2() : x(0), y(0) {}
    explicit tvec2(T a) : x(a), y(a) {}
    template <typename A, typename B> tvec2(A a, B b) : x((T)a), y((T)b) {}
    template <typename U> tvec2(const tvec2<U>& v) : x((T)v.x), y((T)v.y) {}
    template <typename U> tvec2(const tvec3<U>& v);
    T& operator[](int i) { return i == 0 ? x : y; }
    const T& operator[](int i) const { return i == 0 ? x : y; }
};
template <typename T> struct tvec3 {
    union { T x, r, s; }; union { T y, g, t; }; union { T z, b, p; };
    tvec3() : x(0), y(0), z(0) {}
    explicit tvec3(T a) : x(a), y(a), z(a) {}
    template <typename A, typename B, typename C> tvec3(A a, B b_, C c) : x((T)a), y((T)b_), z((T)c) {}
    template <typename U> tvec3(const tvec2<U>& v, T c) : x((T)v.x), y((T)v.y), z(c) {}
    template <typename U> tvec3(const tvec3<U>& v) : x((T)v.x), y((T)v.y), z((T)v.z) {}
    template <typename U> tvec3(const tvec4<U>& v);
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    tvec3& operator-=(const tvec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    tvec3& operator*=(T s_) { x *= s_; y *= s_; z *= s_; return *this; }
    tvec3& operator/=(T s_) { x /= s_; y /= s_; z /= s_; return *this; }
};
template <typename T> struct tvec4 {
    union { T x, r, s; }; union { T y, g, t; }; union { T z, b, p; }; union { T w, a, q; };
    tvec4() : x(0), y(0), z(0), w(0) {}
    explicit tvec4(T v) : x(v), y(v), z(v), w(v) {}
    template <typename A, typename B, typename C, typename D> tvec4(A a_, B b_, C c, D d) : x((T)a_), y((T)b_), z((T)c), w((T)d) {}
    template <typename U> tvec4(const tvec3<U>& v, T d) : x((T)v.x), y((T)v.y), z((T)v.z), w(d) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};
template <typename T> template <typename U> tvec2<T>::tvec2(const tvec3<U>& v) : x((T)v.x), y((T)v.y) {}
template <typename T> template <typename U> tvec3<T>::tvec3(const tvec4<U>& v) : x((T)v.x), y((T)v.y), z((T)v.z) {}

typedef tvec2<float> vec2; typedef tvec3<float> vec3; typedef tvec4<float> vec4;
typedef tvec2<int> ivec2; typedef tvec3<int> ivec3; typedef tvec4<int> ivec4;

#define ZGLM_OPS2(op) \
    template <typename T> tvec2<T> operator op(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x op b.x, a.y op b.y); } \
    template <typename T> tvec2<T> operator op(const tvec2<T>& a, T s) { return tvec2<T>(a.x op s, a.y op s); } \
    template <typename T> tvec2<T> operator op(T s, const tvec2<T>& a) { return tvec2<T>(s op a.x, s op a.y); }
#define ZGLM_OPS3(op) \
    template <typename T> tvec3<T> operator op(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x op b.x, a.y op b.y, a.z op b.z); } \
    template <typename T> tvec3<T> operator op(const tvec3<T>& a, T s) { return tvec3<T>(a.x op s, a.y op s, a.z op s); } \
    template <typename T> tvec3<T> operator op(T s, const tvec3<T>& a) { return tvec3<T>(s op a.x, s op a.y, s op a.z); }
#define ZGLM_OPS4(op) \
    template <typename T> tvec4<T> operator op(const tvec4<T>& a, const tvec4<T>& b) { return tvec4<T>(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); } \
    template <typename T> tvec4<T> operator op(const tvec4<T>& a, T s) { return tvec4<T>(a.x op s, a.y op s, a.z op s, a.w op s); } \
    template <typename T> tvec4<T> operator op(T s, const tvec4<T>& a) { return tvec4<T>(s op a.x, s op a.y, s op a.z, s op a.w); }
ZGLM_OPS2(+) ZGLM_OPS2(-) ZGLM_OPS2(*) ZGLM_OPS2(/)
ZGLM_OPS3(+) ZGLM_OPS3(-) ZGLM_OPS3(*) ZGLM_OPS3(/)
ZGLM_OPS4(+) ZGLM_OPS4(-) ZGLM_OPS4(*) ZGLM_OPS4(/)
#undef ZGLM_OPS2
#undef ZGLM_OPS3
#undef ZGLM_OPS4
template <typename T> tvec3<T> operator-(const tvec3<T>& a) { return tvec3<T>(-a.x, -a.y, -a.z); }
template <typename T> bool operator==(const tvec2<T>& a, const tvec2<T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> bool operator!=(const tvec2<T>& a, const tvec2<T>& b) { return !(a == b); }
template <typename T> bool operator==(const tvec3<T>& a, const tvec3<T>& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <typename T> bool operator!=(const tvec3<T>& a, const tvec3<T>& b) { return !(a == b); }

// func_trigonometric / func_exponential / func_common (scalar paths)
inline float radians(float d) { return d * 0.01745329251994329576923690768489f; }
inline float degrees(float r) { return r * 57.295779513082320876798154814105f; }
inline float sin(float x) { return std::sin(x); }
inline float cos(float x) { return std::cos(x); }
inline float tan(float x) { return std::tan(x); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
template <typename T> T min(T x, T y) { return (y < x) ? y : x; }
template <typename T> T max(T x, T y) { return (x < y) ? y : x; }
template <typename T> tvec2<T> min(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(min(a.x, b.x), min(a.y, b.y)); }
template <typename T> tvec2<T> max(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(max(a.x, b.x), max(a.y, b.y)); }
template <typename T> tvec3<T> min(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
template <typename T> tvec3<T> max(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }

// func_geometric
inline float dot(const vec2& a, const vec2& b) { vec2 t(a * b); return t.x + t.y; }
inline float dot(const vec3& a, const vec3& b) { vec3 t(a * b); return t.x + t.y + t.z; }
inline float dot(const vec4& a, const vec4& b) { vec4 t(a * b); return (t.x + t.y) + (t.z + t.w); }
inline vec3 cross(const vec3& x, const vec3& y) { return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
inline float length(const vec2& v) { return std::sqrt(dot(v, v)); }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
inline vec2 normalize(const vec2& v) { return v * inversesqrt(dot(v, v)); }

struct mat4;
struct mat3 {
    vec3 c[3];
    mat3() {}
    explicit mat3(float d) { c[0] = vec3(d, 0, 0); c[1] = vec3(0, d, 0); c[2] = vec3(0, 0, d); }
    mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    mat3(const mat4& m);
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
struct mat4 {
    vec4 c[4];
    mat4() {}
    explicit mat4(float d) { c[0] = vec4(d, 0, 0, 0); c[1] = vec4(0, d, 0, 0); c[2] = vec4(0, 0, d, 0); c[3] = vec4(0, 0, 0, d); }
    mat4(const vec4& a, const vec4& b, const vec4& d, const vec4& e) { c[0] = a; c[1] = b; c[2] = d; c[3] = e; }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline mat3::mat3(const mat4& m) { c[0] = vec3(m[0]); c[1] = vec3(m[1]); c[2] = vec3(m[2]); }

// type_mat3x3.inl / type_mat4x4.inl
inline vec3 operator*(const mat3& m, const vec3& v) {
    return vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
                m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
                m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}
inline vec4 operator*(const mat4& m, const vec4& v) {
    const vec4 Mov0(v[0]), Mov1(v[1]);
    const vec4 Mul0 = m[0] * Mov0, Mul1 = m[1] * Mov1;
    const vec4 Add0 = Mul0 + Mul1;
    const vec4 Mov2(v[2]), Mov3(v[3]);
    const vec4 Mul2 = m[2] * Mov2, Mul3 = m[3] * Mov3;
    const vec4 Add1 = Mul2 + Mul3;
    return Add0 + Add1;
}
inline mat4 operator*(const mat4& m1, const mat4& m2) {
    const vec4 A0 = m1[0], A1 = m1[1], A2 = m1[2], A3 = m1[3];
    const vec4 B0 = m2[0], B1 = m2[1], B2 = m2[2], B3 = m2[3];
    mat4 r;
    r[0] = A0 * B0[0] + A1 * B0[1] + A2 * B0[2] + A3 * B0[3];
    r[1] = A0 * B1[0] + A1 * B1[1] + A2 * B1[2] + A3 * B1[3];
    r[2] = A0 * B2[0] + A1 * B2[1] + A2 * B2[2] + A3 * B2[3];
    r[3] = A0 * B3[0] + A1 * B3[1] + A2 * B3[2] + A3 * B3[3];
    return r;
}
inline mat4 operator*(const mat4& m, float s) { return mat4(m[0] * s, m[1] * s, m[2] * s, m[3] * s); }

// func_matrix.inl
inline mat3 transpose(const mat3& m) {
    mat3 r;
    r[0][0] = m[0][0]; r[0][1] = m[1][0]; r[0][2] = m[2][0];
    r[1][0] = m[0][1]; r[1][1] = m[1][1]; r[1][2] = m[2][1];
    r[2][0] = m[0][2]; r[2][1] = m[1][2]; r[2][2] = m[2][2];
    return r;
}
inline mat4 transpose(const mat4& m) {
    mat4 r;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r[i][j] = m[j][i];
    return r;
}
inline mat3 inverse(const mat3& m) {
    float OneOverDeterminant = 1.0f / (
        + m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2])
        - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])
        + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
    mat3 Inverse;
    Inverse[0][0] = + (m[1][1] * m[2][2] - m[2][1] * m[1][2]) * OneOverDeterminant;
    Inverse[1][0] = - (m[1][0] * m[2][2] - m[2][0] * m[1][2]) * OneOverDeterminant;
    Inverse[2][0] = + (m[1][0] * m[2][1] - m[2][0] * m[1][1]) * OneOverDeterminant;
    Inverse[0][1] = - (m[0][1] * m[2][2] - m[2][1] * m[0][2]) * OneOverDeterminant;
    Inverse[1][1] = + (m[0][0] * m[2][2] - m[2][0] * m[0][2]) * OneOverDeterminant;
    Inverse[2][1] = - (m[0][0] * m[2][1] - m[2][0] * m[0][1]) * OneOverDeterminant;
    Inverse[0][2] = + (m[0][1] * m[1][2] - m[1][1] * m[0][2]) * OneOverDeterminant;
    Inverse[1][2] = - (m[0][0] * m[1][2] - m[1][0] * m[0][2]) * OneOverDeterminant;
    Inverse[2][2] = + (m[0][0] * m[1][1] - m[1][0] * m[0][1]) * OneOverDeterminant;
    return Inverse;
}
inline mat4 inverse(const mat4& m) {
    float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
    float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
    float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
    float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
    float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
    float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
    float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
    float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    vec4 Fac0(Coef00, Coef00, Coef02, Coef03);
    vec4 Fac1(Coef04, Coef04, Coef06, Coef07);
    vec4 Fac2(Coef08, Coef08, Coef10, Coef11);
    vec4 Fac3(Coef12, Coef12, Coef14, Coef15);
    vec4 Fac4(Coef16, Coef16, Coef18, Coef19);
    vec4 Fac5(Coef20, Coef20, Coef22, Coef23);
    vec4 Vec0(m[1][0], m[0][0], m[0][0], m[0][0]);
    vec4 Vec1(m[1][1], m[0][1], m[0][1], m[0][1]);
    vec4 Vec2(m[1][2], m[0][2], m[0][2], m[0][2]);
    vec4 Vec3(m[1][3], m[0][3], m[0][3], m[0][3]);
    vec4 Inv0(Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2);
    vec4 Inv1(Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4);
    vec4 Inv2(Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5);
    vec4 Inv3(Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5);
    vec4 SignA(+1.0f, -1.0f, +1.0f, -1.0f);
    vec4 SignB(-1.0f, +1.0f, -1.0f, +1.0f);
    mat4 Inverse(Inv0 * SignA, Inv1 * SignB, Inv2 * SignA, Inv3 * SignB);
    vec4 Row0(Inverse[0][0], Inverse[1][0], Inverse[2][0], Inverse[3][0]);
    vec4 Dot0(m[0] * Row0);
    float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
    float OneOverDeterminant = 1.0f / Dot1;
    return Inverse * OneOverDeterminant;
}

}  // namespace glm
