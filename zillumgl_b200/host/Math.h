// Host-side linear algebra for scene preparation (stands in for glm, which the reference
// uses un-vendored: src/accelerator/AABB.h:6-8, src/core/Camera.h:6-8).  IEEE binary32,
// compiled with -ffp-contract=off so the flattened scene is reproducible bit for bit.
#pragma once
#include <cmath>
#include <cstdint>

namespace zillum {

struct Vec2f { float x = 0, y = 0; };
struct Vec3f {
    float x = 0, y = 0, z = 0;
    Vec3f() {}
    explicit Vec3f(float a) : x(a), y(a), z(a) {}
    Vec3f(float a, float b, float c) : x(a), y(b), z(c) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
inline Vec3f operator+(Vec3f a, Vec3f b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3f operator-(Vec3f a, Vec3f b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3f operator*(Vec3f a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3f operator/(Vec3f a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline Vec3f operator-(Vec3f a) { return {-a.x, -a.y, -a.z}; }
inline float dot(Vec3f a, Vec3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3f cross(Vec3f a, Vec3f b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float length(Vec3f a) { return std::sqrt(dot(a, a)); }
inline Vec3f normalize(Vec3f a) { float inv = 1.0f / std::sqrt(dot(a, a)); return a * inv; }
inline float fminf2(float a, float b) { return (b < a) ? b : a; }   // glm::min
inline float fmaxf2(float a, float b) { return (a < b) ? b : a; }   // glm::max
inline Vec3f vmin(Vec3f a, Vec3f b) { return {fminf2(a.x, b.x), fminf2(a.y, b.y), fminf2(a.z, b.z)}; }
inline Vec3f vmax(Vec3f a, Vec3f b) { return {fmaxf2(a.x, b.x), fmaxf2(a.y, b.y), fmaxf2(a.z, b.z)}; }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

struct Mat3f { Vec3f c0, c1, c2; };   // column-major
inline Vec3f operator*(const Mat3f& m, Vec3f v) { return m.c0 * v.x + m.c1 * v.y + m.c2 * v.z; }
inline Mat3f transpose(const Mat3f& m) {
    return {{m.c0.x, m.c1.x, m.c2.x}, {m.c0.y, m.c1.y, m.c2.y}, {m.c0.z, m.c1.z, m.c2.z}};
}
inline Mat3f inverse(const Mat3f& m) {   // cofactor form
    float a00 = m.c0.x, a01 = m.c0.y, a02 = m.c0.z;
    float a10 = m.c1.x, a11 = m.c1.y, a12 = m.c1.z;
    float a20 = m.c2.x, a21 = m.c2.y, a22 = m.c2.z;
    float k00 = a11 * a22 - a21 * a12;
    float k10 = a01 * a22 - a21 * a02;
    float k20 = a01 * a12 - a11 * a02;
    float det = a00 * k00 - a10 * k10 + a20 * k20;
    float inv = 1.0f / det;
    Mat3f r;
    r.c0 = {k00 * inv, -k10 * inv, k20 * inv};
    r.c1 = {-(a10 * a22 - a20 * a12) * inv, (a00 * a22 - a20 * a02) * inv, -(a00 * a12 - a10 * a02) * inv};
    r.c2 = {(a10 * a21 - a20 * a11) * inv, -(a00 * a21 - a20 * a01) * inv, (a00 * a11 - a10 * a01) * inv};
    return r;
}
// Rodrigues rotation about a unit axis (the 3x3 block glm::rotate builds).
inline Mat3f rotation(float angle, Vec3f axisIn) {
    float c = std::cos(angle), s = std::sin(angle);
    Vec3f axis = normalize(axisIn);
    Vec3f t = axis * (1.0f - c);
    Mat3f r;
    r.c0 = {c + t.x * axis.x, t.x * axis.y + s * axis.z, t.x * axis.z - s * axis.y};
    r.c1 = {t.y * axis.x - s * axis.z, c + t.y * axis.y, t.y * axis.z + s * axis.x};
    r.c2 = {t.z * axis.x + s * axis.y, t.z * axis.y - s * axis.x, c + t.z * axis.z};
    return r;
}
inline Mat3f operator*(const Mat3f& a, const Mat3f& b) { return {a * b.c0, a * b.c1, a * b.c2}; }

// Affine transform = 3x3 linear part + translation (the 4x4 model matrix of Model.cpp:62-72).
struct Affine {
    Mat3f m{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    Vec3f t;
    Vec3f point(Vec3f v) const { return (m.c0 * v.x + m.c1 * v.y) + (m.c2 * v.z + t); }
};

}  // namespace zillum
