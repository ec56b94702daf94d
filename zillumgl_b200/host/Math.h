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

// The normal matrix of Scene.cpp:153-154: mat3(transpose(inverse(model))) with `model` the 4x4 matrix.  glm inverts a mat4 through
// its 2x2 sub-determinants of rows 1-3 (func_matrix.inl, compute_inverse for 4x4) and divides by the determinant expanded along
// row 0; inverting only the 3x3 linear part is the same matrix on paper but rounds differently once scale and rotation mix
// (1 ulp on ~15 % of the normals of a randomly transformed instance: tests/test_ref_parity.py::test_random_scene_xml_equals_reference_scene).
// So: the 4x4 expansion, operation for operation, on [m | t; 0 0 0 1].
inline Mat3f normalMatrix(const Affine& a) {
    const float M[4][4] = {{a.m.c0.x, a.m.c0.y, a.m.c0.z, 0.0f}, {a.m.c1.x, a.m.c1.y, a.m.c1.z, 0.0f},
                           {a.m.c2.x, a.m.c2.y, a.m.c2.z, 0.0f}, {a.t.x, a.t.y, a.t.z, 1.0f}};       // M[column][row]
    // sub(r, s, p, q) = the 2x2 determinant of rows r < s, columns p < q
    auto sub = [&](int r, int s, int p, int q) { return M[p][r] * M[q][s] - M[q][r] * M[p][s]; };
    // six factor vectors, one per row pair; their lanes run over the column pairs (2,3), (2,3), (1,3), (1,2)
    static const int colP[4] = {2, 2, 1, 1}, colQ[4] = {3, 3, 3, 2};
    float inv[4][4];                                                                                   // inv[column][row], before the division
    for (int l = 0; l < 4; l++) {
        const int p = colP[l], q = colQ[l];
        const float f0 = sub(2, 3, p, q), f1 = sub(1, 3, p, q), f2 = sub(1, 2, p, q), f3 = sub(0, 3, p, q), f4 = sub(0, 2, p, q), f5 = sub(0, 1, p, q);
        const int c = l == 0 ? 1 : 0;                      // lane 0 multiplies by column 1 of the source, the other lanes by column 0
        const float v0 = M[c][0], v1 = M[c][1], v2 = M[c][2], v3 = M[c][3];
        const float sa = (l & 1) ? -1.0f : 1.0f, sb = -sa;
        inv[0][l] = ((v1 * f0 - v2 * f1) + v3 * f2) * sa;
        inv[1][l] = ((v0 * f0 - v2 * f3) + v3 * f4) * sb;
        inv[2][l] = ((v0 * f1 - v1 * f3) + v3 * f5) * sa;
        inv[3][l] = ((v0 * f2 - v1 * f4) + v2 * f5) * sb;
    }
    const float d0 = M[0][0] * inv[0][0], d1 = M[0][1] * inv[1][0], d2 = M[0][2] * inv[2][0], d3 = M[0][3] * inv[3][0];
    const float oneOverDet = 1.0f / ((d0 + d1) + (d2 + d3));
    // transpose, keep the upper-left 3x3: column j of the result is row j of the inverse
    Mat3f n;
    n.c0 = {inv[0][0] * oneOverDet, inv[1][0] * oneOverDet, inv[2][0] * oneOverDet};
    n.c1 = {inv[0][1] * oneOverDet, inv[1][1] * oneOverDet, inv[2][1] * oneOverDet};
    n.c2 = {inv[0][2] * oneOverDet, inv[1][2] * oneOverDet, inv[2][2] * oneOverDet};
    return n;
}

}  // namespace zillum
