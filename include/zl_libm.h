/* zl_libm.h — the transcendental functions of the shading path, pinned.
 *
 * The reference leaves sin / cos / atan / asin / log / pow to the GL driver (GLSL 4.50 §4.7.1:
 * precision is implementation-defined), so nothing in the reference fixes their bits.  This
 * header fixes them ONCE for every implementation of the path in this repository: the CUDA
 * kernels (csrc/, compiled --fmad=false), the CPU oracle (oracle/, -ffp-contract=off) and the
 * GL-emulation shim under which the reference's own GLSL text is compiled (oracle/ref_shim/).
 * Every function is written with IEEE-754 + - * / and comparisons only (binary32; zl_powf
 * goes through binary64), in one fixed evaluation order, with no FMA and no table, so all
 * three produce the same bits on an x86 host and on an sm_100a device.  sqrt and division are
 * correctly rounded on both sides already (nvcc -prec-sqrt/-prec-div default true) and stay
 * the hardware's.
 *
 * Used at: math.glsl:25-41 (toConcentricDisk), :58-71 (sphereToPlane / planeToSphere),
 * :133-137 (angleBetween), :180-185 (rotateZ); microfacet.glsl:92-112 (gtr1, gtr1SampleWm);
 * light.glsl:194 (envSampleWi); post_proc.glsl:57 (gamma).
 *
 * Accuracy (tests/test_libm.py, against binary64 libm): sin / cos < 2.5 ulp for |x| <= 12867
 * (argument reduction by a four-part pi/2 whose first three parts carry 11 bits each, exact
 * for |k| < 2^13), atan2 < 3.5 ulp, asin / acos < 3 ulp, log < 1 ulp, pow / exp <= 0.5 ulp + 2^-28.
 * Outside the documented range results stay deterministic but lose accuracy; |x| >= 2^20 or
 * a non-finite argument gives sin = x - x (0 or NaN), cos = 1 + (x - x).
 */
#ifndef ZL_LIBM_H
#define ZL_LIBM_H

#include <stdint.h>
#if defined(__CUDACC__)
#define ZL_LIBM_FN __host__ __device__ __forceinline__
#else
#include <math.h>
#include <string.h>
#define ZL_LIBM_FN static inline
#endif

ZL_LIBM_FN uint32_t zl_f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
ZL_LIBM_FN float zl_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
ZL_LIBM_FN uint64_t zl_d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
ZL_LIBM_FN double zl_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
ZL_LIBM_FN float zl_fabsf(float x) { return zl_u2f(zl_f2u(x) & 0x7fffffffu); }
ZL_LIBM_FN float zl_copysignf(float mag, float sgn) { return zl_u2f((zl_f2u(mag) & 0x7fffffffu) | (zl_f2u(sgn) & 0x80000000u)); }

/* ---- sin / cos ------------------------------------------------------------------------
 * k = nearest integer to x * 2/pi; r = x - k * pi/2 in four steps (Cody-Waite);
 * minimax polynomials on [-pi/4, pi/4] (the classic single-precision coefficient sets). */
ZL_LIBM_FN void zl_sincosf(float x, float* sOut, float* cOut) {
    if (!(zl_fabsf(x) < 1048576.0f)) { float z = x - x; *sOut = z; *cOut = 1.0f + z; return; }
    const float kf = rintf(x * 0.636619772367581343f);
    const int k = (int)kf;
    float r = x - kf * 1.5703125f;                    /* 0x1.92p+0   (11 bits) */
    r = r - kf * 0.0004837512969970703f;              /* 0x1.fb4p-12 (11 bits) */
    r = r - kf * 7.549533620476723e-08f;              /* 0x1.444p-24 (11 bits) */
    r = r - kf * 2.5633440682570896e-12f;             /* remainder of pi/2     */
    const float z = r * r;
    float s = -1.9515295891e-4f;
    s = s * z + 8.3321608736e-3f;
    s = s * z - 1.6666654611e-1f;
    s = s * z * r + r;
    float c = 2.443315711809948e-5f;
    c = c * z - 1.388731625493765e-3f;
    c = c * z + 4.166664568298827e-2f;
    c = c * z * z - 0.5f * z + 1.0f;
    switch (k & 3) {
    case 0: *sOut = s; *cOut = c; break;
    case 1: *sOut = c; *cOut = -s; break;
    case 2: *sOut = -s; *cOut = -c; break;
    default: *sOut = -c; *cOut = s; break;
    }
}
ZL_LIBM_FN float zl_sinf(float x) { float s, c; zl_sincosf(x, &s, &c); return s; }
ZL_LIBM_FN float zl_cosf(float x) { float s, c; zl_sincosf(x, &s, &c); return c; }

/* ---- atan / atan2 -------------------------------------------------------------------- */
/* t in [0, 1]: one reduction at tan(pi/8), degree-9 odd polynomial */
ZL_LIBM_FN float zl_atan01f(float t) {
    float base = 0.0f;
    if (t > 0.4142135623730950f) { base = 0.785398163397448309616f; t = (t - 1.0f) / (t + 1.0f); }
    const float z = t * t;
    float p = 8.05374449538e-2f;
    p = p * z - 1.38776856032e-1f;
    p = p * z + 1.99777106478e-1f;
    p = p * z - 3.33329491539e-1f;
    p = p * z * t + t;
    return base + p;
}
ZL_LIBM_FN float zl_atan2f(float y, float x) {
    if (x != x || y != y) return x + y;
    const float ax = zl_fabsf(x), ay = zl_fabsf(y);
    const bool xneg = (zl_f2u(x) >> 31) != 0;
    float r;
    if (ay == 0.0f) r = xneg ? 3.14159265358979323846f : 0.0f;
    else if (ax == ay) r = xneg ? 2.35619449019234492885f : 0.785398163397448309616f;   /* incl. inf, inf */
    else if (ax > ay) {
        r = zl_atan01f(ay / ax);                                                   /* finite / inf = 0 */
        if (xneg) r = 3.14159265358979323846f - r;
    } else {
        r = 1.57079632679489661923f - zl_atan01f(ax / ay);
        if (xneg) r = 3.14159265358979323846f - r;
    }
    return zl_copysignf(r, y);
}
ZL_LIBM_FN float zl_atanf(float x) { return zl_atan2f(x, 1.0f); }

/* ---- asin / acos --------------------------------------------------------------------- */
ZL_LIBM_FN float zl_asinf(float x) {
    const float a = zl_fabsf(x);
    if (!(a <= 1.0f)) return (x - x) / (x - x);       /* NaN outside [-1, 1] (GLSL: undefined) */
    float z, t; bool big = a > 0.5f;
    if (big) { z = 0.5f * (1.0f - a); t = sqrtf(z); } else { t = a; z = t * t; }
    float p = 4.2163199048e-2f;
    p = p * z + 2.4181311049e-2f;
    p = p * z + 4.5470025998e-2f;
    p = p * z + 7.4953002686e-2f;
    p = p * z + 1.6666752422e-1f;
    p = p * z * t + t;
    if (big) p = 1.57079632679489661923f - (p + p);
    return zl_copysignf(p, x);
}
ZL_LIBM_FN float zl_acosf(float x) {
    const float a = zl_fabsf(x);
    if (!(a > 0.5f)) return 1.57079632679489661923f - zl_asinf(x);             /* also NaN */
    if (a > 1.0f) return (x - x) / (x - x);
    const float z = 0.5f * (1.0f - a), t = sqrtf(z);
    float p = 4.2163199048e-2f;
    p = p * z + 2.4181311049e-2f;
    p = p * z + 4.5470025998e-2f;
    p = p * z + 7.4953002686e-2f;
    p = p * z + 1.6666752422e-1f;
    p = p * z * t + t;
    p = p + p;
    return x > 0.0f ? p : 3.14159265358979323846f - p;
}

/* ---- log ------------------------------------------------------------------------------
 * x = m * 2^e, m in [sqrt(1/2), sqrt(2)); degree-9 polynomial in (m - 1); e * ln2 in two parts. */
ZL_LIBM_FN float zl_logf(float x) {
    if (x != x) return x;
    if (x < 0.0f) return (x - x) / (x - x);
    if (x == 0.0f) return zl_u2f(0xff800000u);
    uint32_t u = zl_f2u(x);
    if (u == 0x7f800000u) return x;
    int e = 0;
    if (u < 0x00800000u) { x = x * 8388608.0f; u = zl_f2u(x); e = -23; }          /* subnormal */
    e += (int)(u >> 23) - 126;                                                  /* m in [0.5, 1) */
    float m = zl_u2f((u & 0x007fffffu) | 0x3f000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = m + m - 1.0f; } else m = m - 1.0f;
    const float z = m * m;
    float y = 7.0376836292e-2f;
    y = y * m - 1.1514610310e-1f;
    y = y * m + 1.1676998740e-1f;
    y = y * m - 1.2420140846e-1f;
    y = y * m + 1.4249322787e-1f;
    y = y * m - 1.6668057665e-1f;
    y = y * m + 2.0000714765e-1f;
    y = y * m - 2.4999993993e-1f;
    y = y * m + 3.3333331174e-1f;
    y = y * m * z;
    const float fe = (float)e;
    y = y + fe * -2.12194440e-4f;
    y = y - 0.5f * z;
    float r = m + y;
    r = r + fe * 0.693359375f;
    return r;
}

/* ---- pow (binary64 inside: ln by the atanh series, exp by Taylor; one rounding to binary32) */
ZL_LIBM_FN double zl_log_d(double x) {           /* x finite, > 0, normal in binary64 */
    const uint64_t u = zl_d2u(x);
    int e = (int)(u >> 52) - 1022;
    double m = zl_u2d((u & 0x000fffffffffffffull) | 0x3fe0000000000000ull);     /* [0.5, 1) */
    if (m < 0.70710678118654752) { e -= 1; m = m + m; }                          /* [sqrt(1/2), sqrt(2)) */
    const double s = (m - 1.0) / (m + 1.0), s2 = s * s;
    double p = 1.0 / 21.0;
    p = p * s2 + 1.0 / 19.0;
    p = p * s2 + 1.0 / 17.0;
    p = p * s2 + 1.0 / 15.0;
    p = p * s2 + 1.0 / 13.0;
    p = p * s2 + 1.0 / 11.0;
    p = p * s2 + 1.0 / 9.0;
    p = p * s2 + 1.0 / 7.0;
    p = p * s2 + 1.0 / 5.0;
    p = p * s2 + 1.0 / 3.0;
    p = p * s2 * s + s;
    const double de = (double)e;
    return (p + p) + de * 1.9082149292705877e-10 + de * 0.6931471803691238;
}
ZL_LIBM_FN double zl_exp_d(double t) {           /* |t| <= 200 */
    const double kd = rint(t * 1.4426950408889634);
    const int k = (int)kd;
    double r = t - kd * 0.6931471803691238;
    r = r - kd * 1.9082149292705877e-10;
    double p = 1.0 / 6227020800.0;               /* 1/13! */
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    return p * zl_u2d((uint64_t)(k + 1023) << 52);                               /* |k| <= 289 */
}
ZL_LIBM_FN float zl_powf(float x, float y) {
    if (y == 0.0f || x == 1.0f) return 1.0f;
    if (x != x || y != y) return x + y;
    if (x < 0.0f) return (x - x) / (x - x);       /* GLSL pow: undefined for x < 0 */
    const float inf = zl_u2f(0x7f800000u);
    if (x == 0.0f) return y > 0.0f ? 0.0f : inf;
    if (x == inf) return y > 0.0f ? inf : 0.0f;
    if (y == inf) return x > 1.0f ? inf : 0.0f;
    if (y == -inf) return x > 1.0f ? 0.0f : inf;
    const double t = (double)y * zl_log_d((double)x);                            /* binary32 x is normal in binary64 */
    if (t > 89.0) return inf;
    if (t < -104.0) return 0.0f;
    return (float)zl_exp_d(t);
}
ZL_LIBM_FN float zl_expf(float x) {
    if (x != x) return x;
    if (x > 89.0f) return zl_u2f(0x7f800000u);
    if (x < -104.0f) return 0.0f;
    return (float)zl_exp_d((double)x);
}

#endif /* ZL_LIBM_H */
