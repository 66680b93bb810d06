/*
 * sp_detmath.h -- portable, bit-reproducible arithmetic used where a floating-point result decides an
 * integer output (inlier masks, plane labels, extreme-point indices).
 *
 * Everything here is built from IEEE-754 +,-,*,/ and sqrt only, evaluated in a fixed order, so the same
 * source gives the same bits under gcc (-ffp-contract=off) and nvcc (-fmad=false, default -prec-div/-prec-sqrt).
 * libm's atan2/sin/cos are NOT bit-identical between glibc and CUDA, which is why the closed-form 3x3
 * eigen-solve of PCL (pcl::eigen33 / pcl::computeRoots, upstream common/impl/eigen.hpp -- the solver behind
 * SACSegmentation's optimizeCoefficients, called from
 * /root/reference/modules/stairperception/StairPerception.cpp:785 and :862) is restated on top of det_* functions.
 * The oracle can also run the same solve on libm ("faithful" mode) to show both agree to the last few ulps.
 *
 * Header-only; included by the CUDA library (device + host) and by the CPU oracle.
 */
#ifndef SP_DETMATH_H_
#define SP_DETMATH_H_

#include <math.h>
#include <float.h>

#ifdef __CUDACC__
#define SP_HD __host__ __device__ __forceinline__
#else
#define SP_HD static inline
#endif

/* ---- double-precision kernels (basic ops only) ------------------------------------------------------- */

/* atan(t) for t >= 0, by two half-angle reductions and a Taylor series (|t| <= 0.199 after reduction). */
SP_HD double det_atan_pos(double t)
{
    int inv = 0;
    if (t > 1.0) { t = 1.0 / t; inv = 1; }
    t = t / (1.0 + sqrt(1.0 + t * t));
    t = t / (1.0 + sqrt(1.0 + t * t));
    const double t2 = t * t;
    /* sum_{k=0}^{15} (-1)^k t^(2k+1)/(2k+1), Horner from the tail */
    double s = 1.0 / 31.0;
    s = 1.0 / 29.0 - t2 * s;
    s = 1.0 / 27.0 - t2 * s;
    s = 1.0 / 25.0 - t2 * s;
    s = 1.0 / 23.0 - t2 * s;
    s = 1.0 / 21.0 - t2 * s;
    s = 1.0 / 19.0 - t2 * s;
    s = 1.0 / 17.0 - t2 * s;
    s = 1.0 / 15.0 - t2 * s;
    s = 1.0 / 13.0 - t2 * s;
    s = 1.0 / 11.0 - t2 * s;
    s = 1.0 / 9.0 - t2 * s;
    s = 1.0 / 7.0 - t2 * s;
    s = 1.0 / 5.0 - t2 * s;
    s = 1.0 / 3.0 - t2 * s;
    s = 1.0 - t2 * s;
    double a = 4.0 * (t * s);
    if (inv) a = 1.5707963267948966 - a;
    return a;
}

/* atan2(y, x) for y >= 0 (the only case the eigen-solve needs: y = sqrt(-q)). Result in [0, pi]. */
SP_HD double det_atan2_ypos(double y, double x)
{
    if (x > 0.0) return det_atan_pos(y / x);
    if (x < 0.0) return 3.141592653589793 - det_atan_pos(y / (-x));
    return (y > 0.0) ? 1.5707963267948966 : 0.0;
}

/* cos and sin for |x| <= ~1.1 (theta = atan2(..)/3 lies in [0, pi/3]); Taylor, Horner from the tail. */
SP_HD double det_cos_small(double x)
{
    /* cos x = 1 - x^2/(1*2) (1 - x^2/(3*4) (1 - x^2/(5*6) ( ... ))) */
    const double x2 = x * x;
    double s = 1.0;
    for (int k = 13; k >= 0; --k) s = 1.0 - (x2 * s) / (double)((2 * k + 1) * (2 * k + 2));
    return s;
}

SP_HD double det_sin_small(double x)
{
    /* sin x = x (1 - x^2/(2*3) (1 - x^2/(4*5) ( ... ))) */
    const double x2 = x * x;
    double s = 1.0;
    for (int k = 13; k >= 1; --k) s = 1.0 - (x2 * s) / (double)((2 * k) * (2 * k + 1));
    return x * s;
}


/* ---- log and acos from basic operations --------------------------------------------------------------- */
/* ln(x) for finite x > 0: x = m * 2^e with m in [sqrt(1/2), sqrt(2)); ln m = 2 atanh((m-1)/(m+1)).
 * Used for RANSAC's k = log(1-p)/log(1-w^3) (pcl::RandomSampleConsensus::computeModel). */
SP_HD double det_log(double x)
{
    long long bits;
#ifdef __CUDA_ARCH__
    bits = __double_as_longlong(x);
#else
    { union { double d; long long i; } u; u.d = x; bits = u.i; }
#endif
    int e = (int)((bits >> 52) & 0x7ff);
    if (e == 0) { /* subnormal: scale up by 2^54 */
        x = x * 18014398509481984.0;
#ifdef __CUDA_ARCH__
        bits = __double_as_longlong(x);
#else
        { union { double d; long long i; } u; u.d = x; bits = u.i; }
#endif
        e = (int)((bits >> 52) & 0x7ff) - 54;
    }
    e -= 1023;
    bits = (bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL;
    double m;
#ifdef __CUDA_ARCH__
    m = __longlong_as_double(bits);
#else
    { union { double d; long long i; } u; u.i = bits; m = u.d; }
#endif
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    const double s = (m - 1.0) / (m + 1.0);
    const double s2 = s * s;
    double t = 1.0 / 27.0;
    t = 1.0 / 25.0 + s2 * t;
    t = 1.0 / 23.0 + s2 * t;
    t = 1.0 / 21.0 + s2 * t;
    t = 1.0 / 19.0 + s2 * t;
    t = 1.0 / 17.0 + s2 * t;
    t = 1.0 / 15.0 + s2 * t;
    t = 1.0 / 13.0 + s2 * t;
    t = 1.0 / 11.0 + s2 * t;
    t = 1.0 / 9.0 + s2 * t;
    t = 1.0 / 7.0 + s2 * t;
    t = 1.0 / 5.0 + s2 * t;
    t = 1.0 / 3.0 + s2 * t;
    t = 1.0 + s2 * t;
    return (double)e * 0.6931471805599453 + 2.0 * (s * t);
}

/* acos(x) = atan2(sqrt((1-x)(1+x)), x); NaN for |x| > 1 exactly like libm.  Result in [0, pi]. */
SP_HD double det_acos(double x)
{
    const double y2 = (1.0 - x) * (1.0 + x);
    if (!(y2 >= 0.0)) return (double)NAN;
    return det_atan2_ypos(sqrt(y2), x);
}
SP_HD float det_acosf(float x) { return (float)det_acos((double)x); }

/* ---- float front-ends: evaluate in double, round once ----------------------------------------------- */
SP_HD float det_atan2f_ypos(float y, float x) { return (float)det_atan2_ypos((double)y, (double)x); }
SP_HD float det_cosf_small(float x) { return (float)det_cos_small((double)x); }
SP_HD float det_sinf_small(float x) { return (float)det_sin_small((double)x); }

/* ---- pcl::computeRoots2 / computeRoots / eigen33 restated in float (Scalar = float) ------------------ */
/* USE_LIBM selects libm trig (oracle "faithful" mode, host only); 0 = portable det_* functions. */

SP_HD void det_roots2f(float b, float c, float *roots)
{
    roots[0] = 0.0f;
    float d = (float)((double)(b * b) - 4.0 * (double)c);     /* Scalar (b*b - 4.0*c): double expression narrowed */
    if (d < 0.0f) d = 0.0f;
    float sd = sqrtf(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}

template <int USE_LIBM>
SP_HD void det_roots3f(const float *m /* row-major 3x3 symmetric */, float *roots)
{
    const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
    float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
    float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
    float c2 = m00 + m11 + m22;
    if (fabsf(c0) < FLT_EPSILON) { det_roots2f(c2, c1, roots); return; }
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = sqrtf(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    float rho = sqrtf(-a_over_3);
    float theta, cos_theta, sin_theta;
#ifndef __CUDA_ARCH__
    if (USE_LIBM) {
        theta = atan2f(sqrtf(-q), half_b) * s_inv3;
        cos_theta = cosf(theta);
        sin_theta = sinf(theta);
    } else
#endif
    {
        theta = det_atan2f_ypos(sqrtf(-q), half_b) * s_inv3;
        cos_theta = det_cosf_small(theta);
        sin_theta = det_sinf_small(theta);
    }
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    float t;
    if (roots[0] >= roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
    if (roots[1] >= roots[2]) {
        t = roots[1]; roots[1] = roots[2]; roots[2] = t;
        if (roots[0] >= roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
    }
    if (roots[0] <= 0.0f) det_roots2f(c2, c1, roots);
}

/* smallest eigenpair of a symmetric 3x3 (pcl::eigen33(mat, eigenvalue, eigenvector)) */
template <int USE_LIBM>
SP_HD void det_eigen33f(const float *mat /* row-major 3x3 */, float *eigenvalue, float *evec)
{
    float scale = 0.0f;
    for (int i = 0; i < 9; ++i) { float a = fabsf(mat[i]); if (a > scale) scale = a; }
    if (scale <= FLT_MIN) scale = 1.0f;
    float s[9];
    for (int i = 0; i < 9; ++i) s[i] = mat[i] / scale;
    float roots[3];
    det_roots3f<USE_LIBM>(s, roots);
    *eigenvalue = roots[0] * scale;
    s[0] -= roots[0]; s[4] -= roots[0]; s[8] -= roots[0];
    /* rows r0 = s[0..2], r1 = s[3..5], r2 = s[6..8] */
    float v1[3] = { s[1] * s[5] - s[2] * s[4], s[2] * s[3] - s[0] * s[5], s[0] * s[4] - s[1] * s[3] };   /* r0 x r1 */
    float v2[3] = { s[1] * s[8] - s[2] * s[7], s[2] * s[6] - s[0] * s[8], s[0] * s[7] - s[1] * s[6] };   /* r0 x r2 */
    float v3[3] = { s[4] * s[8] - s[5] * s[7], s[5] * s[6] - s[3] * s[8], s[3] * s[7] - s[4] * s[6] };   /* r1 x r2 */
    float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
    float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
    float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
    const float *v; float l;
    if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
    else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
    else { v = v3; l = l3; }
    float n = sqrtf(l);
    evec[0] = v[0] / n; evec[1] = v[1] / n; evec[2] = v[2] / n;
}

/* ---- symmetric 3x3 eigen-decomposition, cyclic Jacobi, float, fixed sweep count ---------------------- */
/* Stands in for Eigen::JacobiSVD of the (symmetric PSD) covariance at StairPerception.cpp:1241: singular
 * values = eigenvalues (descending), U = eigenvectors (columns).  evecs is row-major, column j = vector j. */
SP_HD void det_jacobi3f(const float *cov /* row-major */, float *evals, float *evecs)
{
    float a[3][3] = { { cov[0], cov[1], cov[2] }, { cov[1], cov[4], cov[5] }, { cov[2], cov[5], cov[8] } };
    float v[3][3] = { { 1.f, 0.f, 0.f }, { 0.f, 1.f, 0.f }, { 0.f, 0.f, 1.f } };
    for (int sweep = 0; sweep < 12; ++sweep) {
        for (int pi = 0; pi < 3; ++pi) {
            const int p = (pi == 2) ? 1 : 0;
            const int q = (pi == 0) ? 1 : 2;
            const float apq = a[p][q];
            if (apq == 0.0f) continue;
            const float theta = (a[q][q] - a[p][p]) / (2.0f * apq);
            const float at = fabsf(theta);
            float t = 1.0f / (at + sqrtf(theta * theta + 1.0f));
            if (theta < 0.0f) t = -t;
            const float c = 1.0f / sqrtf(t * t + 1.0f);
            const float s = t * c;
            const int r = 3 - p - q;
            const float app = a[p][p], aqq = a[q][q];
            a[p][p] = app - t * apq;
            a[q][q] = aqq + t * apq;
            a[p][q] = 0.0f; a[q][p] = 0.0f;
            const float arp = a[r][p], arq = a[r][q];
            a[r][p] = c * arp - s * arq; a[p][r] = a[r][p];
            a[r][q] = s * arp + c * arq; a[q][r] = a[r][q];
            for (int k = 0; k < 3; ++k) {
                const float vkp = v[k][p], vkq = v[k][q];
                v[k][p] = c * vkp - s * vkq;
                v[k][q] = s * vkp + c * vkq;
            }
        }
    }
    int o0 = 0, o1 = 1, o2 = 2, t;
    float d[3] = { a[0][0], a[1][1], a[2][2] };
    if (d[o0] < d[o1]) { t = o0; o0 = o1; o1 = t; }
    if (d[o1] < d[o2]) { t = o1; o1 = o2; o2 = t; if (d[o0] < d[o1]) { t = o0; o0 = o1; o1 = t; } }
    const int ord[3] = { o0, o1, o2 };
    for (int j = 0; j < 3; ++j) {
        evals[j] = d[ord[j]];
        for (int k = 0; k < 3; ++k) evecs[k * 3 + j] = v[k][ord[j]];
    }
}

#endif /* SP_DETMATH_H_ */
