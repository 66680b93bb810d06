/*
 * indep_math.h -- the oracle's OWN double-precision numerics.   TEST INFRASTRUCTURE ONLY.
 *
 * The product and the oracle's default mode share include/sp_detmath.h (one source compiled by nvcc and g++), which makes
 * their bit-exact agreement blind to a wrong restatement of pcl::eigen33 / Eigen::JacobiSVD / acos / log.  Nothing in this
 * file shares code, structure or algorithm with that header:
 *   - symmetric 3x3 eigen-decomposition: classical Jacobi in double, pivot = largest off-diagonal element, rotation
 *     angle from libm atan2, iterated to convergence (the shared header has a closed-form trigonometric root finder in float
 *     for the smallest pair, and a fixed-sweep cyclic float Jacobi with the tangent formula for the full decomposition);
 *   - acos / log: libm.
 * Two uses (oracle/sp_oracle.cpp): the AUDIT, which re-solves every eigen problem, acos and log of a frame independently
 * and records the largest deviation of the shared-header result, and eig_math = 2, which runs the whole pipeline on these
 * routines so that the product can be compared with it at the stated tolerance (1e-4).
 */
#ifndef ORC_INDEP_MATH_H_
#define ORC_INDEP_MATH_H_

#include <cmath>
#include <cstring>

namespace indep {

/* A = V diag(lam) V^T for symmetric A given by its upper triangle {xx, xy, xz, yy, yz, zz}.
 * lam ascending, vec[k] = unit eigenvector of lam[k]. */
inline void sym3_eigen(const double up[6], double lam[3], double vec[3][3])
{
    double A[3][3] = { { up[0], up[1], up[2] }, { up[1], up[3], up[4] }, { up[2], up[4], up[5] } };
    double V[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
    double total = 0.0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) total += A[i][j] * A[i][j];
    for (int it = 0; it < 200; ++it) {
        int p = 0, q = 1;
        double big = std::fabs(A[0][1]);
        if (std::fabs(A[0][2]) > big) { big = std::fabs(A[0][2]); p = 0; q = 2; }
        if (std::fabs(A[1][2]) > big) { big = std::fabs(A[1][2]); p = 1; q = 2; }
        if (big * big <= 1e-34 * total || big == 0.0) break;
        const double phi = 0.5 * std::atan2(2.0 * A[p][q], A[q][q] - A[p][p]);
        const double cs = std::cos(phi), sn = std::sin(phi);
        /* A <- G^T A G with G = rotation in the (p, q) plane */
        for (int k = 0; k < 3; ++k) {
            const double akp = A[k][p], akq = A[k][q];
            A[k][p] = cs * akp - sn * akq; A[k][q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < 3; ++k) {
            const double apk = A[p][k], aqk = A[q][k];
            A[p][k] = cs * apk - sn * aqk; A[q][k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < 3; ++k) {
            const double vkp = V[k][p], vkq = V[k][q];
            V[k][p] = cs * vkp - sn * vkq; V[k][q] = sn * vkp + cs * vkq;
        }
    }
    int ord[3] = { 0, 1, 2 };
    for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 3; ++j) if (A[ord[j]][ord[j]] < A[ord[i]][ord[i]]) { const int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    for (int k = 0; k < 3; ++k) {
        lam[k] = A[ord[k]][ord[k]];
        double n = 0.0;
        for (int r = 0; r < 3; ++r) n += V[r][ord[k]] * V[r][ord[k]];
        n = std::sqrt(n);
        for (int r = 0; r < 3; ++r) vec[k][r] = V[r][ord[k]] / n;
    }
}

/* angle (radians) between two directions, sign of either ignored */
inline double axis_angle(const double a[3], const double b[3])
{
    const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
    const double s = std::sqrt(cx * cx + cy * cy + cz * cz), c = std::fabs(a[0] * b[0] + a[1] * b[1] + a[2] * b[2]);
    return std::atan2(s, c);
}

/* what the audit keeps per frame */
struct Audit {
    long n_eigen33 = 0, n_jacobi = 0, n_acos = 0, n_log = 0;
    double eigen33_max_angle = 0.0;          /* smallest-eigenvector direction, radians */
    double eigen33_max_angle_wellsep = 0.0;  /* the same over matrices whose two smallest eigenvalues are well separated */
    double eigen33_max_val = 0.0;            /* |lambda0 - ref| / lambda_max */
    double jacobi_max_val = 0.0;             /* max |eval - ref| / lambda_max */
    double jacobi_max_angle = 0.0;           /* eigenvector directions (pairs separated by > 1e-3 lambda_max only) */
    double acos_max_abs = 0.0, log_max_rel = 0.0;
    void reset() { *this = Audit(); }
};

} // namespace indep
#endif
