/*
 * sp_oracle.cpp -- CPU ORACLE for the stair-perception segmentation front-end.   TEST INFRASTRUCTURE ONLY.
 *
 * Nothing here is shipped or called by the product (stair_perception_b200 / libsp_b200.so).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load liborc.so.
 *
 * It restates, stage by stage and without PCL / Eigen / Boost (none exist in this image), the reference path
 *   StairDetection::process()   /root/reference/modules/stairperception/StairPerception.cpp:5-214
 * and the PCL 1.7.2-1.8.1 algorithms that path calls (recalled from upstream; the upstream file each rule
 * comes from is named at the function).  PARITY UNPINNED at stage level: the reference ships no tests, golden
 * vectors or fixtures for any intermediate stage and PCL cannot be built here.  External anchor: the published
 * 4-step table for samples/0001 (pic/screenshot2.png), checked at mm level by tests/test_oracle_anchor.py.
 *
 * Build: g++ -O2 -ffp-contract=off (no FMA contraction: the reference is a plain x86-64 build).
 */
#include "sp_oracle.h"
#include "../include/sp_detmath.h"
#include "indep_math.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <random>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

const float kNaN = std::numeric_limits<float>::quiet_NaN();

struct Tap { std::vector<uint8_t> bytes; };

struct SegPlane {                 /* one extracted plane: coefficients + pixel indices in cloud order */
    float coef[4];
    std::vector<int> idx;         /* pixel indices */
};

enum { ST_LEVEL = 0, ST_NORMALS, ST_CROP, ST_VOXEL, ST_FILTER, ST_SEG_H, ST_SEG_V, ST_MERGE, ST_INFO, ST_COUNT };
const char *kStageNames[ST_COUNT] = { "level_mask", "normals", "crop", "voxel", "filter_split", "segment_h",
                                      "segment_v", "merge", "plane_info_sort" };

} // namespace

struct orc_ctx {
    orc_params p;
    int normal_arith = 1;    /* 0 = PCL integral images (double), 1 = fixed-order double box sums */
    int voxel_order = 0;     /* 0 = stable, 1 = std::sort on the key only */
    int eig_math = 0;        /* 0 = portable det_* math (shared with the product), 1 = the same restatement on libm trig,
                                2 = the oracle's own numerics (oracle/indep_math.h: double Jacobi, libm acos / log) */
    bool audit_on = false;   /* set by orc_process(keep_taps = 1): the timed baseline does not pay for the audit */
    indep::Audit audit;      /* every eigen-solve, acos and log of the last frame re-done independently: largest deviations */
    int knn_threads = 8;     /* host threads of the kNN normal search (a3') */
    int acos_float = 0;      /* a8: 0 = acos((double)nx), 1 = (double)acosf(nx)  (overload ambiguity) */
    std::map<std::string, Tap> taps;
    double stage_ms[ST_COUNT];
    /* working state of the current frame */
    int W = 0, H = 0, N = 0;
    std::vector<float> x, y, z, nx, ny, nz;
    std::vector<orc_plane_record> records;
    std::vector<int> plane_idx;
    uint32_t rand_pos = 0;   /* position in the glibc rand() stream for this frame */
    std::vector<int32_t> rand_stream;
};

namespace {

template <class T> void put_tap(orc_ctx *c, const char *name, const T *data, size_t n)
{
    Tap &t = c->taps[name];
    t.bytes.resize(n * sizeof(T));
    if (n) memcpy(t.bytes.data(), data, n * sizeof(T));
}
template <class T> void put_tap(orc_ctx *c, const char *name, const std::vector<T> &v) { put_tap(c, name, v.data(), v.size()); }

inline bool finite3(float a, float b, float c) { return std::isfinite(a) && std::isfinite(b) && std::isfinite(c); }

/* ---- the places where sp_detmath.h decides a result, each with its independent twin ------------------------------------ */
inline void cov_upper(const float *cov9, double up[6]) { up[0] = cov9[0]; up[1] = cov9[1]; up[2] = cov9[2]; up[3] = cov9[4]; up[4] = cov9[5]; up[5] = cov9[8]; }

/* smallest eigenpair of the refit / kNN covariance (pcl::eigen33).  mode 2 solves it with indep::sym3_eigen; the sign of an
 * eigenvector is free, mode 2 makes its largest component positive.  audit (may be NULL) gets the deviation of modes 0 / 1. */
void smallest_eigenpair(int mode, const float *cov9, float *ev, float *evec, indep::Audit *audit)
{
    double up[6], lam[3], vec[3][3];
    if (mode == 2 || audit) { cov_upper(cov9, up); indep::sym3_eigen(up, lam, vec); }
    if (mode == 2) {
        int big = 0;
        for (int k = 1; k < 3; ++k) if (std::fabs(vec[0][k]) > std::fabs(vec[0][big])) big = k;
        const double sg = vec[0][big] < 0 ? -1.0 : 1.0;
        *ev = (float)lam[0];
        for (int k = 0; k < 3; ++k) evec[k] = (float)(sg * vec[0][k]);
        return;
    }
    if (mode == 1) det_eigen33f<1>(cov9, ev, evec); else det_eigen33f<0>(cov9, ev, evec);
    if (audit && std::isfinite(evec[0]) && lam[2] > 0.0) {
        const double got[3] = { evec[0], evec[1], evec[2] };
        const double ang = indep::axis_angle(got, vec[0]);
        audit->n_eigen33++;
        audit->eigen33_max_angle = std::max(audit->eigen33_max_angle, ang);
        if (lam[1] - lam[0] > 0.05 * lam[2]) audit->eigen33_max_angle_wellsep = std::max(audit->eigen33_max_angle_wellsep, ang);
        audit->eigen33_max_val = std::max(audit->eigen33_max_val, std::fabs((double)*ev - lam[0]) / lam[2]);
    }
}

/* full decomposition of the plane covariance (Eigen::JacobiSVD of a symmetric PSD matrix): descending, vectors as columns */
void full_eigen(int mode, const float *cov9, float *evals, float *evecs, indep::Audit *audit)
{
    double up[6], lam[3], vec[3][3];
    if (mode == 2 || audit) { cov_upper(cov9, up); indep::sym3_eigen(up, lam, vec); }
    if (mode == 2) {
        for (int j = 0; j < 3; ++j) {
            const int k = 2 - j;
            int big = 0;
            for (int r = 1; r < 3; ++r) if (std::fabs(vec[k][r]) > std::fabs(vec[k][big])) big = r;
            const double sg = vec[k][big] < 0 ? -1.0 : 1.0;
            evals[j] = (float)lam[k];
            for (int r = 0; r < 3; ++r) evecs[r * 3 + j] = (float)(sg * vec[k][r]);
        }
        return;
    }
    det_jacobi3f(cov9, evals, evecs);
    if (audit && lam[2] > 0.0) {
        audit->n_jacobi++;
        for (int j = 0; j < 3; ++j) {
            const int k = 2 - j;
            audit->jacobi_max_val = std::max(audit->jacobi_max_val, std::fabs((double)evals[j] - lam[k]) / lam[2]);
            const double gap = std::min(k > 0 ? lam[k] - lam[k - 1] : 1e300, k < 2 ? lam[k + 1] - lam[k] : 1e300);
            if (gap > 1e-3 * lam[2]) {
                const double got[3] = { evecs[0 * 3 + j], evecs[1 * 3 + j], evecs[2 * 3 + j] };
                audit->jacobi_max_angle = std::max(audit->jacobi_max_angle, indep::axis_angle(got, vec[k]));
            }
        }
    }
}

inline float merge_acos(int mode, float dot, indep::Audit *audit)
{
    const float ref = (float)acos((double)dot);
    if (mode != 0) return ref;
    const float a = det_acosf(dot);
    if (audit && std::isfinite(ref)) { audit->n_acos++; audit->acos_max_abs = std::max(audit->acos_max_abs, (double)std::fabs(a - ref)); }
    return a;
}

inline double ransac_log(int mode, double x, indep::Audit *audit)
{
    const double ref = log(x);
    if (mode != 0) return ref;
    const double v = det_log(x);
    if (audit && ref != 0.0) { audit->n_log++; audit->log_max_rel = std::max(audit->log_max_rel, std::fabs((v - ref) / ref)); }
    return v;
}

/* pcl::deg2rad(float), upstream common/impl/angles.hpp */
inline float deg2radf(float a) { return a * 0.017453293f; }

/* ---------------------------------------------------------------------------------------------------- */
/* a1: levelling, /root/reference/test/reader.cpp:273-324.  Stateless variant: a NaN-x input yields NaN xyz
 * (the reference leaves the previous frame's value in its reused output buffer).                        */
void level_and_mask(orc_ctx *c, const uint8_t *rec, const float *R)
{
    const int N = c->N;
    c->x.resize(N); c->y.resize(N); c->z.resize(N);
    for (int i = 0; i < N; ++i) {
        float px, py, pz; uint32_t rgba;
        memcpy(&px, rec + 16 * (size_t)i, 4); memcpy(&py, rec + 16 * (size_t)i + 4, 4);
        memcpy(&pz, rec + 16 * (size_t)i + 8, 4); memcpy(&rgba, rec + 16 * (size_t)i + 12, 4);
        float ox = px, oy = py, oz = pz;
        if (R) {
            if (!std::isnan(px)) {
                ox = R[0] * px + R[1] * py + R[2] * pz;      /* left-to-right float, reader.cpp:307-318 */
                oy = R[3] * px + R[4] * py + R[5] * pz;
                oz = R[6] * px + R[7] * py + R[8] * pz;
            } else { ox = oy = oz = kNaN; }
        }
        /* a2: remove_nan_points, StairPerception.cpp:338-356.  PointXYZRGBA packs b,g,r,a little-endian. */
        const uint32_t b = rgba & 255u, g = (rgba >> 8) & 255u, r = (rgba >> 16) & 255u;
        if (r == 0 || g == 0 || b == 0) { ox = oy = oz = kNaN; }
        c->x[i] = ox; c->y[i] = oy; c->z[i] = oz;
    }
}

/* ---------------------------------------------------------------------------------------------------- */
/* a3: pcl::IntegralImageNormalEstimation, AVERAGE_3D_GRADIENT, MaxDepthChangeFactor 0.02, NormalSmoothingSize 5,
 * BORDER_POLICY_IGNORE, no depth-dependent smoothing (StairPerception.cpp:579-592; upstream
 * features/impl/integral_image_normal.hpp, features/impl/integral_image2D.hpp).                          */
void normals_integral(orc_ctx *c)
{
    const int W = c->W, H = c->H, N = c->N;
    const float *X = c->x.data(), *Y = c->y.data(), *Z = c->z.data();
    c->nx.assign(N, kNaN); c->ny.assign(N, kNaN); c->nz.assign(N, kNaN);
    if (W < 3 || H < 3) return;                      /* nothing can lie 5 pixels inside the border; the loops below need 3 x 3 */

    /* depth-change map (on the levelled z = lateral axis: the reference's quirk) */
    std::vector<uint8_t> dcm(N, 255);
    for (int ri = 0; ri < H - 1; ++ri)
        for (int ci = 0; ci < W - 1; ++ci) {
            const int i = ri * W + ci;
            const float d = Z[i], dR = Z[i + 1], dD = Z[i + W];
            const float thr = (0.02f * (fabsf(d) + 1.0f) * 2.0f);
            if (fabsf(d - dR) > thr || !std::isfinite(d) || !std::isfinite(dR)) { dcm[i] = 0; dcm[i + 1] = 0; }
            if (fabsf(d - dD) > thr || !std::isfinite(d) || !std::isfinite(dD)) { dcm[i] = 0; dcm[i + W] = 0; }
        }
    /* distance map: two-pass 1.0 / 1.4 chamfer */
    std::vector<float> dist(N);
    for (int i = 0; i < N; ++i) dist[i] = dcm[i] == 0 ? 0.0f : (float)(W + H);
    for (int ri = 1; ri < H; ++ri) {
        float *prev = &dist[(size_t)(ri - 1) * W], *cur = &dist[(size_t)ri * W];
        for (int ci = 1; ci < W - 1; ++ci) {
            const float upLeft = prev[ci - 1] + 1.4f, up = prev[ci] + 1.0f, upRight = prev[ci + 1] + 1.4f;
            const float left = cur[ci - 1] + 1.0f, center = cur[ci];
            const float m = std::min(std::min(upLeft, up), std::min(left, upRight));
            if (m < center) cur[ci] = m;
        }
    }
    for (int ri = H - 2; ri >= 0; --ri) {
        float *next = &dist[(size_t)(ri + 1) * W], *cur = &dist[(size_t)ri * W];
        for (int ci = W - 2; ci >= 1; --ci) {
            const float lowerLeft = next[ci - 1] + 1.4f, lower = next[ci] + 1.0f, lowerRight = next[ci + 1] + 1.4f;
            const float right = cur[ci + 1] + 1.0f, center = cur[ci];
            const float m = std::min(std::min(lowerLeft, lower), std::min(right, lowerRight));
            if (m < center) cur[ci] = m;
        }
    }
    /* gradient fields (float, zero outside [1,H-2]x[1,W-2]) */
    std::vector<float> gx((size_t)N * 3, 0.0f), gy((size_t)N * 3, 0.0f);
    for (int ri = 1; ri < H - 1; ++ri)
        for (int ci = 1; ci < W - 1; ++ci) {
            const int i = ri * W + ci;
            gx[3 * (size_t)i + 0] = X[i + 1] - X[i - 1]; gx[3 * (size_t)i + 1] = Y[i + 1] - Y[i - 1]; gx[3 * (size_t)i + 2] = Z[i + 1] - Z[i - 1];
            gy[3 * (size_t)i + 0] = X[i + W] - X[i - W]; gy[3 * (size_t)i + 1] = Y[i + W] - Y[i - W]; gy[3 * (size_t)i + 2] = Z[i + W] - Z[i - W];
        }
    auto elem_finite = [](const float *e) { return std::isfinite(e[0] + (e[1] + e[2])); };

    std::vector<double> IX, IY; std::vector<unsigned> CX, CY;
    const int W1 = W + 1;
    if (c->normal_arith == 0) {
        auto build = [&](const std::vector<float> &g, std::vector<double> &I, std::vector<unsigned> &C) {
            I.assign((size_t)(H + 1) * W1 * 3, 0.0); C.assign((size_t)(H + 1) * W1, 0u);
            for (int r = 0; r < H; ++r) {
                double *prev = &I[(size_t)r * W1 * 3], *cur = &I[(size_t)(r + 1) * W1 * 3];
                unsigned *cprev = &C[(size_t)r * W1], *ccur = &C[(size_t)(r + 1) * W1];
                cur[0] = cur[1] = cur[2] = 0.0; ccur[0] = 0;
                for (int col = 0; col < W; ++col) {
                    for (int k = 0; k < 3; ++k)
                        cur[3 * (col + 1) + k] = prev[3 * (col + 1) + k] + cur[3 * col + k] - prev[3 * col + k];
                    ccur[col + 1] = cprev[col + 1] + ccur[col] - cprev[col];
                    const float *e = &g[3 * ((size_t)r * W + col)];
                    if (elem_finite(e)) {
                        for (int k = 0; k < 3; ++k) cur[3 * (col + 1) + k] += (double)e[k];
                        ++ccur[col + 1];
                    }
                }
            }
        };
        build(gx, IX, CX); build(gy, IY, CY);
    }

    const int border = 5;
    for (int ri = border; ri < H - border; ++ri)
        for (int ci = border; ci < W - border; ++ci) {
            const int i = ri * W + ci;
            if (!std::isfinite(Z[i])) continue;
            const float s = std::min(dist[i], 5.0f);
            if (!(s > 2.0f)) continue;
            const int k = (int)s, sx = ci - k / 2, sy = ri - k / 2;
            double sgx[3], sgy[3]; unsigned cntx = 0, cnty = 0;
            if (c->normal_arith == 0) {
                const size_t ul = (size_t)sy * W1 + sx, ur = ul + k, ll = (size_t)(sy + k) * W1 + sx, lr = ll + k;
                cntx = CX[lr] + CX[ul] - CX[ur] - CX[ll];
                cnty = CY[lr] + CY[ul] - CY[ur] - CY[ll];
                for (int q = 0; q < 3; ++q) {
                    sgx[q] = IX[3 * lr + q] + IX[3 * ul + q] - IX[3 * ur + q] - IX[3 * ll + q];
                    sgy[q] = IY[3 * lr + q] + IY[3 * ul + q] - IY[3 * ur + q] - IY[3 * ll + q];
                }
            } else {
                /* fixed-order definition the CUDA kernel reproduces: S = sum over rows (top to bottom) of the
                 * row sums (left to right), all in double, non-finite elements contributing +0.0 */
                for (int q = 0; q < 3; ++q) sgx[q] = sgy[q] = 0.0;
                for (int rr = 0; rr < k; ++rr) {
                    double rx[3] = { 0, 0, 0 }, ry[3] = { 0, 0, 0 };
                    for (int cc = 0; cc < k; ++cc) {
                        const size_t j = (size_t)(sy + rr) * W + (sx + cc);
                        const float *ex = &gx[3 * j], *ey = &gy[3 * j];
                        if (elem_finite(ex)) { ++cntx; for (int q = 0; q < 3; ++q) rx[q] += (double)ex[q]; }
                        if (elem_finite(ey)) { ++cnty; for (int q = 0; q < 3; ++q) ry[q] += (double)ey[q]; }
                    }
                    for (int q = 0; q < 3; ++q) { sgx[q] += rx[q]; sgy[q] += ry[q]; }
                }
            }
            if (cntx == 0 || cnty == 0) continue;
            double n0 = sgy[1] * sgx[2] - sgy[2] * sgx[1];
            double n1 = sgy[2] * sgx[0] - sgy[0] * sgx[2];
            double n2 = sgy[0] * sgx[1] - sgy[1] * sgx[0];
            const double L = n0 * n0 + (n1 * n1 + n2 * n2);
            if (L == 0.0) continue;
            const double sq = sqrt(L);
            float fx = (float)(n0 / sq), fy = (float)(n1 / sq), fz = (float)(n2 / sq);
            /* pcl::flipNormalTowardsViewpoint, viewpoint = sensor origin 0 (features/normal_3d.h) */
            const float vx = 0.0f - X[i], vy = 0.0f - Y[i], vz = 0.0f - Z[i];
            const float cs = (vx * fx + vy * fy + vz * fz);
            if (cs < 0) { fx *= -1; fy *= -1; fz *= -1; }
            c->nx[i] = fx; c->ny[i] = fy; c->nz[i] = fz;
        }
}

/* ---------------------------------------------------------------------------------------------------- */
/* a3': normalEstimationOMP (StairPerception.cpp:524-541) -> pcl::NormalEstimationOMP, k-nearest-neighbour search
 * (upstream features/impl/normal_3d_omp.hpp, features/normal_3d.h, common/impl/centroid.hpp, common/impl/eigen.hpp).
 * Query = search surface = every finite point of the frame.  Neighbours: the k points with the smallest
 * flann::L2_Simple squared distance (sequential float sum of squared differences), in ascending (distance, index)
 * order -- FLANN's order on distance ties is unspecified; the index tie-break is this oracle's rule.  The exact
 * search runs on a uniform grid with shell expansion (any exact method gives the same list). */
void normals_knn(orc_ctx *c, int k, int nthreads)
{
    const int N = c->N;
    const float *X = c->x.data(), *Y = c->y.data(), *Z = c->z.data();
    c->nx.assign(N, kNaN); c->ny.assign(N, kNaN); c->nz.assign(N, kNaN);
    std::vector<int> valid; valid.reserve(N);
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int i = 0; i < N; ++i)
        if (finite3(X[i], Y[i], Z[i])) {
            valid.push_back(i);
            mn[0] = std::min(mn[0], X[i]); mx[0] = std::max(mx[0], X[i]);
            mn[1] = std::min(mn[1], Y[i]); mx[1] = std::max(mx[1], Y[i]);
            mn[2] = std::min(mn[2], Z[i]); mx[2] = std::max(mx[2], Z[i]);
        }
    const int n = (int)valid.size();
    if (n == 0 || k < 1) return;
    const double ex = (double)mx[0] - mn[0], ey = (double)mx[1] - mn[1], ez = (double)mx[2] - mn[2];
    double h = sqrt(std::max(1e-12, (ex * ey + ey * ez + ex * ez) * k / (3.14159265358979 * n)));
    h = std::max(h, 1e-4);
    const int gx = std::min(1024, (int)(ex / h) + 1), gy = std::min(1024, (int)(ey / h) + 1), gz = std::min(1024, (int)(ez / h) + 1);
    const double hx = ex / gx + 1e-9, hy = ey / gy + 1e-9, hz = ez / gz + 1e-9;       /* actual cell edges */
    auto cell_of = [&](int i, int &a, int &b, int &d) {
        a = std::min(gx - 1, std::max(0, (int)(((double)X[i] - mn[0]) / hx)));
        b = std::min(gy - 1, std::max(0, (int)(((double)Y[i] - mn[1]) / hy)));
        d = std::min(gz - 1, std::max(0, (int)(((double)Z[i] - mn[2]) / hz)));
    };
    std::vector<int> cstart((size_t)gx * gy * gz + 1, 0), order(n);
    for (int i : valid) { int a, b, d; cell_of(i, a, b, d); ++cstart[((size_t)d * gy + b) * gx + a + 1]; }
    for (size_t q = 1; q < cstart.size(); ++q) cstart[q] += cstart[q - 1];
    { std::vector<int> fill(cstart.begin(), cstart.end() - 1);
      for (int i : valid) { int a, b, d; cell_of(i, a, b, d); order[fill[((size_t)d * gy + b) * gx + a]++] = i; } }
    const double hmin = std::min(hx, std::min(hy, hz));
    auto work = [&](int lo, int hi) {
        std::vector<std::pair<float, int>> cand;
        for (int vi = lo; vi < hi; ++vi) {
            const int i = valid[vi];
            int a, b, d; cell_of(i, a, b, d);
            const int want = std::min(k, n);
            cand.clear();
            for (int shell = 0;; ++shell) {
                for (int dz = -shell; dz <= shell; ++dz) for (int dy = -shell; dy <= shell; ++dy) for (int dx = -shell; dx <= shell; ++dx) {
                    if (std::max(std::abs(dx), std::max(std::abs(dy), std::abs(dz))) != shell) continue;
                    const int ca = a + dx, cb = b + dy, cd = d + dz;
                    if (ca < 0 || cb < 0 || cd < 0 || ca >= gx || cb >= gy || cd >= gz) continue;
                    const size_t cell = ((size_t)cd * gy + cb) * gx + ca;
                    for (int q = cstart[cell]; q < cstart[cell + 1]; ++q) {
                        const int j = order[q];
                        const float d0 = X[i] - X[j], d1 = Y[i] - Y[j], d2 = Z[i] - Z[j];
                        float dist = 0.f; dist += d0 * d0; dist += d1 * d1; dist += d2 * d2;
                        cand.push_back({ dist, j });
                    }
                }
                const bool all = shell >= std::max(gx, std::max(gy, gz));
                if ((int)cand.size() >= want || all) {
                    std::sort(cand.begin(), cand.end());
                    if (all) break;
                    if ((int)cand.size() >= want) {
                        const double reach = shell * hmin * (1.0 - 1e-6);                /* everything nearer than this has been seen */
                        if ((double)cand[want - 1].first < reach * reach) break;
                    }
                }
            }
            const int cnt = std::min((int)cand.size(), want);
            /* computeMeanAndCovarianceMatrix (float, single pass, neighbour order) + solvePlaneParameters */
            float accu[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
            for (int q = 0; q < cnt; ++q) {
                const int j = cand[q].second;
                const float px = X[j], py = Y[j], pz = Z[j];
                accu[0] += px * px; accu[1] += px * py; accu[2] += px * pz;
                accu[3] += py * py; accu[4] += py * pz; accu[5] += pz * pz;
                accu[6] += px; accu[7] += py; accu[8] += pz;
            }
            const float fc = (float)cnt;
            for (int q = 0; q < 9; ++q) accu[q] = accu[q] / fc;
            float cov[9];
            cov[0] = accu[0] - accu[6] * accu[6]; cov[1] = accu[1] - accu[6] * accu[7]; cov[2] = accu[2] - accu[6] * accu[8];
            cov[4] = accu[3] - accu[7] * accu[7]; cov[5] = accu[4] - accu[7] * accu[8]; cov[8] = accu[5] - accu[8] * accu[8];
            cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
            float ev, evec[3];
            smallest_eigenpair(c->eig_math, cov, &ev, evec, nullptr);      /* (the audit of eigen33 runs on the refits: single-threaded) */
            float fx = evec[0], fy = evec[1], fz = evec[2];
            const float vx = 0.0f - X[i], vy = 0.0f - Y[i], vz = 0.0f - Z[i];
            const float cs = (vx * fx + vy * fy + vz * fz);
            if (cs < 0) { fx *= -1; fy *= -1; fz *= -1; }
            c->nx[i] = fx; c->ny[i] = fy; c->nz[i] = fz;
        }
    };
    nthreads = std::max(1, std::min(nthreads, 64));
    if (nthreads == 1) work(0, n);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(work, (int)((int64_t)n * t / nthreads), (int)((int64_t)n * (t + 1) / nthreads));
        for (auto &t : th) t.join();
    }
}

/* ---------------------------------------------------------------------------------------------------- */
/* a4: three PassThrough passes == one predicate (StairPerception.cpp:394-466; upstream filters/impl/passthrough.hpp) */
void crop(orc_ctx *c, std::vector<int> &out)
{
    const float x0 = (float)c->p.x_min, x1 = (float)c->p.x_max, y0 = (float)c->p.y_min, y1 = (float)c->p.y_max,
                z0 = (float)c->p.z_min, z1 = (float)c->p.z_max;
    out.clear();
    for (int i = 0; i < c->N; ++i) {
        const float px = c->x[i], py = c->y[i], pz = c->z[i];
        if (!finite3(px, py, pz)) continue;
        if (px < x0 || px > x1) continue;
        if (py < y0 || py > y1) continue;
        if (pz < z0 || pz > z1) continue;
        out.push_back(i);
    }
}

/* a5: pcl::VoxelGridIndices::applyFilterIndices, /root/reference/modules/stairperception/voxel_grid_indices.hpp:277-517.
 * returns 0 ok, 2 = overflow guard (:295-303). */
int voxel_pick(orc_ctx *c, const std::vector<int> &in, std::vector<int> &out, std::vector<uint32_t> &key_of_pixel, float *minmax6)
{
    out.clear();
    key_of_pixel.assign(c->N, 0xFFFFFFFFu);
    const float lx = (float)c->p.voxel_x, ly = (float)c->p.voxel_y, lz = (float)c->p.voxel_z;
    const float ix = 1.0f / lx, iy = 1.0f / ly, iz = 1.0f / lz;                 /* voxel_grid_indices.h:131 */
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int i : in) {                                                          /* getMinMax3D */
        mn[0] = std::min(mn[0], c->x[i]); mx[0] = std::max(mx[0], c->x[i]);
        mn[1] = std::min(mn[1], c->y[i]); mx[1] = std::max(mx[1], c->y[i]);
        mn[2] = std::min(mn[2], c->z[i]); mx[2] = std::max(mx[2], c->z[i]);
    }
    for (int k = 0; k < 3; ++k) { minmax6[k] = mn[k]; minmax6[3 + k] = mx[k]; }
    const int64_t dx = (int64_t)((mx[0] - mn[0]) * ix) + 1, dy = (int64_t)((mx[1] - mn[1]) * iy) + 1,
                  dz = (int64_t)((mx[2] - mn[2]) * iz) + 1;
    if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) return 2;
    const int minb[3] = { (int)floor(mn[0] * ix), (int)floor(mn[1] * iy), (int)floor(mn[2] * iz) };
    const int maxb[3] = { (int)floor(mx[0] * ix), (int)floor(mx[1] * iy), (int)floor(mx[2] * iz) };
    const int div[3] = { maxb[0] - minb[0] + 1, maxb[1] - minb[1] + 1, maxb[2] - minb[2] + 1 };
    const int mul[3] = { 1, div[0], div[0] * div[1] };
    struct KI { uint32_t key; int idx; };
    std::vector<KI> v; v.reserve(in.size());
    for (int i : in) {
        const int a = (int)(floor(c->x[i] * ix) - (float)minb[0]);
        const int b = (int)(floor(c->y[i] * iy) - (float)minb[1]);
        const int d = (int)(floor(c->z[i] * iz) - (float)minb[2]);
        const int idx = a * mul[0] + b * mul[1] + d * mul[2];
        v.push_back({ (uint32_t)idx, i });
        key_of_pixel[i] = (uint32_t)idx;
    }
    if (c->voxel_order == 0) std::stable_sort(v.begin(), v.end(), [](const KI &a, const KI &b) { return a.key < b.key; });
    else std::sort(v.begin(), v.end(), [](const KI &a, const KI &b) { return a.key < b.key; });
    size_t first = 0;
    while (first < v.size()) {
        size_t last = first + 1;
        while (last < v.size() && v[last].key == v[first].key) ++last;
        /* pcl::CentroidPoint: sequential float sum, divided by the count (common/impl/centroid.hpp) */
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (size_t li = first; li < last; ++li) { sx += c->x[v[li].idx]; sy += c->y[v[li].idx]; sz += c->z[v[li].idx]; }
        const float n = (float)(last - first);
        const float cx = sx / n, cy = sy / n, cz = sz / n;
        float min_dis = FLT_MAX; int p_idx = -1;
        for (size_t li = first; li < last; ++li) {
            const int j = v[li].idx;
            const float dis = (cx - c->x[j]) * (cx - c->x[j]) + (cy - c->y[j]) * (cy - c->y[j]) + (cz - c->z[j]) * (cz - c->z[j]);
            if (min_dis > dis) { min_dis = dis; p_idx = j; }
        }
        out.push_back(p_idx);
        first = last;
    }
    return 0;
}

/* ---------------------------------------------------------------------------------------------------- */
/* a9: pcl::EuclideanClusterExtraction (StairPerception.cpp:935-963; upstream segmentation/impl/extract_clusters.hpp,
 * kdtree_flann radiusSearch with L2_Simple and RadiusResultSet's strict '<').  `pix` = pixel indices of the cloud.
 * Output: clusters (each a list of positions in `pix`, ascending), ordered by size descending, ties by smaller
 * first member (documented rule; std::sort on reverse iterators is unspecified on ties).  root_of[pos] = smallest
 * member position of the point's cluster, or -1 when the cluster is outside [min,max].                       */
void euclidean_clusters(orc_ctx *c, const std::vector<int> &pix, std::vector<std::vector<int>> &clusters, std::vector<int> &root_of)
{
    const int n = (int)pix.size();
    clusters.clear(); root_of.assign(n, -1);
    if (n == 0) return;
    const double radius = (double)(float)c->p.cluster_tolerance;
    const float r2 = (float)(radius * radius);
    const float cell = (float)radius;
    struct Key { int a, b, d; bool operator==(const Key &o) const { return a == o.a && b == o.b && d == o.d; } };
    struct KH { size_t operator()(const Key &k) const { return (size_t)k.a * 73856093u ^ (size_t)k.b * 19349663u ^ (size_t)k.d * 83492791u; } };
    std::unordered_map<Key, std::vector<int>, KH> grid;
    std::vector<Key> keys(n);
    for (int i = 0; i < n; ++i) {
        const int p = pix[i];
        keys[i] = { (int)floorf(c->x[p] / cell), (int)floorf(c->y[p] / cell), (int)floorf(c->z[p] / cell) };
        grid[keys[i]].push_back(i);
    }
    std::vector<char> processed(n, 0);
    std::vector<int> queue;
    for (int i = 0; i < n; ++i) {
        if (processed[i]) continue;
        queue.clear(); queue.push_back(i); processed[i] = 1;
        size_t sq = 0;
        while (sq < queue.size()) {
            const int q = queue[sq++];
            const int pq = pix[q];
            const float qx = c->x[pq], qy = c->y[pq], qz = c->z[pq];
            for (int da = -1; da <= 1; ++da) for (int db = -1; db <= 1; ++db) for (int dd = -1; dd <= 1; ++dd) {
                auto it = grid.find(Key{ keys[q].a + da, keys[q].b + db, keys[q].d + dd });
                if (it == grid.end()) continue;
                for (int j : it->second) {
                    if (processed[j]) continue;
                    const int pj = pix[j];
                    /* flann::L2_Simple: result += diff*diff, sequential float */
                    float d0 = qx - c->x[pj], d1 = qy - c->y[pj], d2 = qz - c->z[pj];
                    float dist = 0.f; dist += d0 * d0; dist += d1 * d1; dist += d2 * d2;
                    if (dist < r2) { processed[j] = 1; queue.push_back(j); }
                }
            }
        }
        if ((int)queue.size() >= c->p.min_cluster_size && (int)queue.size() <= c->p.max_cluster_size) {
            std::vector<int> m(queue);
            std::sort(m.begin(), m.end());
            clusters.push_back(std::move(m));
        }
    }
    std::stable_sort(clusters.begin(), clusters.end(), [](const std::vector<int> &a, const std::vector<int> &b) {
        if (a.size() != b.size()) return a.size() > b.size();
        return a[0] < b[0];
    });
    for (auto &m : clusters) for (int j : m) root_of[j] = m[0];
}

/* ---------------------------------------------------------------------------------------------------- */
/* a10 / a11: pcl::SACSegmentation::segment with SACMODEL_PLANE / SACMODEL_PARALLEL_PLANE, SAC_RANSAC
 * (upstream segmentation/impl/sac_segmentation.hpp, sample_consensus/impl/ransac.hpp, sac_model.h,
 *  impl/sac_model_plane.hpp, impl/sac_model_parallel_plane.hpp, common/impl/centroid.hpp, common/impl/eigen.hpp). */
struct Pts { std::vector<float> x, y, z; std::vector<int> pix; int n() const { return (int)x.size(); } };

struct SegCfg {
    int max_iters; float thr; bool parallel; double sin_angle; int eig_math; indep::Audit *audit;
};

/* Eigen packet reduction of a 4-float dot: (a0+a1)+(a2+a3) */
inline float dot4(float a, float b, float c, float d, float x, float y, float z, float w) { return (a * x + b * y) + (c * z + d * w); }

bool sample_degenerate(const Pts &P, int i0, int i1, int i2)
{
    /* dy1dy2 = (p1-p0)/(p2-p0); collinear iff all three ratios equal (isSampleGood / computeModelCoefficients) */
    const float r0 = (P.x[i1] - P.x[i0]) / (P.x[i2] - P.x[i0]);
    const float r1 = (P.y[i1] - P.y[i0]) / (P.y[i2] - P.y[i0]);
    const float r2 = (P.z[i1] - P.z[i0]) / (P.z[i2] - P.z[i0]);
    return (r0 == r1) && (r2 == r1);
}

void model_from_sample(const Pts &P, int i0, int i1, int i2, float *m)
{
    const float ax = P.x[i1] - P.x[i0], ay = P.y[i1] - P.y[i0], az = P.z[i1] - P.z[i0];
    const float bx = P.x[i2] - P.x[i0], by = P.y[i2] - P.y[i0], bz = P.z[i2] - P.z[i0];
    float n0 = ay * bz - az * by, n1 = az * bx - ax * bz, n2 = ax * by - ay * bx;
    const float nrm = sqrtf((n0 * n0 + n1 * n1) + (n2 * n2 + 0.0f));
    n0 = n0 / nrm; n1 = n1 / nrm; n2 = n2 / nrm;
    m[0] = n0; m[1] = n1; m[2] = n2;
    m[3] = -1 * dot4(n0, n1, n2, 0.0f, P.x[i0], P.y[i0], P.z[i0], 1.0f);
}

bool model_valid(const SegCfg &cfg, const float *m)
{
    if (!cfg.parallel) return true;
    /* SampleConsensusModelParallelPlane::isModelValid (PCL 1.8 form): |axis . n^| > |sin eps| => invalid */
    const float nrm = sqrtf((m[0] * m[0] + m[1] * m[1]) + (m[2] * m[2] + 0.0f));
    const float c0 = m[0] / nrm;
    return !((double)fabsf(c0) > cfg.sin_angle);
}

inline bool is_inlier(const float *m, float px, float py, float pz, float thr)
{
    return fabsf(dot4(m[0], m[1], m[2], m[3], px, py, pz, 1.0f)) < thr;
}

int count_within(const SegCfg &cfg, const Pts &P, const float *m)
{
    if (!model_valid(cfg, m)) return 0;
    int n = 0;
    for (int i = 0; i < P.n(); ++i) if (is_inlier(m, P.x[i], P.y[i], P.z[i], cfg.thr)) ++n;
    return n;
}

void select_within(const SegCfg &cfg, const Pts &P, const float *m, std::vector<int> &inl)
{
    inl.clear();
    if (!model_valid(cfg, m)) return;
    for (int i = 0; i < P.n(); ++i) if (is_inlier(m, P.x[i], P.y[i], P.z[i], cfg.thr)) inl.push_back(i);
}

/* one row of the RANSAC call log */
struct CallLog { int cluster, call, n, scored, b0, b1, b2, best_count, n_inliers, emitted; };

bool segment_once(const SegCfg &cfg, const Pts &P, std::vector<int> &inliers, float *coef, CallLog &lg)
{
    const int n = P.n();
    inliers.clear();
    lg.n = n; lg.scored = 0; lg.b0 = lg.b1 = lg.b2 = -1; lg.best_count = 0; lg.n_inliers = 0; lg.emitted = 0;
    if (n < 3) return false;                                       /* getSamples: too few indices */
    std::mt19937 rng(12345u);                                      /* boost::mt19937, PCL's fixed seed */
    auto rnd = [&]() -> uint32_t { return (uint32_t)(rng() >> 1); };  /* boost::uniform_int<>(0, INT_MAX) */
    std::vector<int> sh(n);
    for (int i = 0; i < n; ++i) sh[i] = i;
    int iterations = 0, best = -std::numeric_limits<int>::max();
    double k = 1.0;
    const double log_probability = ransac_log(cfg.eig_math, 1.0 - 0.99, cfg.audit);
    const double one_over = 1.0 / (double)n;
    unsigned skipped = 0; const unsigned max_skip = (unsigned)cfg.max_iters * 10u;
    float best_m[4] = { 0, 0, 0, 0 }; bool have = false;
    while (iterations < k && skipped < max_skip) {
        int s[3]; bool good = false;
        for (int it = 0; it < 1000 && !good; ++it) {               /* max_sample_checks_ = 1000 */
            for (int i = 0; i < 3; ++i) std::swap(sh[i], sh[i + (rnd() % (uint32_t)(n - i))]);
            s[0] = sh[0]; s[1] = sh[1]; s[2] = sh[2];
            good = !sample_degenerate(P, s[0], s[1], s[2]);
        }
        if (!good) break;
        if (sample_degenerate(P, s[0], s[1], s[2])) { ++skipped; continue; }   /* computeModelCoefficients: never after a good sample */
        float m[4];
        model_from_sample(P, s[0], s[1], s[2], m);
        const int cnt = count_within(cfg, P, m);
        ++lg.scored;
        if (cnt > best) {
            best = cnt; have = true; memcpy(best_m, m, sizeof(m));
            lg.b0 = s[0]; lg.b1 = s[1]; lg.b2 = s[2]; lg.best_count = cnt;
            const double w = (double)best * one_over;
            double p_no = 1.0 - (cfg.eig_math ? pow(w, 3.0) : w * w * w);
            p_no = std::max(std::numeric_limits<double>::epsilon(), p_no);
            p_no = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no);
            k = log_probability / ransac_log(cfg.eig_math, p_no, cfg.audit);
        }
        ++iterations;
        if (iterations > cfg.max_iters) break;
    }
    if (!have) return false;
    std::vector<int> inl;
    select_within(cfg, P, best_m, inl);
    /* optimizeModelCoefficients */
    float refined[4]; memcpy(refined, best_m, sizeof(refined));
    if (inl.size() >= 4) {
        float accu[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        for (int i : inl) {
            const float px = P.x[i], py = P.y[i], pz = P.z[i];
            accu[0] += px * px; accu[1] += px * py; accu[2] += px * pz;
            accu[3] += py * py; accu[4] += py * pz; accu[5] += pz * pz;
            accu[6] += px; accu[7] += py; accu[8] += pz;
        }
        const float cnt = (float)inl.size();
        for (int q = 0; q < 9; ++q) accu[q] = accu[q] / cnt;
        float cov[9];
        cov[0] = accu[0] - accu[6] * accu[6]; cov[1] = accu[1] - accu[6] * accu[7]; cov[2] = accu[2] - accu[6] * accu[8];
        cov[4] = accu[3] - accu[7] * accu[7]; cov[5] = accu[4] - accu[7] * accu[8]; cov[8] = accu[5] - accu[8] * accu[8];
        cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
        float ev, evec[3];
        smallest_eigenpair(cfg.eig_math, cov, &ev, evec, cfg.audit);
        refined[0] = evec[0]; refined[1] = evec[1]; refined[2] = evec[2];
        refined[3] = -1 * dot4(evec[0], evec[1], evec[2], 0.0f, accu[6], accu[7], accu[8], 1.0f);
        if (!model_valid(cfg, refined)) memcpy(refined, best_m, sizeof(refined));
    }
    memcpy(coef, refined, sizeof(refined));
    select_within(cfg, P, refined, inliers);
    lg.n_inliers = (int)inliers.size();
    return true;
}

/* caller loops, StairPerception.cpp:773-808 (H) and :853-884 (V) */
void segment_planes(orc_ctx *c, const std::vector<int> &pix, bool vertical, std::vector<SegPlane> &planes,
                    std::vector<int> &root_of, int *n_clusters, std::vector<CallLog> &logs)
{
    std::vector<std::vector<int>> clusters;
    euclidean_clusters(c, pix, clusters, root_of);
    *n_clusters = (int)clusters.size();
    SegCfg cfg;
    cfg.max_iters = c->p.seg_max_iters; cfg.thr = (float)c->p.seg_threshold; cfg.parallel = vertical;
    const float eps = deg2radf((float)c->p.seg_plane_angle_diff);
    cfg.sin_angle = fabs(sin((double)eps)); cfg.eig_math = c->eig_math; cfg.audit = c->audit_on ? &c->audit : nullptr;
    const int rest = c->p.seg_rest_point;
    for (size_t ci = 0; ci < clusters.size(); ++ci) {
        Pts P;
        for (int j : clusters[ci]) { const int p = pix[j]; P.x.push_back(c->x[p]); P.y.push_back(c->y[p]); P.z.push_back(c->z[p]); P.pix.push_back(p); }
        int call = 0;
        for (;;) {
            if (!vertical && P.n() < rest) break;
            std::vector<int> inl; float coef[4] = { 0, 0, 0, 0 };
            CallLog lg; lg.cluster = (int)ci; lg.call = call++;
            segment_once(cfg, P, inl, coef, lg);
            const bool emit = (int)inl.size() >= rest;
            lg.emitted = emit ? 1 : 0;
            logs.push_back(lg);
            if (!emit) break;
            SegPlane sp; memcpy(sp.coef, coef, sizeof(coef));
            Pts Q; size_t t = 0;
            for (int i = 0; i < P.n(); ++i) {
                if (t < inl.size() && inl[t] == i) { sp.idx.push_back(P.pix[i]); ++t; }
                else { Q.x.push_back(P.x[i]); Q.y.push_back(P.y[i]); Q.z.push_back(P.z[i]); Q.pix.push_back(P.pix[i]); }
            }
            planes.push_back(std::move(sp));
            P = std::move(Q);
            if (P.n() < rest) break;
        }
    }
}

/* ---------------------------------------------------------------------------------------------------- */
/* glibc rand(): TYPE_3 additive feedback r[i] = r[i-3] + r[i-31], output >> 1 (glibc stdlib/random_r.c) */
void glibc_rand_stream(uint32_t seed, int n, int32_t *out)
{
    if (seed == 0) seed = 1;
    std::vector<int32_t> r(344 + n);
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; ++i) {
        int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        int64_t w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (int32_t)w;
    }
    for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
    for (int i = 34; i < 344; ++i) r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
    for (int i = 344; i < 344 + n; ++i) {
        r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
        out[i - 344] = (int32_t)(((uint32_t)r[i]) >> 1);
    }
}

int next_rand(orc_ctx *c)
{
    if (c->rand_stream.empty()) { c->rand_stream.resize(1 << 16); glibc_rand_stream(1, 1 << 16, c->rand_stream.data()); }
    return c->rand_stream[(c->rand_pos++) & 0xFFFF];
}

/* a12: mergeSeperatedPlanes, StairPerception.cpp:965-1090 (in place, rand() % number-of-planes quirk kept) */
void merge_planes(orc_ctx *c, std::vector<SegPlane> &planes)
{
    const int S = 20;
    const double ath = (double)deg2radf((float)c->p.merge_angle_diff);
    const float thr = (float)c->p.merge_threshold;
    struct P3 { float x, y, z; };
    std::vector<std::vector<P3>> samples;
    auto sample = [&](size_t i) {
        std::vector<P3> v;
        for (int j = 0; j < S; ++j) {
            const int pi = planes[i].idx[(size_t)next_rand(c) % planes.size()];
            v.push_back({ c->x[pi], c->y[pi], c->z[pi] });
        }
        return v;
    };
    for (size_t i = 0; i < planes.size(); ++i) samples.push_back(sample(i));
    for (size_t i = 0; i < planes.size(); ++i) {
        const float a1 = planes[i].coef[0], b1 = planes[i].coef[1], c1 = planes[i].coef[2], d1 = planes[i].coef[3];
        const std::vector<P3> s1 = samples[i];
        for (size_t j = i + 1; j < planes.size(); ++j) {
            const float a2 = planes[j].coef[0], b2 = planes[j].coef[1], c2 = planes[j].coef[2], d2 = planes[j].coef[3];
            const float dot = a1 * a2 + (b1 * b2 + c1 * c2);                   /* Eigen Vector3f dot: v0 + (v1 + v2) */
            const float angle = merge_acos(c->eig_math, dot, c->audit_on ? &c->audit : nullptr);
            if (((double)angle < ath) || (M_PI - (double)angle < ath)) {
                size_t count = 0;
                for (int k = 0; k < S; ++k) {
                    const float d = fabsf(a2 * s1[k].x + b2 * s1[k].y + c2 * s1[k].z + d2);
                    if (d <= thr) count++;
                }
                const std::vector<P3> &s2 = samples[j];
                for (int k = 0; k < S; ++k) {
                    const float d = fabsf(a1 * s2[k].x + b1 * s2[k].y + c1 * s2[k].z + d1);
                    if (d <= thr) count++;
                }
                if (count > (size_t)S) {
                    const int n1 = (int)planes[i].idx.size(), n2 = (int)planes[j].idx.size(), n = n1 + n2;
                    float a, b, cc, d;
                    if (dot > 0) { a = (a1 * n1 + a2 * n2) / n; b = (b1 * n1 + b2 * n2) / n; cc = (c1 * n1 + c2 * n2) / n; d = (d1 * n1 + d2 * n2) / n; }
                    else { a = (a1 * n1 - a2 * n2) / n; b = (b1 * n1 - b2 * n2) / n; cc = (c1 * n1 - c2 * n2) / n; d = (d1 * n1 - d2 * n2) / n; }
                    planes[i].idx.insert(planes[i].idx.end(), planes[j].idx.begin(), planes[j].idx.end());
                    planes[i].coef[0] = a; planes[i].coef[1] = b; planes[i].coef[2] = cc; planes[i].coef[3] = d;
                    planes.erase(planes.begin() + j);
                    samples[i] = sample(i);
                    samples.erase(samples.begin() + j);
                    break;
                }
            }
        }
    }
}

/* a13: computeVectorPlaneInfo, StairPerception.cpp:1196-1325.  Eigen::JacobiSVD of the symmetric PSD covariance is
 * replaced by a fixed-sweep Jacobi eigen-solve (include/sp_detmath.h); consumers use sign-invariant quantities. */
void plane_info(orc_ctx *c, const SegPlane &sp, int type, orc_plane_record &r)
{
    memset(&r, 0, sizeof(r));
    const int n = (int)sp.idx.size();
    r.type = type; r.ptype = 3; r.count = n;
    memcpy(r.coef, sp.coef, sizeof(r.coef));
    float xc = 0, yc = 0, zc = 0;
    for (int i : sp.idx) { xc += c->x[i]; yc += c->y[i]; zc += c->z[i]; }
    const float fn = (float)(size_t)n;
    const float cx = xc / fn, cy = yc / fn, cz = zc / fn;
    r.center[0] = cx; r.center[1] = cy; r.center[2] = cz;
    float s[6] = { 0, 0, 0, 0, 0, 0 };   /* xx xy xz yy yz zz, sequential float */
    for (int i : sp.idx) {
        const float px = c->x[i] - cx, py = c->y[i] - cy, pz = c->z[i] - cz;
        s[0] += px * px; s[1] += px * py; s[2] += px * pz; s[3] += py * py; s[4] += py * pz; s[5] += pz * pz;
    }
    float cov[9] = { s[0] / fn, s[1] / fn, s[2] / fn, s[1] / fn, s[3] / fn, s[4] / fn, s[2] / fn, s[4] / fn, s[5] / fn };
    float evals[3], evecs[9];
    full_eigen(c->eig_math, cov, evals, evecs, c->audit_on ? &c->audit : nullptr);
    for (int j = 0; j < 3; ++j) { r.eval[j] = evals[j]; for (int k = 0; k < 3; ++k) r.evec[3 * j + k] = evecs[k * 3 + j]; }
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    int imn[3] = { -1, -1, -1 }, imx[3] = { -1, -1, -1 };
    for (int i : sp.idx) {
        const float px = c->x[i] - cx, py = c->y[i] - cy, pz = c->z[i] - cz;
        for (int a = 0; a < 3; ++a) {
            const float pj = px * r.evec[3 * a + 0] + (py * r.evec[3 * a + 1] + pz * r.evec[3 * a + 2]);   /* Vector3f dot */
            if (pj > mx[a]) { mx[a] = pj; imx[a] = i; }
            else if (pj < mn[a]) { mn[a] = pj; imn[a] = i; }      /* the reference's else-if quirk, :1279-1310 */
        }
    }
    for (int a = 0; a < 3; ++a) {
        r.pmin[a] = mn[a]; r.pmax[a] = mx[a]; r.idx_min[a] = imn[a]; r.idx_max[a] = imx[a];
        for (int k = 0; k < 3; ++k) { r.pt_min[a][k] = kNaN; r.pt_max[a][k] = kNaN; }
        if (imn[a] >= 0) { r.pt_min[a][0] = c->x[imn[a]]; r.pt_min[a][1] = c->y[imn[a]]; r.pt_min[a][2] = c->z[imn[a]]; }
        if (imx[a] >= 0) { r.pt_max[a][0] = c->x[imx[a]]; r.pt_max[a][1] = c->y[imx[a]]; r.pt_max[a][2] = c->z[imx[a]]; }
    }
    for (int k = 0; k < 3; ++k) r.pcenter[k] = (r.pt_max[0][k] + r.pt_min[0][k]) / 2;
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void dump_planes(orc_ctx *c, const char *prefix, const std::vector<SegPlane> &planes)
{
    std::vector<float> coef; std::vector<int> cnt, idx;
    for (auto &p : planes) { coef.insert(coef.end(), p.coef, p.coef + 4); cnt.push_back((int)p.idx.size()); idx.insert(idx.end(), p.idx.begin(), p.idx.end()); }
    put_tap(c, (std::string(prefix) + "_coef").c_str(), coef);
    put_tap(c, (std::string(prefix) + "_count").c_str(), cnt);
    put_tap(c, (std::string(prefix) + "_idx").c_str(), idx);
}

} // namespace

/* ==================================================================================================== */
extern "C" {

orc_ctx *orc_create(const orc_params *p) { orc_ctx *c = new orc_ctx(); c->p = *p; memset(c->stage_ms, 0, sizeof(c->stage_ms)); return c; }
void orc_destroy(orc_ctx *c) { delete c; }

int orc_set_mode(orc_ctx *c, const char *name, int value)
{
    const std::string n(name);
    if (n == "normal_arith") c->normal_arith = value;
    else if (n == "voxel_order") c->voxel_order = value;
    else if (n == "eig_math") c->eig_math = value;
    else if (n == "acos_float") c->acos_float = value;
    else if (n == "knn_threads") c->knn_threads = value;
    else return -1;
    return 0;
}

int orc_process(orc_ctx *c, const void *xyzrgba, int width, int height, const float *R9, int show_mode, int normal_mode, int keep_taps)
{
    if (!c || !xyzrgba || width <= 0 || height <= 0) return -1;
    c->taps.clear(); c->records.clear(); c->plane_idx.clear(); c->rand_pos = 0;
    memset(c->stage_ms, 0, sizeof(c->stage_ms));
    c->audit.reset(); c->audit_on = keep_taps != 0;
    c->W = width; c->H = height; c->N = width * height;
    int32_t counts[16]; memset(counts, 0, sizeof(counts));   /* same order as sp_frame_result's leading ints */
    auto finish = [&](int status) { counts[0] = status; put_tap(c, "counts", counts, 16); return (int)c->records.size(); };
    if (show_mode == 0) return finish(0);

    double t0 = now_ms(), t1;
    level_and_mask(c, (const uint8_t *)xyzrgba, R9);
    t1 = now_ms(); c->stage_ms[ST_LEVEL] = t1 - t0; t0 = t1;
    if (normal_mode == 0) normals_integral(c);
    else normals_knn(c, c->p.normal_compute_points, c->knn_threads);
    t1 = now_ms(); c->stage_ms[ST_NORMALS] = t1 - t0; t0 = t1;
    if (keep_taps) {
        std::vector<float> xyz((size_t)c->N * 3), nrm((size_t)c->N * 3);
        for (int i = 0; i < c->N; ++i) {
            xyz[3 * (size_t)i] = c->x[i]; xyz[3 * (size_t)i + 1] = c->y[i]; xyz[3 * (size_t)i + 2] = c->z[i];
            nrm[3 * (size_t)i] = c->nx[i]; nrm[3 * (size_t)i + 1] = c->ny[i]; nrm[3 * (size_t)i + 2] = c->nz[i];
        }
        put_tap(c, "xyz", xyz); put_tap(c, "normals", nrm);
    }
    int nvalid = 0;
    for (int i = 0; i < c->N; ++i) if (finite3(c->x[i], c->y[i], c->z[i])) ++nvalid;
    counts[1] = nvalid;

    std::vector<int> crop_idx;
    crop(c, crop_idx);
    counts[2] = (int)crop_idx.size();
    t1 = now_ms(); c->stage_ms[ST_CROP] = t1 - t0; t0 = t1;
    if (keep_taps) put_tap(c, "crop_idx", crop_idx);
    if (show_mode == 1) return finish(0);
    if (crop_idx.size() < 100) return finish(1);                       /* StairPerception.cpp:56 */

    std::vector<int> voxel_idx; std::vector<uint32_t> key_of_pixel; float minmax[6];
    const int vst = voxel_pick(c, crop_idx, voxel_idx, key_of_pixel, minmax);
    counts[3] = (int)voxel_idx.size();
    t1 = now_ms(); c->stage_ms[ST_VOXEL] = t1 - t0; t0 = t1;
    if (keep_taps) { put_tap(c, "voxel_idx", voxel_idx); put_tap(c, "voxel_key", key_of_pixel); put_tap(c, "minmax", minmax, 6); }

    /* a6 removeNaNNormalsFromPointCloud; a7 removeClosetPoints (:619-639); a8 extractPerpendicularPlanePoints (:670-713) */
    std::vector<int> nonan_idx, noleg_idx, h_idx, v_idx;
    for (int i : voxel_idx) if (finite3(c->nx[i], c->ny[i], c->nz[i])) nonan_idx.push_back(i);
    counts[4] = (int)nonan_idx.size();
    if (keep_taps) put_tap(c, "nonan_idx", nonan_idx);
    if (show_mode == 2) return finish(vst);
    const float th = (float)c->p.noleg_distance;
    for (int i : nonan_idx) {
        const float px = c->x[i], py = c->y[i], pz = c->z[i];
        if (px * px + py * py + pz * pz > th * th) noleg_idx.push_back(i);
    }
    counts[5] = (int)noleg_idx.size();
    if (keep_taps) put_tap(c, "noleg_idx", noleg_idx);
    if (show_mode == 3) return finish(vst);
    const double t_par = (double)deg2radf((float)c->p.parallel_angle_diff);
    const double t_perp = (double)deg2radf((float)c->p.perpendicular_angle_diff);
    for (int i : noleg_idx) {
        const float d = c->nx[i] * 1.0f + (c->ny[i] * 0.0f + c->nz[i] * 0.0f);
        const double angle = c->acos_float ? (double)acosf(d) : acos((double)d);
        if ((angle < t_par) || (M_PI - angle < t_par)) h_idx.push_back(i);
        else if (fabs(angle - M_PI / 2) < t_perp) v_idx.push_back(i);
    }
    counts[6] = (int)h_idx.size(); counts[7] = (int)v_idx.size();
    t1 = now_ms(); c->stage_ms[ST_FILTER] = t1 - t0; t0 = t1;
    if (keep_taps) { put_tap(c, "h_idx", h_idx); put_tap(c, "v_idx", v_idx); }
    if (show_mode == 4 || show_mode == 5) return finish(vst);

    std::vector<SegPlane> planes_h, planes_v; std::vector<int> root_h, root_v; std::vector<CallLog> log_h, log_v;
    int ncl_h = 0, ncl_v = 0;
    segment_planes(c, h_idx, false, planes_h, root_h, &ncl_h, log_h);
    counts[8] = ncl_h; counts[10] = (int)planes_h.size();
    t1 = now_ms(); c->stage_ms[ST_SEG_H] = t1 - t0; t0 = t1;
    if (keep_taps) { put_tap(c, "h_root", root_h); dump_planes(c, "h_seg", planes_h); put_tap(c, "h_log", (const int *)log_h.data(), log_h.size() * 10); }
    if (show_mode == 6) return finish(vst);
    segment_planes(c, v_idx, true, planes_v, root_v, &ncl_v, log_v);
    counts[9] = ncl_v; counts[11] = (int)planes_v.size();
    t1 = now_ms(); c->stage_ms[ST_SEG_V] = t1 - t0; t0 = t1;
    if (keep_taps) { put_tap(c, "v_root", root_v); dump_planes(c, "v_seg", planes_v); put_tap(c, "v_log", (const int *)log_v.data(), log_v.size() * 10); }
    if (show_mode == 7) return finish(vst);

    merge_planes(c, planes_h);
    counts[12] = (int)planes_h.size();
    if (keep_taps) dump_planes(c, "h_merge", planes_h);
    if (show_mode == 8) { t1 = now_ms(); c->stage_ms[ST_MERGE] = t1 - t0; return finish(vst); }
    merge_planes(c, planes_v);
    counts[13] = (int)planes_v.size();
    t1 = now_ms(); c->stage_ms[ST_MERGE] = t1 - t0; t0 = t1;
    if (keep_taps) dump_planes(c, "v_merge", planes_v);
    if (show_mode == 9) return finish(vst);

    std::vector<orc_plane_record> rec_h(planes_h.size()), rec_v(planes_v.size());
    for (size_t i = 0; i < planes_h.size(); ++i) plane_info(c, planes_h[i], 1, rec_h[i]);
    for (size_t i = 0; i < planes_v.size(); ++i) plane_info(c, planes_v[i], 0, rec_v[i]);
    /* a14: sortPlanesByXValues, StairPerception.cpp:1383-1441 */
    std::vector<float> chx, cvx;
    for (auto &r : rec_h) chx.push_back(r.center[0]);
    for (auto &r : rec_v) cvx.push_back(r.center[0]);
    const size_t total = rec_h.size() + rec_v.size();
    for (size_t i = 0; i < total; ++i) {
        float xmax = -FLT_MAX; size_t hi = 0, vi = 0; bool in_h = false;
        for (size_t h = 0; h < chx.size(); ++h) if (chx[h] > xmax) { xmax = chx[h]; hi = h; in_h = true; }
        for (size_t v = 0; v < cvx.size(); ++v) if (cvx[v] > xmax) { xmax = cvx[v]; vi = v; in_h = false; }
        const SegPlane *sp; orc_plane_record r;
        if (in_h) { r = rec_h[hi]; sp = &planes_h[hi]; chx[hi] = -FLT_MAX; }
        else {
            if (cvx.empty()) break;                     /* reference would read an uninitialised index here */
            r = rec_v[vi]; sp = &planes_v[vi]; cvx[vi] = -FLT_MAX;
        }
        r.first = (int)c->plane_idx.size();
        c->plane_idx.insert(c->plane_idx.end(), sp->idx.begin(), sp->idx.end());
        c->records.push_back(r);
    }
    counts[14] = (int)c->records.size(); counts[15] = (int)c->plane_idx.size();
    t1 = now_ms(); c->stage_ms[ST_INFO] = t1 - t0;
    if (keep_taps) { put_tap(c, "planes", c->records); put_tap(c, "plane_idx", c->plane_idx); }
    return finish(vst);
}

int64_t orc_process_batch(const orc_params *p, const void *xyzrgba, int width, int height, const float *R9, int nframes,
                          int nthreads, int normal_arith, int32_t *plane_counts)
{
    if (nthreads < 1) nthreads = 1;
    std::vector<int64_t> totals(nthreads, 0);
    std::vector<std::thread> th;
    const size_t fb = (size_t)width * height * 16;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t]() {
            orc_ctx *c = orc_create(p);
            c->normal_arith = normal_arith;
            for (int f = t; f < nframes; f += nthreads) {
                const int np = orc_process(c, (const uint8_t *)xyzrgba + fb * f, width, height, R9 ? R9 + 9 * (size_t)f : nullptr, 10, 0, 0);
                if (plane_counts) plane_counts[f] = np;
                totals[t] += np > 0 ? np : 0;
            }
            orc_destroy(c);
        });
    for (auto &t : th) t.join();
    int64_t s = 0; for (auto v : totals) s += v;
    return s;
}

int64_t orc_tap_bytes(orc_ctx *c, const char *name) { auto it = c->taps.find(name); return it == c->taps.end() ? -1 : (int64_t)it->second.bytes.size(); }
int64_t orc_tap_copy(orc_ctx *c, const char *name, void *dst, int64_t cap)
{
    auto it = c->taps.find(name);
    if (it == c->taps.end()) return -1;
    const int64_t n = std::min<int64_t>(cap, (int64_t)it->second.bytes.size());
    if (n > 0) memcpy(dst, it->second.bytes.data(), (size_t)n);
    return n;
}
int64_t orc_tap_names(orc_ctx *c, char *dst, int64_t cap)
{
    std::string s;
    for (auto &kv : c->taps) { s += kv.first; s += '\n'; }
    if (dst && cap > 0) { const int64_t n = std::min<int64_t>(cap - 1, (int64_t)s.size()); memcpy(dst, s.data(), (size_t)n); dst[n] = 0; }
    return (int64_t)s.size() + 1;
}

/* the audit of the last orc_process(keep_taps = 1) call, see oracle/indep_math.h: 11 doubles */
int orc_audit(orc_ctx *c, double *out, int cap)
{
    const indep::Audit &a = c->audit;
    const double v[11] = { (double)a.n_eigen33, (double)a.n_jacobi, (double)a.n_acos, (double)a.n_log, a.eigen33_max_angle, a.eigen33_max_angle_wellsep,
                           a.eigen33_max_val, a.jacobi_max_val, a.jacobi_max_angle, a.acos_max_abs, a.log_max_rel };
    const int n = std::min(cap, 11);
    for (int i = 0; i < n; ++i) out[i] = v[i];
    return n;
}
int orc_stage_ms(orc_ctx *c, double *out, int cap) { int n = std::min(cap, (int)ST_COUNT); for (int i = 0; i < n; ++i) out[i] = c->stage_ms[i]; return n; }
const char *orc_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : ""; }

/* computeRotationMatrix, /root/reference/test/reader.cpp:228-271, the way Eigen 3 evaluates it in a plain x86-64 (SSE2) build
 * [Eigen-recall: Geometry/AngleAxis.h, Geometry/Quaternion.h, Geometry/arch/Geometry_SSE.h, Core/arch/SSE/PacketMath.h]:
 *   - AngleAxisf * AngleAxisf * AngleAxisf is a chain of QUATERNION products, converted to a matrix once;
 *   - the float quaternion product is the SSE kernel: four lanes (x, y, z, w) of
 *        (a * b.wwww - a.zxyx * b.yzxx) + (+,+,+,-)(a.yzxz * b.zxyz + a.wwwy * b.xyzy)
 *   - normalized() divides by sqrt of the SSE2 horizontal sum (c0 + c2) + (c1 + c3) of the squared coefficients.
 * Written as packet arithmetic over 4-float arrays, unlike the product's scalar form (sp_b200.cu), on purpose. */
typedef float packet4[4];                                /* coefficient order of Eigen::Quaternionf: x, y, z, w */
static void swz(const packet4 a, int p, int q, int r, int s, packet4 o) { const float t[4] = { a[p], a[q], a[r], a[s] }; memcpy(o, t, sizeof(t)); }
static void quat_mul_sse(const packet4 a, const packet4 b, packet4 res)
{
    packet4 a1, b1, a2, b2, a3, b3, bw, s1, s2, lhs;
    swz(a, 1, 2, 0, 2, a1); swz(b, 2, 0, 1, 2, b1);
    swz(a, 3, 3, 3, 1, a2); swz(b, 0, 1, 2, 1, b2);
    swz(a, 2, 0, 1, 0, a3); swz(b, 1, 2, 0, 0, b3);
    swz(b, 3, 3, 3, 3, bw);
    for (int i = 0; i < 4; ++i) { s1[i] = a1[i] * b1[i]; s2[i] = a2[i] * b2[i]; }
    for (int i = 0; i < 4; ++i) { const float m = a[i] * bw[i], n = a3[i] * b3[i]; lhs[i] = m - n; }
    for (int i = 0; i < 4; ++i) { float t = s1[i] + s2[i]; if (i == 3) t = -t; res[i] = lhs[i] + t; }
}
static void quat_to_matrix(const packet4 q, float *m)
{
    /* Eigen::QuaternionBase::toRotationMatrix */
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    m[0] = 1.f - (tyy + tzz); m[1] = txy - twz; m[2] = txz + twy;
    m[3] = txy + twz; m[4] = 1.f - (txx + tzz); m[5] = tyz - twx;
    m[6] = txz - twy; m[7] = tyz + twx; m[8] = 1.f - (txx + tyy);
}
static void finish_rotation(const float *m2, float *R9)
{
    /* m3 * m2 * m1 with m1 = [[0,0,1],[-1,0,0],[0,-1,0]], m3 = [[0,0,-1],[1,0,0],[0,-1,0]]: a signed permutation of m2's
     * entries (every product term is 0 or +-m2(i,j): exact), written out: R(i,j) = sum_k sum_l m3(i,k) m2(k,l) m1(l,j) */
    R9[0] = m2[7];  R9[1] = m2[8];  R9[2] = -m2[6];
    R9[3] = -m2[1]; R9[4] = -m2[2]; R9[5] = m2[0];
    R9[6] = m2[4];  R9[7] = m2[5];  R9[8] = -m2[3];
}
void orc_rotation_from_euler(float pitch, float roll, float yaw, float *R9)
{
    const float ang[3] = { (float)(roll / 180.0 * M_PI), (float)(pitch / 180.0 * M_PI), (float)(yaw / 180.0 * M_PI) };   /* about X, Y, Z */
    packet4 q[3];
    for (int k = 0; k < 3; ++k) {
        const float ha = 0.5f * ang[k], sn = sinf(ha);
        for (int i = 0; i < 3; ++i) q[k][i] = sn * (i == k ? 1.0f : 0.0f);
        q[k][3] = cosf(ha);
    }
    packet4 t, r;
    quat_mul_sse(q[0], q[1], t); quat_mul_sse(t, q[2], r);
    float m2[9]; quat_to_matrix(r, m2);
    finish_rotation(m2, R9);
}
void orc_rotation_from_quat(float w, float x, float y, float z, float *R9)
{
    packet4 c = { x, y, z, w }, sq;
    for (int i = 0; i < 4; ++i) sq[i] = c[i] * c[i];
    const float n = sqrtf((sq[0] + sq[2]) + (sq[1] + sq[3]));
    for (int i = 0; i < 4; ++i) c[i] = c[i] / n;
    float m2[9]; quat_to_matrix(c, m2);
    finish_rotation(m2, R9);
}

void orc_mt19937_draws(uint32_t seed, int n, int32_t *out) { std::mt19937 r(seed); for (int i = 0; i < n; ++i) out[i] = (int32_t)(r() >> 1); }
void orc_glibc_rand(uint32_t seed, int n, int32_t *out) { glibc_rand_stream(seed, n, out); }
void orc_eigen33(const float *cov9, int libm, float *ev, float *evec) { if (libm) det_eigen33f<1>(cov9, ev, evec); else det_eigen33f<0>(cov9, ev, evec); }
void orc_jacobi3(const float *cov9, float *evals3, float *evecs9) { det_jacobi3f(cov9, evals3, evecs9); }

/* K2G::updateCloud + K2G::prepareMake3D, /root/reference/modules/k2g/k2g.cpp:231-308, :506-520: undistorted depth (mm) and
 * registered colour -> the XYZRGBA cloud (16-byte records).  depth_type 0 = float32, 1 = uint16. */
void orc_depth_to_cloud(const void *depth, int depth_type, const uint8_t *bgrx, int W, int H, float fx, float fy, float cx, float cy,
                        int mirror, void *out_records)
{
    std::vector<float> colmap(W), rowmap(H);
    for (int i = 0; i < W; i++) colmap[i] = (float)((i - cx + 0.5) / fx);
    for (int i = 0; i < H; i++) rowmap[i] = (float)((i - cy + 0.5) / fy);
    uint8_t *out = (uint8_t *)out_records;
    for (int y = 0; y < H; ++y) {
        const float dy = rowmap[y];
        for (int x = 0; x < W; ++x) {
            const int sx = mirror ? (W - 1 - x) : x;                       /* cv::flip(.., 1) before the loop */
            const size_t si = (size_t)y * W + sx;
            const float raw = depth_type == 0 ? ((const float *)depth)[si] : (float)((const uint16_t *)depth)[si];
            const float depth_value = raw / 1000.0f;
            float px, py, pz; uint32_t rgba;
            if (!std::isnan(depth_value) && !(std::abs(depth_value) < 0.0001)) {
                px = colmap[x] * depth_value; py = dy * depth_value; pz = depth_value;
                rgba = 0xFFFFFFFFu;
                if (bgrx) rgba = (uint32_t)bgrx[4 * si] | ((uint32_t)bgrx[4 * si + 1] << 8) | ((uint32_t)bgrx[4 * si + 2] << 16) | 0xFF000000u;
            } else { px = py = pz = kNaN; rgba = 0; }
            uint8_t *o = out + 16 * ((size_t)y * W + x);
            memcpy(o, &px, 4); memcpy(o + 4, &py, 4); memcpy(o + 8, &pz, 4); memcpy(o + 12, &rgba, 4);
        }
    }
}

/* findMinMaxProjPoint, StairPerception.cpp:2780-2808, on plane `plane` of the last frame processed */
int orc_minmax_proj(orc_ctx *c, int plane, const float *center3, const float *dir3, float *max_xyz, float *min_xyz, int32_t *max_pixel, int32_t *min_pixel)
{
    if (!c || plane < 0 || plane >= (int)c->records.size() || !dir3) return -1;
    const orc_plane_record &r = c->records[plane];
    const float *cen = center3 ? center3 : r.center;
    float max_proj = -FLT_MAX, min_proj = FLT_MAX; int imax = -1, imin = -1;
    for (int i = 0; i < r.count; ++i) {
        const int pix = c->plane_idx[r.first + i];
        const float px = c->x[pix] - cen[0], py = c->y[pix] - cen[1], pz = c->z[pix] - cen[2];
        const float proj = dir3[0] * px + (dir3[1] * py + dir3[2] * pz);      /* Eigen's unrolled 3-term reduction */
        if (proj > max_proj) { max_proj = proj; imax = pix; }
        if (proj < min_proj) { min_proj = proj; imin = pix; }
    }
    for (int k = 0; k < 3; ++k) { if (max_xyz) max_xyz[k] = kNaN; if (min_xyz) min_xyz[k] = kNaN; }
    if (imax >= 0 && max_xyz) { max_xyz[0] = c->x[imax]; max_xyz[1] = c->y[imax]; max_xyz[2] = c->z[imax]; }
    if (imin >= 0 && min_xyz) { min_xyz[0] = c->x[imin]; min_xyz[1] = c->y[imin]; min_xyz[2] = c->z[imin]; }
    if (max_pixel) *max_pixel = imax;
    if (min_pixel) *min_pixel = imin;
    return 0;
}

} /* extern "C" */
