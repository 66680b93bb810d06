/*
 * sp_strip.cuh -- K1K2 (levelling + masking + organised normals + crop statistics) as a persistent, row-marching kernel fed by TMA.
 *
 * Same outputs, bit for bit, as the tile kernel k_ingest_normals (sp_front.cuh), which restates
 *   rotationCloudPoints            /root/reference/test/reader.cpp:273-324
 *   remove_nan_points              /root/reference/modules/stairperception/StairPerception.cpp:338-356
 *   normalEstimationIntegral       StairPerception.cpp:579-592  (pcl::IntegralImageNormalEstimation, AVERAGE_3D_GRADIENT)
 *   the crop predicate / getMinMax3D of passThroughCloudANDNormal and VoxelGridIndices (StairPerception.cpp:394-466,
 *   voxel_grid_indices.hpp:292)
 * but organised around the data flow instead of around a tile:
 *
 *   work item   = (frame, column strip, row band).  A CTA of ST_T threads owns a strip of up to ST_T - 10 image columns (+5 halo
 *                 columns each side), one thread per column, and marches down the rows of its band.  Everything a column needs
 *                 from the rows above / below is the thread's own history; only the left / right neighbours go through shared
 *                 memory.  Nothing is re-read from HBM except the 5 + 5 halo columns and the 5 + 5 halo rows of a band.
 *   input       = one TMA tensor copy per image row (cp.async.bulk.tensor.4d, two boxes of 144 pixels x 16 B) into a 3-slot ring,
 *                 completion on an mbarrier; rows / columns outside the image arrive as NaN (the tensor map's out-of-bounds fill),
 *                 which is exactly "invalid pixel" for every later stage, so the kernel has no bounds logic for loads.
 *   software pipeline, per step (ingest row i):
 *       D  finish row r = i - 7: smoothing window per pixel from the bit rows of rows r - 4 .. r + 4, 5 x 5 box sums =
 *          own column sum + 4 neighbours' column sums (exact double sums), cross product, normalise, flip, store
 *       E  the few pixels with a window of 4, 3, 2 (next to a depth discontinuity) or with an inexact neighbourhood: ordered
 *          sums straight from the gradient ring (one list per row, full warps)
 *       A  ingest row i: level + mask, store the levelled point, crop statistics, bit rows (valid, risk)
 *       B  depth-change pair flags of rows i - 1 (horizontal pairs) and i - 2 (vertical pairs) as warp ballots
 *       C  gradients of row j = i - 4 (zeroed where not finite), finite-element bit rows, running 5-row column sums
 *          (add row j, drop row j - 5: the sums stay exact, see k1_inexact_risk), column sums of centre row i - 6 to shared memory
 *   two CTA barriers per step; every other hand-off is a warp ballot or the thread's own registers.
 */
#pragma once
#include "sp_common.cuh"
#include "sp_front.cuh"
#include <cuda.h>

#define ST_T 288                          /* threads per CTA = columns of a strip including the halo */
#define ST_NW (ST_T / 32)
#define ST_HALO 5
#define ST_SWMAX (ST_T - 2 * ST_HALO)     /* owned columns per strip, at most */
#define ST_BOX (ST_T / 2)                 /* pixels per TMA box (a box dimension is limited to 256) */
#define ST_RAW 3                          /* TMA ring slots (rows in flight) */
#define ST_PR 6                           /* levelled rows kept: i - 5 .. i */
#define ST_GR 6                           /* gradient rows kept: j - 5 .. j */
#define ST_WR 16                          /* bit-row ring */
#define ST_LAG 7                          /* row r is finished when row r + 7 is ingested */

struct __align__(128) StSmem {
    float4 raw[ST_RAW][ST_T];             /* TMA destination: raw records {x, y, z, rgba} */
    float4 P[ST_PR][ST_T];                /* levelled + masked points {x, y, z, 0} */
    float4 GA[ST_GR][ST_T];               /* gradients {dx.x, dx.y, dx.z, dy.x} */
    float2 GB[ST_GR][ST_T];               /*           {dy.y, dy.z} */
    double2 CS[3][ST_T];                  /* 5-row column sums of the six gradient channels, centre row i - 6 */
    /* bit rows, one word per warp (+ a zero guard word at either end): [row ring][warp + 1] */
    unsigned hp[ST_WR][ST_NW + 2], vp[ST_WR][ST_NW + 2];      /* depth-change pair flags: (c, c + 1) and (r, r + 1) */
    unsigned fa[ST_WR][ST_NW + 2], fb[ST_WR][ST_NW + 2];      /* finite dx / dy gradient elements */
    unsigned rk[ST_WR][ST_NW + 2], vm[ST_WR][ST_NW + 2];      /* inexact-risk coordinates; pixels that may get a normal */
    unsigned short slow[2][ST_T + 32];    /* per finished row: pixels that take the ordered sums (column | window << 12) */
    int nslow[2];
    int anyrisk;
    unsigned long long mbar[ST_RAW];
};

__device__ __forceinline__ unsigned st_sptr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void st_mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_sptr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void st_mbar_expect(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_sptr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "ST_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra ST_DONE;\n"
        "bra ST_WAIT;\n"
        "ST_DONE:\n"
        "}\n" ::"r"(st_sptr(bar)), "r"(parity) : "memory");
}
/* one box of ST_BOX pixels of image row `row` of frame `frame`, starting at column `col` (any of them may lie outside: NaN fill) */
__device__ __forceinline__ void st_tma_row(const CUtensorMap *tmap, void *dst, unsigned long long *bar, int col, int row, int frame)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(st_sptr(dst)), "l"((unsigned long long)tmap), "r"(0), "r"(col), "r"(row), "r"(frame), "r"(st_sptr(bar)) : "memory");
}

/* 64-bit window of a bit row around warp `wid`: bit 16 + k = column k of the warp, 16 columns of margin either side */
__device__ __forceinline__ unsigned long long st_local(const unsigned (*rows)[ST_NW + 2], int slot, int wid)
{
    const unsigned *w = rows[slot & (ST_WR - 1)] + wid;
    return ((unsigned long long)w[0] >> 16) | ((unsigned long long)w[1] << 16) | ((unsigned long long)w[2] << 48);
}

/* k1_finish with the normalisation off the slow path.  The definition stays fx = (float)(n0 / sqrt(L)) with IEEE double sqrt and
 * division (two Newton-heavy library routines per pixel on this part); here: q = n0 * rs with rs = 1 / sqrt(L) from the hardware
 * approximation (rsqrt.approx.f64, >= 17 good bits) refined by two coupled Newton steps in FMA arithmetic.  q and the defined
 * quotient differ by at most 3 ulp of double (measured over 3e8 cases, oracle-side test tests/test_oracle_cpu.py::test_fast_normalise),
 * so they round to the same float unless q lies within a few ulp of a float rounding boundary: the low 29 mantissa bits of q
 * within 256 of 0x10000000.  Those cases (one component in a million), tiny quotients and extreme L take the defined route. */
__device__ __forceinline__ bool st_round_safe(double q)
{
    const int lo = __double2loint(q), hi = __double2hiint(q);
    const int dist = (lo & 0x1fffffff) - 0x10000000;
    const unsigned ex = (unsigned)(hi >> 20) & 0x7ffu;
    return (abs(dist) > 256 && ex > 1023u - 100u) || q == 0.0;
}
__device__ __forceinline__ void st_finish(const SpDev &d, size_t gi, float px, float py, float pz,
                                          double sgx0, double sgx1, double sgx2, double sgy0, double sgy1, double sgy2)
{
    float onx = __int_as_float(0x7fc00000), ony = onx, onz = onx;
    const double n0 = sgy1 * sgx2 - sgy2 * sgx1;
    const double n1 = sgy2 * sgx0 - sgy0 * sgx2;
    const double n2 = sgy0 * sgx1 - sgy1 * sgx0;
    const double L = n0 * n0 + (n1 * n1 + n2 * n2);
    if (L != 0.0) {
        float fx = 0.f, fy = 0.f, fz = 0.f;
        bool ok = L > 1e-200 && L < 1e200;
        if (ok) {
            double y;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(L));
            double h = 0.5 * y, g = L * y;
            double r = __fma_rn(-g, h, 0.5);
            g = __fma_rn(g, r, g); h = __fma_rn(h, r, h);
            r = __fma_rn(-g, h, 0.5);
            h = __fma_rn(h, r, h);
            const double rs = h + h;
            const double q0 = n0 * rs, q1 = n1 * rs, q2 = n2 * rs;
            ok = st_round_safe(q0) && st_round_safe(q1) && st_round_safe(q2);
            fx = (float)q0; fy = (float)q1; fz = (float)q2;
        }
        if (!ok) {
            const double sq = sqrt(L);
            fx = (float)(n0 / sq); fy = (float)(n1 / sq); fz = (float)(n2 / sq);
        }
        const float vx = 0.0f - px, vy = 0.0f - py, vz = 0.0f - pz;
        const float cs = (vx * fx + vy * fy + vz * fz);
        if (cs < 0) { fx *= -1; fy *= -1; fz *= -1; }
        onx = fx; ony = fy; onz = fz;
    }
    d.nrm[gi] = make_float4(onx, ony, onz, 0.f);
}

/* ordered K x K sums of the pixel in ring column t (k1_normal of the tile kernel, read from the gradient ring) */
template <int K>
__device__ __forceinline__ void st_slow_normal(const StSmem &s, const SpDev &d, int t, int nrow /* row r as a band-local index */, size_t gi)
{
    const int half = K >> 1;
    double sgx0 = 0.0, sgx1 = 0.0, sgx2 = 0.0, sgy0 = 0.0, sgy1 = 0.0, sgy2 = 0.0;
#pragma unroll
    for (int a = 0; a < K; ++a) {
        double rx0 = 0.0, rx1 = 0.0, rx2 = 0.0, ry0 = 0.0, ry1 = 0.0, ry2 = 0.0;
        const int slot = (nrow - half + a) % ST_GR;
#pragma unroll
        for (int b = 0; b < K; ++b) {
            const float4 ga = s.GA[slot][t - half + b];
            const float2 gb = s.GB[slot][t - half + b];
            rx0 += (double)ga.x; rx1 += (double)ga.y; rx2 += (double)ga.z;
            ry0 += (double)ga.w; ry1 += (double)gb.x; ry2 += (double)gb.y;
        }
        sgx0 += rx0; sgx1 += rx1; sgx2 += rx2; sgy0 += ry0; sgy1 += ry1; sgy2 += ry2;
    }
    const float4 pw = d.xyzw[gi];
    st_finish(d, gi, pw.x, pw.y, pw.z, sgx0, sgx1, sgx2, sgy0, sgy1, sgy2);
}

__global__ void __launch_bounds__(ST_T, 2) k_ingest_normals_strip(SpDev d, const __grid_constant__ CUtensorMap tmap, int ns, int SW, int nb, int BR)
{
    extern __shared__ __align__(128) unsigned char st_smem_raw[];
    StSmem &s = *reinterpret_cast<StSmem *>(st_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt = (1u << lane) - 1u, bit = 1u << lane;
    const int W = d.W, H = d.H, N = d.N;
    const float qnan = __int_as_float(0x7fc00000);
    const bool rot = d.R != nullptr;

    if (tid == 0) {
        for (int k = 0; k < ST_RAW; ++k) st_mbar_init(&s.mbar[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int loads = 0;                                                        /* row loads consumed by this CTA so far: slot and parity */
    const int items = d.B * ns * nb;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int band = item % nb, strip = (item / nb) % ns, f = item / (nb * ns);
        const int r0 = band * BR, r1 = min(H, r0 + BR);
        const int s0 = strip * SW, swid = min(SW, W - s0);
        const int c = s0 - ST_HALO + tid;                                 /* this thread's image column (may lie outside) */
        const bool col_owned = tid >= ST_HALO && tid < ST_HALO + swid;
        const int i0 = r0 - ST_HALO;                                      /* first row ingested */
        const int n_ingest = (r1 - r0) + 2 * ST_HALO;
        const int nsteps = (r1 - r0) + ST_HALO + ST_LAG;
        const size_t fbase = (size_t)f * N;
        float R[9];
        if (rot) {
#pragma unroll
            for (int k = 0; k < 9; ++k) R[k] = __ldg(d.R + 9 * f + k);
        }
        /* clear the rings of the previous item */
        for (int k = tid; k < ST_WR * (ST_NW + 2); k += ST_T) {
            (&s.hp[0][0])[k] = 0u; (&s.vp[0][0])[k] = 0u; (&s.fa[0][0])[k] = 0u; (&s.fb[0][0])[k] = 0u; (&s.rk[0][0])[k] = 0u; (&s.vm[0][0])[k] = 0u;
        }
#pragma unroll
        for (int k = 0; k < ST_GR; ++k) { s.GA[k][tid] = make_float4(0.f, 0.f, 0.f, 0.f); s.GB[k][tid] = make_float2(0.f, 0.f); }
        if (tid == 0) { s.nslow[0] = 0; s.nslow[1] = 0; s.anyrisk = 0; }
        if (tid == 0) {                                                   /* rows i0 and i0 + 1 on their way */
            for (int n = 0; n < 2 && n < n_ingest; ++n) {
                const int slot = (loads + n) % ST_RAW;
                st_mbar_expect(&s.mbar[slot], ST_T * 16);
                st_tma_row(&tmap, &s.raw[slot][0], &s.mbar[slot], s0 - ST_HALO, i0 + n, f);
                st_tma_row(&tmap, &s.raw[slot][ST_BOX], &s.mbar[slot], s0 - ST_HALO + ST_BOX, i0 + n, f);
            }
        }
        __syncthreads();

        float z1 = qnan, z2 = qnan;                                       /* levelled z of rows i - 1, i - 2 of this column */
        double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0, w4 = 0.0, w5 = 0.0;/* running 5-row column sums of the six gradient channels */
        int dirty = 0;                                                    /* risk mode: steps until the column sums are clean again */
        bool riskmode = false;
        float mnx = FLT_MAX, mny = FLT_MAX, mnz = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX, mxz = -FLT_MAX;
        int nvalid = 0, ncrop = 0;

        for (int n = 0; n < nsteps; ++n) {
            const int i = i0 + n;                                         /* row ingested in this step */
            if (tid == 0 && n + 2 < n_ingest) {                           /* slot of row i - 1 is free again: row i + 2 */
                const int slot = (loads + n + 2) % ST_RAW;
                st_mbar_expect(&s.mbar[slot], ST_T * 16);
                st_tma_row(&tmap, &s.raw[slot][0], &s.mbar[slot], s0 - ST_HALO, i + 2, f);
                st_tma_row(&tmap, &s.raw[slot][ST_BOX], &s.mbar[slot], s0 - ST_HALO + ST_BOX, i + 2, f);
            }
            riskmode = riskmode || (s.anyrisk != 0);                      /* CTA-uniform: written before the last barrier */

            /* ---------------------------------------------------------------- D: finish row r */
            const int r = i - ST_LAG, nr = n - ST_LAG;                    /* nr = band-local index of row r (slot arithmetic) */
            const bool fin = r >= r0 && r < r1;                           /* CTA-uniform */
            if (fin) {
                /* smoothing window per pixel (PCL's 1.0 / 1.4 chamfer map, capped at 5) from bit rows: lanes 0..8 take the flag
                 * rows r - 4 .. r + 4, lanes 9..13 the gradient rows r - 2 .. r + 2, lanes 14..20 the risk rows r - 3 .. r + 3; the
                 * per-row contributions are OR-reduced across the warp.  A flagged pixel at row distance a and column distance h
                 * allows a window of kcat[a][h] (see k_ingest_normals, S3). */
                unsigned cz0 = 0, cz2 = 0, cz3 = 0, cz4 = 0, ca5 = 0, ca4 = 0, ca3 = 0, ca2 = 0, cb5 = 0, cb4 = 0, cb3 = 0, cb2 = 0, cwr = 0;
                if (lane < 9) {
                    const int q = nr - 4 + lane, dd = lane < 4 ? 4 - lane : lane - 4;
                    const unsigned long long hq = st_local(s.hp, q, wid), vq = st_local(s.vp, q, wid), vm1 = st_local(s.vp, q - 1, wid);
                    unsigned long long D0 = hq | (hq << 1) | vq | vm1;
                    unsigned long long D1 = D0 | (D0 << 1) | (D0 >> 1);
                    unsigned long long D2 = D1 | (D1 << 1) | (D1 >> 1);
                    unsigned long long D3 = D2 | (D2 << 1) | (D2 >> 1);
                    unsigned long long D4 = D3 | (D3 << 1) | (D3 >> 1);
                    const unsigned e0 = (unsigned)(D0 >> 16), e1 = (unsigned)(D1 >> 16), e2 = (unsigned)(D2 >> 16), e3 = (unsigned)(D3 >> 16), e4 = (unsigned)(D4 >> 16);
                    cz0 = dd == 0 ? e2 : dd == 1 ? e1 : dd == 2 ? e0 : 0u;
                    cz2 = dd <= 2 ? e2 : 0u;
                    cz3 = dd <= 2 ? e3 : dd == 3 ? e2 : 0u;
                    cz4 = dd <= 2 ? e4 : dd == 3 ? e3 : e2;
                } else if (lane < 14) {
                    const int a = lane - 9, q = nr - 2 + a;               /* gradient row offset a - 2 */
                    const unsigned long long A = st_local(s.fa, q, wid), Bq = st_local(s.fb, q, wid);
                    const unsigned long long A3 = (A << 1) | A | (A >> 1), B3 = (Bq << 1) | Bq | (Bq >> 1);
                    const unsigned a5 = (unsigned)((A3 | (A << 2) | (A >> 2)) >> 16), b5 = (unsigned)((B3 | (Bq << 2) | (Bq >> 2)) >> 16);
                    const unsigned a4 = (unsigned)((A3 | (A << 2)) >> 16), b4 = (unsigned)((B3 | (Bq << 2)) >> 16);
                    const unsigned a3 = (unsigned)(A3 >> 16), b3 = (unsigned)(B3 >> 16);
                    const unsigned a2 = (unsigned)(((A << 1) | A) >> 16), b2 = (unsigned)(((Bq << 1) | Bq) >> 16);
                    ca5 = a5; cb5 = b5;
                    if (a <= 3) { ca4 = a4; cb4 = b4; }
                    if (a >= 1 && a <= 3) { ca3 = a3; cb3 = b3; }
                    if (a >= 1 && a <= 2) { ca2 = a2; cb2 = b2; }
                } else if (lane < 21 && riskmode) {
                    const unsigned long long K0 = st_local(s.rk, nr - 3 + (lane - 14), wid);
                    const unsigned long long K1 = K0 | (K0 << 1) | (K0 >> 1), K2 = K1 | (K1 << 1) | (K1 >> 1), K3 = K2 | (K2 << 1) | (K2 >> 1);
                    cwr = (unsigned)(K3 >> 16);
                }
                const unsigned z0 = __reduce_or_sync(0xffffffffu, cz0), zz2 = __reduce_or_sync(0xffffffffu, cz2), z3 = __reduce_or_sync(0xffffffffu, cz3),
                               z4 = __reduce_or_sync(0xffffffffu, cz4);
                const unsigned wn5 = __reduce_or_sync(0xffffffffu, ca5) & __reduce_or_sync(0xffffffffu, cb5);
                const unsigned wn4 = __reduce_or_sync(0xffffffffu, ca4) & __reduce_or_sync(0xffffffffu, cb4);
                const unsigned wn3 = __reduce_or_sync(0xffffffffu, ca3) & __reduce_or_sync(0xffffffffu, cb3);
                const unsigned wn2 = __reduce_or_sync(0xffffffffu, ca2) & __reduce_or_sync(0xffffffffu, cb2);
                const unsigned wrisk = riskmode ? __reduce_or_sync(0xffffffffu, cwr) : 0u;
                const unsigned v = s.vm[nr & (ST_WR - 1)][wid + 1];
                const unsigned k5 = v & wn5 & ~z4, k4 = v & wn4 & (z4 & ~z3), k3 = v & wn3 & (z3 & ~zz2), k2 = v & wn2 & (zz2 & ~z0);
                const int kk = (k5 & bit) ? 5 : (k4 & bit) ? 4 : (k3 & bit) ? 3 : (k2 & bit) ? 2 : 0;
                const bool fast = kk == 5 && !(wrisk & bit);
                const bool slow = col_owned && kk != 0 && !fast;
                const size_t gi = fbase + (size_t)r * W + c;
                if (col_owned) {
                    if (fast) {
                        const double2 a0 = s.CS[0][tid - 2], a1 = s.CS[0][tid - 1], a2 = s.CS[0][tid + 1], a3 = s.CS[0][tid + 2];
                        const double2 b0 = s.CS[1][tid - 2], b1 = s.CS[1][tid - 1], b2 = s.CS[1][tid + 1], b3 = s.CS[1][tid + 2];
                        const double2 c0 = s.CS[2][tid - 2], c1 = s.CS[2][tid - 1], c2 = s.CS[2][tid + 1], c3 = s.CS[2][tid + 2];
                        const double g0 = ((a0.x + a1.x) + w0) + (a2.x + a3.x), g1 = ((a0.y + a1.y) + w1) + (a2.y + a3.y);
                        const double g2 = ((b0.x + b1.x) + w2) + (b2.x + b3.x), g3 = ((b0.y + b1.y) + w3) + (b2.y + b3.y);
                        const double g4 = ((c0.x + c1.x) + w4) + (c2.x + c3.x), g5 = ((c0.y + c1.y) + w5) + (c2.y + c3.y);
                        const float4 pw = d.xyzw[gi];                     /* this thread stored it seven steps ago */
                        st_finish(d, gi, pw.x, pw.y, pw.z, g0, g1, g2, g3, g4, g5);
                    } else if (!slow) d.nrm[gi] = make_float4(qnan, qnan, qnan, 0.f);
                }
                const unsigned bs = __ballot_sync(0xffffffffu, slow);
                if (bs) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s.nslow[n & 1], __popc(bs));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (slow) s.slow[n & 1][base + __popc(bs & lt)] = (unsigned short)(tid | (kk << 12));
                }
            }
            __syncthreads();                                              /* column sums read; the row's slow list is complete */

            /* ---------------------------------------------------------------- E: ordered sums of the listed pixels of row r */
            if (fin) {
                const int nsl = s.nslow[n & 1];
                if (tid == 0) s.nslow[(n + 1) & 1] = 0;
                for (int u = tid; u < nsl; u += ST_T) {
                    const int code = s.slow[n & 1][u], t = code & 0xfff, kk = code >> 12;
                    const size_t gi = fbase + (size_t)r * W + (s0 - ST_HALO + t);
                    if (kk == 5) st_slow_normal<5>(s, d, t, nr, gi);
                    else if (kk == 4) st_slow_normal<4>(s, d, t, nr, gi);
                    else if (kk == 3) st_slow_normal<3>(s, d, t, nr, gi);
                    else st_slow_normal<2>(s, d, t, nr, gi);
                }
            }

            /* ---------------------------------------------------------------- A: ingest row i */
            float z0n = qnan;
            if (n < n_ingest) {
                const int slot = (loads + n) % ST_RAW;
                st_mbar_wait(&s.mbar[slot], (unsigned)(((loads + n) / ST_RAW) & 1));
                const float4 p = s.raw[slot][tid];
                float ox = p.x, oy = p.y, oz = p.z;
                if (rot) {
                    if (!isnan(p.x)) {
                        ox = R[0] * p.x + R[1] * p.y + R[2] * p.z;
                        oy = R[3] * p.x + R[4] * p.y + R[5] * p.z;
                        oz = R[6] * p.x + R[7] * p.y + R[8] * p.z;
                    } else { ox = oy = oz = qnan; }
                }
                /* remove_nan_points: r, g or b == 0 marks an invalid pixel; pixels outside the image arrive as NaN */
                if (__vcmpeq4(__float_as_uint(p.w) | 0xff000000u, 0u) || i < 0 || i >= H || c < 0 || c >= W) { ox = oy = oz = qnan; }
                const bool own = col_owned && i >= r0 && i < r1;
                bool crop = false;
                if (own) {
                    d.xyzw[fbase + (size_t)i * W + c] = make_float4(ox, oy, oz, 0.f);
                    if (sp_finite3(ox, oy, oz)) {
                        ++nvalid;
                        if (!(ox < d.c.x0 || ox > d.c.x1) && !(oy < d.c.y0 || oy > d.c.y1) && !(oz < d.c.z0 || oz > d.c.z1)) {
                            crop = true; ++ncrop;
                            mnx = fminf(mnx, ox); mny = fminf(mny, oy); mnz = fminf(mnz, oz);
                            mxx = fmaxf(mxx, ox); mxy = fmaxf(mxy, oy); mxz = fmaxf(mxz, oz);
                        }
                    }
                }
                s.P[n % ST_PR][tid] = make_float4(ox, oy, oz, 0.f);
                z0n = oz;
                const bool risk = k1_inexact_risk(ox) | k1_inexact_risk(oy) | k1_inexact_risk(oz);
                const unsigned brk = __ballot_sync(0xffffffffu, risk);
                const unsigned bvm = __ballot_sync(0xffffffffu, i >= 5 && i < H - 5 && c >= 5 && c < W - 5 && sp_finite(oz));
                const unsigned bc = __ballot_sync(0xffffffffu, crop);
                if (lane == 0) {
                    s.rk[n & (ST_WR - 1)][wid + 1] = brk; s.vm[n & (ST_WR - 1)][wid + 1] = bvm;
                    if (brk) s.anyrisk = 1;
                    if (bc) {                                             /* cropped pixels per 1024-pixel block, for the voxel stage */
                        const int p0 = i * W + (c - lane + 0), left = 1024 - (p0 & 1023);
                        const unsigned lo = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
                        int *vc = d.vcnt + (size_t)f * d.nblk + (p0 >> 10);
                        if (bc & lo) atomicAdd(vc, __popc(bc & lo));
                        if (bc & ~lo) atomicAdd(vc + 1, __popc(bc & ~lo));
                    }
                }
            }

            /* ---------------------------------------------------------------- B: depth-change pair flags (on the levelled z: the reference's quirk) */
            if (n >= 1) {
                const float zr = tid + 1 < ST_T ? s.P[(n - 1) % ST_PR][tid + 1].z : qnan;
                const unsigned bh = __ballot_sync(0xffffffffu, k1_bad_pair(z1, zr));
                if (lane == 0) s.hp[(n - 1) & (ST_WR - 1)][wid + 1] = bh;
            }
            if (n >= 2) {
                const unsigned bv = __ballot_sync(0xffffffffu, k1_bad_pair(z2, z1));
                if (lane == 0) s.vp[(n - 2) & (ST_WR - 1)][wid + 1] = bv;
            }
            z2 = z1; z1 = z0n;

            /* ---------------------------------------------------------------- C: gradients of row j = i - 4, column sums of centre row i - 6 */
            const int j = i - 4, nj = n - 4;
            if (j >= r0 - 2 && j <= r1 + 1) {                             /* CTA-uniform */
                const int tl = tid > 0 ? tid - 1 : 0, tr = tid + 1 < ST_T ? tid + 1 : ST_T - 1;
                const float4 L = s.P[nj % ST_PR][tl], Rr = s.P[nj % ST_PR][tr], U = s.P[(nj - 1) % ST_PR][tid], Dn = s.P[(nj + 1) % ST_PR][tid];
                const float ax = Rr.x - L.x, ay = Rr.y - L.y, az = Rr.z - L.z;
                const float bx = Dn.x - U.x, by = Dn.y - U.y, bz = Dn.z - U.z;
                const bool fa = sp_finite(ax + (ay + az)), fb = sp_finite(bx + (by + bz));
                const float g0 = fa ? ax : 0.0f, g1 = fa ? ay : 0.0f, g2 = fa ? az : 0.0f, g3 = fb ? bx : 0.0f, g4 = fb ? by : 0.0f, g5 = fb ? bz : 0.0f;
                const float4 oa = s.GA[(nj + 1) % ST_GR][tid];            /* row j - 5 (zero before the band's first gradient row) */
                const float2 ob = s.GB[(nj + 1) % ST_GR][tid];
                s.GA[nj % ST_GR][tid] = make_float4(g0, g1, g2, g3);
                s.GB[nj % ST_GR][tid] = make_float2(g4, g5);
                const unsigned ba = __ballot_sync(0xffffffffu, fa), bb = __ballot_sync(0xffffffffu, fb);
                if (lane == 0) { s.fa[nj & (ST_WR - 1)][wid + 1] = ba; s.fb[nj & (ST_WR - 1)][wid + 1] = bb; }
                bool rebuild = false;
                if (riskmode) {
                    /* an inexact element poisons this column's running sums until it has left the window: five steps later the
                     * sums are rebuilt from the ring (all clean again); pixels that would use a poisoned sum are on the slow list
                     * (risk within three pixels), so nothing reads it in between */
                    const unsigned long long K0 = st_local(s.rk, nj - 1, wid) | st_local(s.rk, nj, wid) | st_local(s.rk, nj + 1, wid);
                    const unsigned near = (unsigned)((K0 | (K0 << 1) | (K0 >> 1)) >> 16);
                    if (near & bit) dirty = 6;
                    if (dirty > 0) { --dirty; rebuild = dirty == 0; }
                }
                if (!rebuild) {
                    w0 += (double)g0; w1 += (double)g1; w2 += (double)g2; w3 += (double)g3; w4 += (double)g4; w5 += (double)g5;
                    w0 -= (double)oa.x; w1 -= (double)oa.y; w2 -= (double)oa.z; w3 -= (double)oa.w; w4 -= (double)ob.x; w5 -= (double)ob.y;
                } else {
                    w0 = (double)g0; w1 = (double)g1; w2 = (double)g2; w3 = (double)g3; w4 = (double)g4; w5 = (double)g5;
                    for (int a = 1; a < 5; ++a) {                         /* rows j - 1 .. j - 4 (zero rows before the band count as zero) */
                        const float4 qa = s.GA[(nj - a + ST_GR) % ST_GR][tid];
                        const float2 qb = s.GB[(nj - a + ST_GR) % ST_GR][tid];
                        w0 += (double)qa.x; w1 += (double)qa.y; w2 += (double)qa.z; w3 += (double)qa.w; w4 += (double)qb.x; w5 += (double)qb.y;
                    }
                }
                s.CS[0][tid] = make_double2(w0, w1); s.CS[1][tid] = make_double2(w2, w3); s.CS[2][tid] = make_double2(w4, w5);
            }
            __syncthreads();
        }
        loads += n_ingest;

        /* crop statistics of the band: passThroughCloudANDNormal's counts, getMinMax3D's box */
        nvalid = __reduce_add_sync(0xffffffffu, nvalid); ncrop = __reduce_add_sync(0xffffffffu, ncrop);
        if (ncrop) {                                                      /* warp-uniform */
            const int a0 = __reduce_min_sync(0xffffffffu, sp_f2ord(mnx)), a1 = __reduce_min_sync(0xffffffffu, sp_f2ord(mny)), a2 = __reduce_min_sync(0xffffffffu, sp_f2ord(mnz));
            const int b0 = __reduce_max_sync(0xffffffffu, sp_f2ord(mxx)), b1 = __reduce_max_sync(0xffffffffu, sp_f2ord(mxy)), b2 = __reduce_max_sync(0xffffffffu, sp_f2ord(mxz));
            if (lane == 0) {
                atomicAdd(&d.counts[f * SP_NCOUNT + 2], ncrop);
                int *mm = d.mm + f * 8;
                atomicMin(mm + 0, a0); atomicMin(mm + 1, a1); atomicMin(mm + 2, a2);
                atomicMax(mm + 3, b0); atomicMax(mm + 4, b1); atomicMax(mm + 5, b2);
            }
        }
        if (lane == 0 && nvalid) atomicAdd(&d.counts[f * SP_NCOUNT + 1], nvalid);
    }
}
