/*
 * sp_front.cuh -- streaming kernels of the front-end: levelling + masking + organised normals (K1K2),
 * voxel keys, voxel representative pick + point classification, list scatter.
 */
#pragma once
#include "sp_common.cuh"
#include "sp_async.cuh"

/* ==================================================================================================== */
/* K1K2: fused rotationCloudPoints (test/reader.cpp:273-324) + remove_nan_points (StairPerception.cpp:338-356)
 * + pcl::IntegralImageNormalEstimation/AVERAGE_3D_GRADIENT (StairPerception.cpp:579-592) + the crop predicate
 * and bounding box of passThroughCloudANDNormal / getMinMax3D (StairPerception.cpp:394-466,
 * voxel_grid_indices.hpp:292).
 *
 * One CTA per TW x TH pixel tile of one frame.  The tile plus a 5-pixel halo of raw 16-byte records is read
 * once (coalesced float4 row segments), levelled and masked on the fly and kept in shared memory; everything
 * the normal needs (depth-change flags, the <=5 chamfer window, both gradient fields, the k x k box sums in
 * double) is then computed from shared memory.  HBM traffic per pixel: 16 B in, 24 B out.
 *
 * Normal definition (bit-exact with the oracle's "box64" mode): see DESIGN.md "organised normals".
 */
#define K1_TW 32
#define K1_TH 32
#define K1_HALO 5
#define K1_RW (K1_TW + 2 * K1_HALO)      /* z region: halo 5 (42) */
#define K1_RH (K1_TH + 2 * K1_HALO)
#define K1_XW (K1_TW + 6)                /* x, y region: halo 3 (gradients at halo 2 need +-1) */
#define K1_XH (K1_TH + 6)
#define K1_FW (K1_TW + 8)                /* zero-flag region: halo 4 */
#define K1_FH (K1_TH + 8)
#define K1_GW (K1_TW + 4)                /* gradient region: halo 2 */
#define K1_GH (K1_TH + 4)
#define K1_THREADS 256

#define K1_BAND0 8                        /* tile rows of the first column-sum band of the exact path (computed while S3 runs) ... */
#define K1_BAND 12                        /* ... and of the following ones */
struct K1Smem {                           /* 54.9 KB: four CTAs per SM */
    float sz[K1_RH * K1_RW];              /* sz .. fbhi are contiguous: the exact path re-uses them as its column-sum band */
    float sx[K1_XH * K1_XW], sy[K1_XH * K1_XW];
    unsigned long long dil[5][K1_FH];     /* bit rows of the flag region (bit c = column c): depth-change pixels dilated by 0..4 columns */
    unsigned falo[K1_GH], fahi[K1_GH], fblo[K1_GH], fbhi[K1_GH];   /* finite dx / dy gradient elements per gradient row: columns 0..31, 32..35 */
    alignas(128) float gx[3][K1_GH * K1_GW]; float gy[3][K1_GH * K1_GW];   /* the TMA kernel lands its raw tile here first */
    unsigned kmask[4][K1_TH];             /* per tile row: pixels whose smoothing window is 5, 4, 3, 2 */
    unsigned vmask[K1_TH], imask[K1_TH];  /* pixels that may get a normal (image border, finite z); pixels inside the image */
    unsigned short base[3][K1_TH];        /* where a row's window-5 / window-4 / window-3-and-2 pixels start in their list */
    unsigned short list[K1_TH * K1_TW];   /* pixels that get a normal, by window size: 5 from the front, 4 from the back */
    unsigned short list2[K1_TH * K1_TW / 2];   /* windows of 3 and 2, tagged (pixel | k << 12); a tile with more takes them in place */
    int n5, n4, n32, pad;
    unsigned long long mbar;              /* completion of the raw tile (k_ingest_normals_tma) */
};
static_assert(K1_RH * K1_RW * 16 <= 6 * K1_GH * K1_GW * 4 && offsetof(K1Smem, gx) % 128 == 0, "the raw tile of the TMA kernel aliases the gradient fields");
static_assert(K1_FW <= 64 && K1_TW == 32 && K1_FH % 4 == 0, "bit rows: one 64-bit word per flag row, one warp lane per tile column");
static_assert(6 * K1_BAND0 * K1_GW * sizeof(double) <= (K1_RH * K1_RW + 2 * K1_XH * K1_XW) * sizeof(float), "the first band must stay inside sz + sx + sy (S3 still reads dil)");
static_assert(6 * K1_BAND * K1_GW * sizeof(double) <= offsetof(K1Smem, gx), "column-sum band must fit in front of the gradients");
static_assert(offsetof(K1Smem, dil) % 8 == 0 && (K1_TH - K1_BAND0) % K1_BAND == 0 && 6 * K1_GW <= K1_THREADS - 32, "layout / one thread per (channel, gradient column)");

__device__ __forceinline__ bool k1_bad_pair(float za, float zb)
{
    /* depth-change test of the pair (a = upper/left element), integral_image_normal.hpp */
    const float thr = (0.02f * (fabsf(za) + 1.0f) * 2.0f);
    return (fabsf(za - zb) > thr) || !sp_finite(za) || !sp_finite(zb);
}

/* cross product of the two summed gradients, normalisation, flip towards the origin (PCL's view point), store */
__device__ __forceinline__ void k1_finish(const SpDev &d, size_t gi, float px, float py, float pz,
                                          double sgx0, double sgx1, double sgx2, double sgy0, double sgy1, double sgy2)
{
    float onx = __int_as_float(0x7fc00000), ony = onx, onz = onx;
    const double n0 = sgy1 * sgx2 - sgy2 * sgx1;
    const double n1 = sgy2 * sgx0 - sgy0 * sgx2;
    const double n2 = sgy0 * sgx1 - sgy1 * sgx0;
    const double L = n0 * n0 + (n1 * n1 + n2 * n2);
    if (L != 0.0) {
        const double sq = sqrt(L);
        float fx = (float)(n0 / sq), fy = (float)(n1 / sq), fz = (float)(n2 / sq);
        const float vx = 0.0f - px, vy = 0.0f - py, vz = 0.0f - pz;
        const float cs = (vx * fx + vy * fy + vz * fz);
        if (cs < 0) { fx *= -1; fy *= -1; fz *= -1; }
        onx = fx; ony = fy; onz = fz;
    }
    d.nrm[gi] = make_float4(onx, ony, onz, 0.f);
}

/* k x k box sums of the two gradient fields around tile pixel e, in the oracle's order (rows top to bottom of row sums
 * left to right, double), then k1_finish.  Only pixels whose two windows hold at least one finite gradient element each get
 * here (the window masks of S3 take care of that). */
template <int K, bool FAST = false>
__device__ __forceinline__ void k1_normal(const K1Smem &s, const SpDev &d, int e, size_t tile_base, int W)
{
    const int rr = e / K1_TW, cc = e - rr * K1_TW;
    const int half = K >> 1;
    const int g0 = (rr + 2 - half) * K1_GW + (cc + 2 - half);
    const size_t gi = tile_base + (size_t)rr * W + cc;
    const float4 pw = d.xyzw[gi];                       /* the point itself, for the flip (written by this CTA a barrier ago): requested
                                                           before the sums so that its latency hides behind them */
    double sgx0 = 0.0, sgx1 = 0.0, sgx2 = 0.0, sgy0 = 0.0, sgy1 = 0.0, sgy2 = 0.0;
#pragma unroll
    for (int a = 0; a < K; ++a) {
        double rx0 = 0.0, rx1 = 0.0, rx2 = 0.0, ry0 = 0.0, ry1 = 0.0, ry2 = 0.0;
        const int gb = g0 + a * K1_GW;
#pragma unroll
        for (int b = 0; b < K; ++b) {
            const int ge = gb + b;
            rx0 += (double)s.gx[0][ge]; rx1 += (double)s.gx[1][ge]; rx2 += (double)s.gx[2][ge];
            ry0 += (double)s.gy[0][ge]; ry1 += (double)s.gy[1][ge]; ry2 += (double)s.gy[2][ge];
        }
        sgx0 += rx0; sgx1 += rx1; sgx2 += rx2; sgy0 += ry0; sgy1 += ry1; sgy2 += ry2;
    }
    if (FAST) st_finish(d, gi, pw.x, pw.y, pw.z, sgx0, sgx1, sgx2, sgy0, sgy1, sgy2);
    else k1_finish(d, gi, pw.x, pw.y, pw.z, sgx0, sgx1, sgx2, sgy0, sgy1, sgy2);
}

/* Exact path, first half: thread u < 6 * K1_GW = (channel, gradient column) slides a 5-row column sum down the ROWS tile rows
 * from b0 on and leaves them in csum[channel][row][column] */
template <int ROWS>
__device__ __forceinline__ void k1_column_sums(const K1Smem &s, double *csum, int u, int b0)
{
    const int ch = u / K1_GW, col = u - ch * K1_GW;
    if (ch >= 6) return;
    const float *g = ch < 3 ? s.gx[ch] : s.gy[ch - 3];
    float v[ROWS + 4];
#pragma unroll
    for (int a = 0; a < ROWS + 4; ++a) v[a] = g[(b0 + a) * K1_GW + col];
    double w = (double)v[0] + (double)v[1] + (double)v[2] + (double)v[3];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
        w += (double)v[i + 4];
        csum[(ch * ROWS + i) * K1_GW + col] = w;
        w -= (double)v[i];
    }
}

/* A coordinate that is NaN, zero, or 2^-21 <= |c| < 8 is a multiple of 2^-44 below 8.  If every coordinate of the tile is like
 * that, every gradient element (a float difference of two of them) is a multiple of 2^-44 of magnitude <= 16, and any sum of up to
 * 30 of them is a multiple of 2^-44 below 2^9: 53 bits, exact in double whatever the order of the additions. */
__device__ __forceinline__ bool k1_inexact_risk(float c)
{
    const unsigned u = __float_as_uint(c) & 0x7fffffffu;
    return !(u - 0x35000000u < 0x41000000u - 0x35000000u || u == 0u || u > 0x7f800000u);
}

__global__ void __launch_bounds__(K1_THREADS, 4) k_ingest_normals(SpDev d)
{
    extern __shared__ __align__(16) unsigned char k1_raw[];
    K1Smem &s = *reinterpret_cast<K1Smem *>(k1_raw);
    const int f = blockIdx.z, W = d.W, H = d.H, N = d.N;
    const int r0 = blockIdx.y * K1_TH, c0 = blockIdx.x * K1_TW;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const float4 *in = d.in + (size_t)f * N;
    const float qnan = __int_as_float(0x7fc00000);
    const size_t tile_base = (size_t)f * N + (size_t)r0 * W + c0;

    /* S0: load tile + halo, level, mask -> shared memory.  The owned 32 x 32 block goes first, one warp per tile row (lane =
     * column, coalesced 512-byte rows): besides the shared-memory copy each pixel is stored levelled (packed float4) and enters
     * the crop statistics of passThroughCloudANDNormal / getMinMax3D, the cropped pixels per 1024-pixel block for the voxel stage
     * and the row masks; the 740 halo elements follow in a flat loop. */
    bool risk = false;
    {
        float R[9];
        const bool rot = d.R != nullptr;
        if (rot) {
#pragma unroll
            for (int i = 0; i < 9; ++i) R[i] = __ldg(d.R + 9 * f + i);
        }
        auto level = [&](const float4 &p, float &ox, float &oy, float &oz) {
            ox = p.x; oy = p.y; oz = p.z;
            if (rot) {
                if (!isnan(p.x)) {
                    ox = R[0] * p.x + R[1] * p.y + R[2] * p.z;
                    oy = R[3] * p.x + R[4] * p.y + R[5] * p.z;
                    oz = R[6] * p.x + R[7] * p.y + R[8] * p.z;
                } else { ox = oy = oz = qnan; }
            }
            /* remove_nan_points: r, g or b == 0 marks an invalid pixel (per-byte compare of the three colour bytes) */
            if (__vcmpeq4(__float_as_uint(p.w) | 0xff000000u, 0u)) { ox = oy = oz = qnan; }
        };
        float mnx = FLT_MAX, mny = FLT_MAX, mnz = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX, mxz = -FLT_MAX;
        int nvalid = 0, ncrop = 0;
        const int gc = c0 + lane;
        float4 pv[K1_TH / 8];                                             /* the thread's four records are requested together */
#pragma unroll
        for (int it = 0; it < K1_TH / 8; ++it) {
            const int gr = r0 + it * 8 + wid;
            pv[it] = make_float4(qnan, qnan, qnan, __int_as_float(-1));
            if (gr < H && gc < W) pv[it] = __ldg(in + (size_t)gr * W + gc);
        }
#pragma unroll
        for (int it = 0; it < K1_TH / 8; ++it) {
            const int tr = it * 8 + wid, gr = r0 + tr;
            const bool inside = gr < H && gc < W;
            float ox = qnan, oy = qnan, oz = qnan;
            bool crop = false;
            if (inside) {
                level(pv[it], ox, oy, oz);
                d.xyzw[tile_base + (size_t)tr * W + lane] = make_float4(ox, oy, oz, 0.f);
                if (sp_finite3(ox, oy, oz)) {
                    ++nvalid;
                    if (!(ox < d.c.x0 || ox > d.c.x1) && !(oy < d.c.y0 || oy > d.c.y1) && !(oz < d.c.z0 || oz > d.c.z1)) {
                        crop = true; ++ncrop;
                        mnx = fminf(mnx, ox); mny = fminf(mny, oy); mnz = fminf(mnz, oz);
                        mxx = fmaxf(mxx, ox); mxy = fmaxf(mxy, oy); mxz = fmaxf(mxz, oz);
                    }
                }
            }
            s.sz[(tr + K1_HALO) * K1_RW + lane + K1_HALO] = oz;
            s.sx[(tr + 3) * K1_XW + lane + 3] = ox; s.sy[(tr + 3) * K1_XW + lane + 3] = oy;
            risk |= k1_inexact_risk(ox) | k1_inexact_risk(oy) | k1_inexact_risk(oz);
            const unsigned bc = __ballot_sync(0xffffffffu, crop);
            const unsigned bi = __ballot_sync(0xffffffffu, inside);
            const unsigned bv = __ballot_sync(0xffffffffu, inside && gr >= 5 && gr < H - 5 && gc >= 5 && gc < W - 5 && sp_finite(oz));
            if (lane == 0) {
                s.imask[tr] = bi; s.vmask[tr] = bv;
                if (bc) {
                    const int i0 = gr * W + c0, left = 1024 - (i0 & 1023);               /* pixels of this row inside i0's block */
                    const unsigned lo = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
                    int *vc = d.vcnt + (size_t)f * d.nblk + (i0 >> 10);
                    if (bc & lo) atomicAdd(vc, __popc(bc & lo));
                    if (bc & ~lo) atomicAdd(vc + 1, __popc(bc & ~lo));
                }
            }
        }
        nvalid = __reduce_add_sync(0xffffffffu, nvalid); ncrop = __reduce_add_sync(0xffffffffu, ncrop);
        if (ncrop) {                                                       /* warp-uniform */
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
        /* the halo ring: 5 rows of 42 above, 5 below, 32 rows of 5 + 5 at the sides */
        constexpr int RING_TB = K1_HALO * K1_RW, RING = 2 * RING_TB + K1_TH * 2 * K1_HALO;
        for (int e = tid; e < RING; e += K1_THREADS) {
            int rr, cc;
            if (e < 2 * RING_TB) { const int q = e < RING_TB ? e : e - RING_TB; rr = q / K1_RW; cc = q - rr * K1_RW; if (e >= RING_TB) rr += K1_HALO + K1_TH; }
            else { const int q = e - 2 * RING_TB; rr = q / (2 * K1_HALO); cc = q - rr * (2 * K1_HALO); rr += K1_HALO; if (cc >= K1_HALO) cc += K1_TW; }
            const int gr = r0 - K1_HALO + rr, gc2 = c0 - K1_HALO + cc;
            float ox = qnan, oy = qnan, oz = qnan;
            if (gr >= 0 && gr < H && gc2 >= 0 && gc2 < W) level(__ldg(in + (size_t)gr * W + gc2), ox, oy, oz);
            s.sz[rr * K1_RW + cc] = oz;
            risk |= k1_inexact_risk(ox) | k1_inexact_risk(oy) | k1_inexact_risk(oz);
            const unsigned xr = (unsigned)(rr - 2), xc = (unsigned)(cc - 2);       /* x, y keep a halo of 3 */
            if (xr < (unsigned)K1_XH && xc < (unsigned)K1_XW) { s.sx[xr * K1_XW + xc] = ox; s.sy[xr * K1_XW + xc] = oy; }
        }
    }
    const bool exact = __syncthreads_or(risk) == 0;      /* box sums of this tile are exact in double in any order */

    /* S1 + S2: depth-change flags on the lateral coordinate z (reference quirk), halo 4, as bit rows, four flag rows per warp
     * pass: columns 0..31 of each row (the ballot is the row), then the eight tail columns of the four rows in one go; lanes
     * 0..3 then dilate "their" row: a flagged pixel within h columns, h = 0..4 (PCL's distance map along the row, capped at 5) */
    for (int q = wid; q < K1_FH / 4; q += 8) {
        unsigned long long m = 0ull;
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const int fr = q * 4 + (u < 4 ? u : (lane >> 3));
            const int se = (fr + 1) * K1_RW + (u < 4 ? lane : 32 + (lane & 7)) + 1;
            const float zc = s.sz[se];
            const bool z0 = k1_bad_pair(zc, s.sz[se + 1]) || k1_bad_pair(s.sz[se - 1], zc) ||
                            k1_bad_pair(zc, s.sz[se + K1_RW]) || k1_bad_pair(s.sz[se - K1_RW], zc);
            const unsigned b = __ballot_sync(0xffffffffu, z0);
            if (u < 4) { if (lane == u) m = b; }
            else m |= (unsigned long long)((b >> (8 * (lane & 3))) & 255u) << 32;
        }
        if (lane < 4) {
            const int fr = q * 4 + lane;
            s.dil[0][fr] = m;
#pragma unroll
            for (int h = 1; h < 5; ++h) { m |= (m << 1) | (m >> 1); s.dil[h][fr] = m; }
        }
    }
    /* S4 (independent of S1): gradient fields, halo 2, same row-aligned passes: columns 0..31 of a gradient row per warp, then
     * the four tail columns of eight rows per warp; the ballots are the rows of finite elements */
    for (int q = wid; q < K1_GH + (K1_GH + 7) / 8; q += 8) {
        const bool tail = q >= K1_GH;
        const int rr = tail ? (q - K1_GH) * 8 + (lane >> 2) : q, cc = tail ? 32 + (lane & 3) : lane;
        bool fa = false, fb = false;
        if (rr < K1_GH) {
            const int e = rr * K1_GW + cc;
            const int se = (rr + 3) * K1_RW + cc + 3, xe = (rr + 1) * K1_XW + cc + 1;
            const float ax = s.sx[xe + 1] - s.sx[xe - 1], ay = s.sy[xe + 1] - s.sy[xe - 1], az = s.sz[se + 1] - s.sz[se - 1];
            const float bx = s.sx[xe + K1_XW] - s.sx[xe - K1_XW], by = s.sy[xe + K1_XW] - s.sy[xe - K1_XW], bz = s.sz[se + K1_RW] - s.sz[se - K1_RW];
            fa = sp_finite(ax + (ay + az)); fb = sp_finite(bx + (by + bz));
            s.gx[0][e] = fa ? ax : 0.0f; s.gx[1][e] = fa ? ay : 0.0f; s.gx[2][e] = fa ? az : 0.0f;
            s.gy[0][e] = fb ? bx : 0.0f; s.gy[1][e] = fb ? by : 0.0f; s.gy[2][e] = fb ? bz : 0.0f;
        }
        const unsigned ba = __ballot_sync(0xffffffffu, fa), bb = __ballot_sync(0xffffffffu, fb);
        if (!tail) { if (lane == 0) { s.falo[rr] = ba; s.fblo[rr] = bb; } }
        else if ((lane & 3) == 0 && rr < K1_GH) { s.fahi[rr] = (ba >> (lane & 28)) & 15u; s.fbhi[rr] = (bb >> (lane & 28)) & 15u; }
    }
    __syncthreads();
    /* S3: smoothing size per owned pixel, one thread per tile row.  A flagged pixel at row distance a and column distance h
     * allows a window of kcat[a][h] (chamfer 1.0/1.4: d = 1.4*min + (max-min); d<=2 -> 0 (NaN), (2,3) -> 2, [3,4) -> 3, [4,5) -> 4,
     * >=5 -> 5):  a=0: 0 0 0 3 4 5 | a=1: 0 0 2 3 4 5 | a=2: 0 2 2 3 4 5 | a=3: 3 3 3 4 5 5 | a=4: 4 4 4 5 5 5  (h = 0..5).
     * "window <= k" is therefore an OR of dilated rows; the list offsets of each row come from a warp scan of the counts.
     * Meanwhile warps 1..7 of an exact tile start on the separable sums of S5 (column sums of the first band): the coordinates
     * in sz / sx / sy had their last readers before the barrier above, the band may overwrite them. */
    double *csum = reinterpret_cast<double *>(k1_raw);                      /* [6][band rows][K1_GW], over sz / sx / sy (+ dil, fa / fb rows) */
    if (exact && tid >= 32) k1_column_sums<K1_BAND0>(s, csum, tid - 32, 0);
    if (tid < K1_TH) {
        const int r = tid + 4;
        const unsigned long long (*D)[K1_FH] = s.dil;
        const unsigned long long z0 = D[2][r] | D[1][r - 1] | D[1][r + 1] | D[0][r - 2] | D[0][r + 2];
        const unsigned long long z2 = D[2][r] | D[2][r - 1] | D[2][r + 1] | D[2][r - 2] | D[2][r + 2];
        const unsigned long long z3 = D[3][r] | D[3][r - 1] | D[3][r + 1] | D[3][r - 2] | D[3][r + 2] | D[2][r - 3] | D[2][r + 3];
        const unsigned long long z4 = D[4][r] | D[4][r - 1] | D[4][r + 1] | D[4][r - 2] | D[4][r + 2] | D[3][r - 3] | D[3][r + 3] | D[2][r - 4] | D[2][r + 4];
        /* a normal needs one finite element in each gradient window (PCL's finite-element counts): rows tid + a of the gradient
         * region, columns cc + b; windows of 5: a, b = 0..4; 4: 0..3; 3: 1..3; 2: 1..2 */
        unsigned long long fa[5], fb[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            fa[a] = (unsigned long long)s.falo[tid + a] | ((unsigned long long)s.fahi[tid + a] << 32);
            fb[a] = (unsigned long long)s.fblo[tid + a] | ((unsigned long long)s.fbhi[tid + a] << 32);
        }
        const unsigned long long a2 = fa[1] | fa[2], a3 = a2 | fa[3], a4 = a3 | fa[0], a5 = a4 | fa[4];
        const unsigned long long b2 = fb[1] | fb[2], b3 = b2 | fb[3], b4 = b3 | fb[0], b5 = b4 | fb[4];
        const unsigned w2 = (unsigned)((a2 >> 1) | (a2 >> 2)) & (unsigned)((b2 >> 1) | (b2 >> 2));
        const unsigned w3 = (unsigned)((a3 >> 1) | (a3 >> 2) | (a3 >> 3)) & (unsigned)((b3 >> 1) | (b3 >> 2) | (b3 >> 3));
        const unsigned w4 = (unsigned)(a4 | (a4 >> 1) | (a4 >> 2) | (a4 >> 3)) & (unsigned)(b4 | (b4 >> 1) | (b4 >> 2) | (b4 >> 3));
        const unsigned w5 = (unsigned)(a5 | (a5 >> 1) | (a5 >> 2) | (a5 >> 3) | (a5 >> 4)) & (unsigned)(b5 | (b5 >> 1) | (b5 >> 2) | (b5 >> 3) | (b5 >> 4));
        const unsigned v = s.vmask[tid];
        const unsigned k5 = v & w5 & ~(unsigned)(z4 >> 4), k4 = v & w4 & (unsigned)((z4 & ~z3) >> 4), k3 = v & w3 & (unsigned)((z3 & ~z2) >> 4),
                       k2 = v & w2 & (unsigned)((z2 & ~z0) >> 4);
        s.kmask[0][tid] = k5; s.kmask[1][tid] = k4; s.kmask[2][tid] = k3; s.kmask[3][tid] = k2;
        int c5 = __popc(k5), c4 = __popc(k4), c32 = __popc(k3 | k2);
        int i5 = c5, i4 = c4, i32 = c32;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t5 = __shfl_up_sync(0xffffffffu, i5, o), t4 = __shfl_up_sync(0xffffffffu, i4, o), t32 = __shfl_up_sync(0xffffffffu, i32, o);
            if (lane >= o) { i5 += t5; i4 += t4; i32 += t32; }
        }
        s.base[0][tid] = (unsigned short)(i5 - c5); s.base[1][tid] = (unsigned short)(i4 - c4); s.base[2][tid] = (unsigned short)(i32 - c32);
        if (tid == K1_TH - 1) { s.n5 = i5; s.n4 = i4; s.n32 = i32; }
    }
    __syncthreads();
    /* Pixels without a normal (invalid, border, next to a depth discontinuity: about a third of the tile) get their NaN here;
     * the others go into one list per window size, so that the box sums below run with full warps and fixed trip counts. */
#pragma unroll
    for (int it = 0; it < K1_TH / 8; ++it) {
        const int tr = it * 8 + wid, e = tr * K1_TW + lane;
        const unsigned k5 = s.kmask[0][tr], k4 = s.kmask[1][tr], k3 = s.kmask[2][tr], k2 = s.kmask[3][tr], bit = 1u << lane;
        if (s.imask[tr] & ~(k5 | k4 | k3 | k2) & bit) { const size_t gi = tile_base + (size_t)tr * W + lane; d.nrm[gi] = make_float4(qnan, qnan, qnan, 0.f); }
        if (k5 & bit) s.list[s.base[0][tr] + __popc(k5 & lt)] = (unsigned short)e;
        else if (k4 & bit) s.list[K1_TH * K1_TW - 1 - (s.base[1][tr] + __popc(k4 & lt))] = (unsigned short)e;
        else if ((k3 | k2) & bit) {
            const int pos = s.base[2][tr] + __popc((k3 | k2) & lt);
            if (pos < K1_TH * K1_TW / 2) s.list2[pos] = (unsigned short)(e | ((k3 & bit) ? (3 << 12) : (2 << 12)));
            else if (k3 & bit) k1_normal<3>(s, d, e, tile_base, W);          /* list full (pathological tile): in place */
            else k1_normal<2>(s, d, e, tile_base, W);
        }
    }
    __syncthreads();
    /* S5: box sums + normal.  The windows of 5 of an exact tile take separable sums, K1_BAND tile rows at a time: a thread per
     * (channel, gradient column) slides a 5-row column sum down the band in shared memory, then a thread per pixel adds five
     * neighbouring column sums per channel: exact sums, the same bits as the ordered sums of k1_normal<5>.  Windows of 4, 3, 2
     * (and of 5 in a tile without the exactness guarantee) take the ordered sums; they are handed out from the last thread
     * downwards, to the threads the first band leaves idle. */
    const int n5 = s.n5, n4 = s.n4, n32 = min(s.n32, K1_TH * K1_TW / 2);
    for (int b0 = 0; b0 < (exact ? (n5 ? K1_TH : K1_BAND0) : 0); b0 += (b0 ? K1_BAND : K1_BAND0)) {    /* CTA-uniform bounds */
        const int rows = b0 ? K1_BAND : K1_BAND0;                           /* bands of 8, 12, 12 tile rows */
        const int p0 = s.base[0][b0], p1 = b0 + rows < K1_TH ? s.base[0][b0 + rows] : n5;
        if (b0 > 0) {                                                       /* the first band's column sums are there already */
            __syncthreads();
            if (p1 > p0) k1_column_sums<K1_BAND>(s, csum, tid, b0);
            __syncthreads();
        }
        for (int i = p0 + tid; i < p1; i += K1_THREADS) {
            const int e = s.list[i], rr = e / K1_TW, cc = e - rr * K1_TW;
            const size_t gi = tile_base + (size_t)rr * W + cc;
            const float4 pw = d.xyzw[gi];
            const double *c = csum + (rr - b0) * K1_GW + cc;
            double sg[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const double *q = c + k * rows * K1_GW;
                sg[k] = q[0] + q[1] + q[2] + q[3] + q[4];
            }
            k1_finish(d, gi, pw.x, pw.y, pw.z, sg[0], sg[1], sg[2], sg[3], sg[4], sg[5]);
        }
        if (b0 == 0) {
            for (int i = K1_THREADS - 1 - tid; i < n4; i += K1_THREADS) k1_normal<4>(s, d, s.list[K1_TH * K1_TW - 1 - i], tile_base, W);
            for (int i = K1_THREADS - 1 - tid; i < n32; i += K1_THREADS) {
                const int code = s.list2[i];
                if ((code >> 12) == 3) k1_normal<3>(s, d, code & 0xfff, tile_base, W);
                else k1_normal<2>(s, d, code & 0xfff, tile_base, W);
            }
        }
    }
    if (!exact) {
        for (int i = tid; i < n5; i += K1_THREADS) k1_normal<5>(s, d, s.list[i], tile_base, W);
        for (int i = tid; i < n4; i += K1_THREADS) k1_normal<4>(s, d, s.list[K1_TH * K1_TW - 1 - i], tile_base, W);
        for (int i = tid; i < n32; i += K1_THREADS) {
            const int code = s.list2[i];
            if ((code >> 12) == 3) k1_normal<3>(s, d, code & 0xfff, tile_base, W);
            else k1_normal<2>(s, d, code & 0xfff, tile_base, W);
        }
    }
}

__global__ void __launch_bounds__(K1_THREADS, 4) k_ingest_normals_tma(SpDev d, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(16) unsigned char k1_raw[];
    K1Smem &s = *reinterpret_cast<K1Smem *>(k1_raw);
    const int f = blockIdx.z, W = d.W, H = d.H, N = d.N;
    const int r0 = blockIdx.y * K1_TH, c0 = blockIdx.x * K1_TW;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const float qnan = __int_as_float(0x7fc00000);
    const size_t tile_base = (size_t)f * N + (size_t)r0 * W + c0;

    /* S0: ONE TMA tensor copy brings the tile + halo of raw records (42 x 42 x 16 B; pixels outside the image arrive as NaN = invalid)
     * into the space the gradient fields use later.  A flat pass levels and masks every element into sz / sx / sy; after the
     * barrier the owned 32 x 32 block (one warp per tile row, lane = column) stores the levelled points, gathers the crop
     * statistics of passThroughCloudANDNormal / getMinMax3D, the cropped pixels per 1024-pixel block for the voxel stage and the
     * row masks. */
    float4 *raw = reinterpret_cast<float4 *>(&s.gx[0][0]);
    if (tid == 0) {
        st_mbar_init(&s.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        st_mbar_expect(&s.mbar, K1_RH * K1_RW * 16);
        st_tma_box(&tmap, raw, &s.mbar, c0 - K1_HALO, r0 - K1_HALO, f);
    }
    __syncthreads();
    bool risk = false;
    {
        float R[9];
        const bool rot = d.R != nullptr;
        if (rot) {
#pragma unroll
            for (int i = 0; i < 9; ++i) R[i] = __ldg(d.R + 9 * f + i);
        }
        st_mbar_wait(&s.mbar, 0u);
        for (int e = tid; e < K1_RH * K1_RW; e += K1_THREADS) {
            const float4 p = raw[e];
            float ox = p.x, oy = p.y, oz = p.z;
            if (rot) {
                if (!isnan(p.x)) {
                    ox = R[0] * p.x + R[1] * p.y + R[2] * p.z;
                    oy = R[3] * p.x + R[4] * p.y + R[5] * p.z;
                    oz = R[6] * p.x + R[7] * p.y + R[8] * p.z;
                } else { ox = oy = oz = qnan; }
            }
            /* remove_nan_points: r, g or b == 0 marks an invalid pixel (per-byte compare of the three colour bytes) */
            if (__vcmpeq4(__float_as_uint(p.w) | 0xff000000u, 0u)) { ox = oy = oz = qnan; }
            s.sz[e] = oz;
            risk |= k1_inexact_risk(ox) | k1_inexact_risk(oy) | k1_inexact_risk(oz);
            const int rr = e / K1_RW, cc = e - rr * K1_RW;
            const unsigned xr = (unsigned)(rr - 2), xc = (unsigned)(cc - 2);       /* x, y keep a halo of 3 */
            if (xr < (unsigned)K1_XH && xc < (unsigned)K1_XW) { s.sx[xr * K1_XW + xc] = ox; s.sy[xr * K1_XW + xc] = oy; }
        }
    }
    const bool exact = __syncthreads_or(risk) == 0;      /* box sums of this tile are exact in double in any order */
    {
        float mnx = FLT_MAX, mny = FLT_MAX, mnz = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX, mxz = -FLT_MAX;
        int nvalid = 0, ncrop = 0;
        const int gc = c0 + lane;
#pragma unroll
        for (int it = 0; it < K1_TH / 8; ++it) {
            const int tr = it * 8 + wid, gr = r0 + tr;
            const bool inside = gr < H && gc < W;
            const float ox = s.sx[(tr + 3) * K1_XW + lane + 3], oy = s.sy[(tr + 3) * K1_XW + lane + 3], oz = s.sz[(tr + K1_HALO) * K1_RW + lane + K1_HALO];
            bool crop = false;
            if (inside) {
                d.xyzw[tile_base + (size_t)tr * W + lane] = make_float4(ox, oy, oz, 0.f);
                if (sp_finite3(ox, oy, oz)) {
                    ++nvalid;
                    if (!(ox < d.c.x0 || ox > d.c.x1) && !(oy < d.c.y0 || oy > d.c.y1) && !(oz < d.c.z0 || oz > d.c.z1)) {
                        crop = true; ++ncrop;
                        mnx = fminf(mnx, ox); mny = fminf(mny, oy); mnz = fminf(mnz, oz);
                        mxx = fmaxf(mxx, ox); mxy = fmaxf(mxy, oy); mxz = fmaxf(mxz, oz);
                    }
                }
            }
            const unsigned bc = __ballot_sync(0xffffffffu, crop);
            const unsigned bi = __ballot_sync(0xffffffffu, inside);
            const unsigned bv = __ballot_sync(0xffffffffu, inside && gr >= 5 && gr < H - 5 && gc >= 5 && gc < W - 5 && sp_finite(oz));
            if (lane == 0) {
                s.imask[tr] = bi; s.vmask[tr] = bv;
                if (bc) {
                    const int i0 = gr * W + c0, left = 1024 - (i0 & 1023);               /* pixels of this row inside i0's block */
                    const unsigned lo = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
                    int *vc = d.vcnt + (size_t)f * d.nblk + (i0 >> 10);
                    if (bc & lo) atomicAdd(vc, __popc(bc & lo));
                    if (bc & ~lo) atomicAdd(vc + 1, __popc(bc & ~lo));
                }
            }
        }
        nvalid = __reduce_add_sync(0xffffffffu, nvalid); ncrop = __reduce_add_sync(0xffffffffu, ncrop);
        if (ncrop) {                                                       /* warp-uniform */
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

    /* S1 + S2: depth-change flags on the lateral coordinate z (reference quirk), halo 4, as bit rows, four flag rows per warp
     * pass: columns 0..31 of each row (the ballot is the row), then the eight tail columns of the four rows in one go; lanes
     * 0..3 then dilate "their" row: a flagged pixel within h columns, h = 0..4 (PCL's distance map along the row, capped at 5) */
    for (int q = wid; q < K1_FH / 4; q += 8) {
        unsigned long long m = 0ull;
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const int fr = q * 4 + (u < 4 ? u : (lane >> 3));
            const int se = (fr + 1) * K1_RW + (u < 4 ? lane : 32 + (lane & 7)) + 1;
            const float zc = s.sz[se];
            const bool z0 = k1_bad_pair(zc, s.sz[se + 1]) || k1_bad_pair(s.sz[se - 1], zc) ||
                            k1_bad_pair(zc, s.sz[se + K1_RW]) || k1_bad_pair(s.sz[se - K1_RW], zc);
            const unsigned b = __ballot_sync(0xffffffffu, z0);
            if (u < 4) { if (lane == u) m = b; }
            else m |= (unsigned long long)((b >> (8 * (lane & 3))) & 255u) << 32;
        }
        if (lane < 4) {
            const int fr = q * 4 + lane;
            s.dil[0][fr] = m;
#pragma unroll
            for (int h = 1; h < 5; ++h) { m |= (m << 1) | (m >> 1); s.dil[h][fr] = m; }
        }
    }
    /* S4 (independent of S1): gradient fields, halo 2, same row-aligned passes: columns 0..31 of a gradient row per warp, then
     * the four tail columns of eight rows per warp; the ballots are the rows of finite elements */
    for (int q = wid; q < K1_GH + (K1_GH + 7) / 8; q += 8) {
        const bool tail = q >= K1_GH;
        const int rr = tail ? (q - K1_GH) * 8 + (lane >> 2) : q, cc = tail ? 32 + (lane & 3) : lane;
        bool fa = false, fb = false;
        if (rr < K1_GH) {
            const int e = rr * K1_GW + cc;
            const int se = (rr + 3) * K1_RW + cc + 3, xe = (rr + 1) * K1_XW + cc + 1;
            const float ax = s.sx[xe + 1] - s.sx[xe - 1], ay = s.sy[xe + 1] - s.sy[xe - 1], az = s.sz[se + 1] - s.sz[se - 1];
            const float bx = s.sx[xe + K1_XW] - s.sx[xe - K1_XW], by = s.sy[xe + K1_XW] - s.sy[xe - K1_XW], bz = s.sz[se + K1_RW] - s.sz[se - K1_RW];
            fa = sp_finite(ax + (ay + az)); fb = sp_finite(bx + (by + bz));
            s.gx[0][e] = fa ? ax : 0.0f; s.gx[1][e] = fa ? ay : 0.0f; s.gx[2][e] = fa ? az : 0.0f;
            s.gy[0][e] = fb ? bx : 0.0f; s.gy[1][e] = fb ? by : 0.0f; s.gy[2][e] = fb ? bz : 0.0f;
        }
        const unsigned ba = __ballot_sync(0xffffffffu, fa), bb = __ballot_sync(0xffffffffu, fb);
        if (!tail) { if (lane == 0) { s.falo[rr] = ba; s.fblo[rr] = bb; } }
        else if ((lane & 3) == 0 && rr < K1_GH) { s.fahi[rr] = (ba >> (lane & 28)) & 15u; s.fbhi[rr] = (bb >> (lane & 28)) & 15u; }
    }
    __syncthreads();
    /* S3: smoothing size per owned pixel, one thread per tile row.  A flagged pixel at row distance a and column distance h
     * allows a window of kcat[a][h] (chamfer 1.0/1.4: d = 1.4*min + (max-min); d<=2 -> 0 (NaN), (2,3) -> 2, [3,4) -> 3, [4,5) -> 4,
     * >=5 -> 5):  a=0: 0 0 0 3 4 5 | a=1: 0 0 2 3 4 5 | a=2: 0 2 2 3 4 5 | a=3: 3 3 3 4 5 5 | a=4: 4 4 4 5 5 5  (h = 0..5).
     * "window <= k" is therefore an OR of dilated rows; the list offsets of each row come from a warp scan of the counts.
     * Meanwhile warps 1..7 of an exact tile start on the separable sums of S5 (column sums of the first band): the coordinates
     * in sz / sx / sy had their last readers before the barrier above, the band may overwrite them. */
    double *csum = reinterpret_cast<double *>(k1_raw);                      /* [6][band rows][K1_GW], over sz / sx / sy (+ dil, fa / fb rows) */
    if (exact && tid >= 32) k1_column_sums<K1_BAND0>(s, csum, tid - 32, 0);
    if (tid < K1_TH) {
        const int r = tid + 4;
        const unsigned long long (*D)[K1_FH] = s.dil;
        const unsigned long long z0 = D[2][r] | D[1][r - 1] | D[1][r + 1] | D[0][r - 2] | D[0][r + 2];
        const unsigned long long z2 = D[2][r] | D[2][r - 1] | D[2][r + 1] | D[2][r - 2] | D[2][r + 2];
        const unsigned long long z3 = D[3][r] | D[3][r - 1] | D[3][r + 1] | D[3][r - 2] | D[3][r + 2] | D[2][r - 3] | D[2][r + 3];
        const unsigned long long z4 = D[4][r] | D[4][r - 1] | D[4][r + 1] | D[4][r - 2] | D[4][r + 2] | D[3][r - 3] | D[3][r + 3] | D[2][r - 4] | D[2][r + 4];
        /* a normal needs one finite element in each gradient window (PCL's finite-element counts): rows tid + a of the gradient
         * region, columns cc + b; windows of 5: a, b = 0..4; 4: 0..3; 3: 1..3; 2: 1..2 */
        unsigned long long fa[5], fb[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            fa[a] = (unsigned long long)s.falo[tid + a] | ((unsigned long long)s.fahi[tid + a] << 32);
            fb[a] = (unsigned long long)s.fblo[tid + a] | ((unsigned long long)s.fbhi[tid + a] << 32);
        }
        const unsigned long long a2 = fa[1] | fa[2], a3 = a2 | fa[3], a4 = a3 | fa[0], a5 = a4 | fa[4];
        const unsigned long long b2 = fb[1] | fb[2], b3 = b2 | fb[3], b4 = b3 | fb[0], b5 = b4 | fb[4];
        const unsigned w2 = (unsigned)((a2 >> 1) | (a2 >> 2)) & (unsigned)((b2 >> 1) | (b2 >> 2));
        const unsigned w3 = (unsigned)((a3 >> 1) | (a3 >> 2) | (a3 >> 3)) & (unsigned)((b3 >> 1) | (b3 >> 2) | (b3 >> 3));
        const unsigned w4 = (unsigned)(a4 | (a4 >> 1) | (a4 >> 2) | (a4 >> 3)) & (unsigned)(b4 | (b4 >> 1) | (b4 >> 2) | (b4 >> 3));
        const unsigned w5 = (unsigned)(a5 | (a5 >> 1) | (a5 >> 2) | (a5 >> 3) | (a5 >> 4)) & (unsigned)(b5 | (b5 >> 1) | (b5 >> 2) | (b5 >> 3) | (b5 >> 4));
        const unsigned v = s.vmask[tid];
        const unsigned k5 = v & w5 & ~(unsigned)(z4 >> 4), k4 = v & w4 & (unsigned)((z4 & ~z3) >> 4), k3 = v & w3 & (unsigned)((z3 & ~z2) >> 4),
                       k2 = v & w2 & (unsigned)((z2 & ~z0) >> 4);
        s.kmask[0][tid] = k5; s.kmask[1][tid] = k4; s.kmask[2][tid] = k3; s.kmask[3][tid] = k2;
        int c5 = __popc(k5), c4 = __popc(k4), c32 = __popc(k3 | k2);
        int i5 = c5, i4 = c4, i32 = c32;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t5 = __shfl_up_sync(0xffffffffu, i5, o), t4 = __shfl_up_sync(0xffffffffu, i4, o), t32 = __shfl_up_sync(0xffffffffu, i32, o);
            if (lane >= o) { i5 += t5; i4 += t4; i32 += t32; }
        }
        s.base[0][tid] = (unsigned short)(i5 - c5); s.base[1][tid] = (unsigned short)(i4 - c4); s.base[2][tid] = (unsigned short)(i32 - c32);
        if (tid == K1_TH - 1) { s.n5 = i5; s.n4 = i4; s.n32 = i32; }
    }
    __syncthreads();
    /* Pixels without a normal (invalid, border, next to a depth discontinuity: about a third of the tile) get their NaN here;
     * the others go into one list per window size, so that the box sums below run with full warps and fixed trip counts. */
#pragma unroll
    for (int it = 0; it < K1_TH / 8; ++it) {
        const int tr = it * 8 + wid, e = tr * K1_TW + lane;
        const unsigned k5 = s.kmask[0][tr], k4 = s.kmask[1][tr], k3 = s.kmask[2][tr], k2 = s.kmask[3][tr], bit = 1u << lane;
        if (s.imask[tr] & ~(k5 | k4 | k3 | k2) & bit) { const size_t gi = tile_base + (size_t)tr * W + lane; d.nrm[gi] = make_float4(qnan, qnan, qnan, 0.f); }
        if (k5 & bit) s.list[s.base[0][tr] + __popc(k5 & lt)] = (unsigned short)e;
        else if (k4 & bit) s.list[K1_TH * K1_TW - 1 - (s.base[1][tr] + __popc(k4 & lt))] = (unsigned short)e;
        else if ((k3 | k2) & bit) {
            const int pos = s.base[2][tr] + __popc((k3 | k2) & lt);
            if (pos < K1_TH * K1_TW / 2) s.list2[pos] = (unsigned short)(e | ((k3 & bit) ? (3 << 12) : (2 << 12)));
            else if (k3 & bit) k1_normal<3, true>(s, d, e, tile_base, W);          /* list full (pathological tile): in place */
            else k1_normal<2, true>(s, d, e, tile_base, W);
        }
    }
    __syncthreads();
    /* S5: box sums + normal.  The windows of 5 of an exact tile take separable sums, K1_BAND tile rows at a time: a thread per
     * (channel, gradient column) slides a 5-row column sum down the band in shared memory, then a thread per pixel adds five
     * neighbouring column sums per channel: exact sums, the same bits as the ordered sums of k1_normal<5, true>.  Windows of 4, 3, 2
     * (and of 5 in a tile without the exactness guarantee) take the ordered sums; they are handed out from the last thread
     * downwards, to the threads the first band leaves idle. */
    const int n5 = s.n5, n4 = s.n4, n32 = min(s.n32, K1_TH * K1_TW / 2);
    for (int b0 = 0; b0 < (exact ? (n5 ? K1_TH : K1_BAND0) : 0); b0 += (b0 ? K1_BAND : K1_BAND0)) {    /* CTA-uniform bounds */
        const int rows = b0 ? K1_BAND : K1_BAND0;                           /* bands of 8, 12, 12 tile rows */
        const int p0 = s.base[0][b0], p1 = b0 + rows < K1_TH ? s.base[0][b0 + rows] : n5;
        if (b0 > 0) {                                                       /* the first band's column sums are there already */
            __syncthreads();
            if (p1 > p0) k1_column_sums<K1_BAND>(s, csum, tid, b0);
            __syncthreads();
        }
        for (int i = p0 + tid; i < p1; i += K1_THREADS) {
            const int e = s.list[i], rr = e / K1_TW, cc = e - rr * K1_TW;
            const size_t gi = tile_base + (size_t)rr * W + cc;
            const float4 pw = d.xyzw[gi];
            const double *c = csum + (rr - b0) * K1_GW + cc;
            double sg[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const double *q = c + k * rows * K1_GW;
                sg[k] = q[0] + q[1] + q[2] + q[3] + q[4];
            }
            st_finish(d, gi, pw.x, pw.y, pw.z, sg[0], sg[1], sg[2], sg[3], sg[4], sg[5]);
        }
        if (b0 == 0) {
            for (int i = K1_THREADS - 1 - tid; i < n4; i += K1_THREADS) k1_normal<4, true>(s, d, s.list[K1_TH * K1_TW - 1 - i], tile_base, W);
            for (int i = K1_THREADS - 1 - tid; i < n32; i += K1_THREADS) {
                const int code = s.list2[i];
                if ((code >> 12) == 3) k1_normal<3, true>(s, d, code & 0xfff, tile_base, W);
                else k1_normal<2, true>(s, d, code & 0xfff, tile_base, W);
            }
        }
    }
    if (!exact) {
        for (int i = tid; i < n5; i += K1_THREADS) k1_normal<5, true>(s, d, s.list[i], tile_base, W);
        for (int i = tid; i < n4; i += K1_THREADS) k1_normal<4, true>(s, d, s.list[K1_TH * K1_TW - 1 - i], tile_base, W);
        for (int i = tid; i < n32; i += K1_THREADS) {
            const int code = s.list2[i];
            if ((code >> 12) == 3) k1_normal<3, true>(s, d, code & 0xfff, tile_base, W);
            else k1_normal<2, true>(s, d, code & 0xfff, tile_base, W);
        }
    }
}

/* ==================================================================================================== */
__global__ void k_init_frames(SpDev d, int nB)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nB * SP_NCOUNT) d.counts[i] = 0;
    if (i < nB * 8) { const int k = i & 7; d.mm[i] = (k < 3) ? 0x7fffffff : (int)0x80000000; }
    if (i < nB) d.flags[i] = 0;
    for (int q = i; q < nB * d.nblk; q += gridDim.x * blockDim.x) d.vcnt[q] = 0;
    if (i < 2 * nB) { d.seg_n[i] = 0; d.n_clusters[i] = 0; d.n_cand[i] = 0; d.log_n[i] = 0; d.n_seg[i] = 0; d.n_merged[i] = 0; }
}

/* ==================================================================================================== */
/* Depth-image ingest: K2G::updateCloud (/root/reference/modules/k2g/k2g.cpp:231-308) with the pinhole maps of
 * K2G::prepareMake3D (:506-520), on the device: (undistorted depth in mm, registered colour) -> the 16-byte cloud records
 * the front-end reads.  depth_type 0 = float32 mm (libfreenect2's undistorted frame), 1 = uint16 mm (the saved depth PNG,
 * test/reader.cpp:128).  bgrx = registered colour, 4 bytes per pixel (b, g, r, x) or NULL (every valid pixel white).
 * mirror = the horizontal flip of K2G(true) (k2g.cpp:247-253, test/reader.cpp:9). */
__global__ void __launch_bounds__(256) k_depth_to_records(const void *depth, int depth_type, const unsigned char *bgrx, const float *colmap,
                                                          const float *rowmap, int W, int H, int mirror, float4 *out)
{
    const int f = blockIdx.y, N = W * H;
    const size_t base = (size_t)f * N;
    const float qnan = __int_as_float(0x7fc00000);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const int y = i / W, x = i - y * W;
        const int sx = mirror ? (W - 1 - x) : x;
        const size_t si = base + (size_t)y * W + sx;
        const float dmm = depth_type == 0 ? reinterpret_cast<const float *>(depth)[si] : (float)reinterpret_cast<const unsigned short *>(depth)[si];
        const float dv = dmm / 1000.0f;
        float4 o;
        if (!isnan(dv) && !(fabsf(dv) < 0.0001)) {
            o.x = colmap[x] * dv; o.y = rowmap[y] * dv; o.z = dv;
            unsigned rgba = 0xFFFFFFFFu;
            if (bgrx) { const uchar4 c = reinterpret_cast<const uchar4 *>(bgrx)[si]; rgba = (unsigned)c.x | ((unsigned)c.y << 8) | ((unsigned)c.z << 16) | 0xFF000000u; }
            o.w = __uint_as_float(rgba);
        } else {
            o.x = o.y = o.z = qnan;
            o.w = __uint_as_float(0u);                /* b = g = r = (uint8)NaN: 0 on x86, and alpha is left unset */
        }
        out[base + i] = o;
    }
}
