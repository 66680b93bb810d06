/*
 * sp_front.cuh -- streaming kernels of the front-end: levelling + masking + organised normals (K1K2),
 * voxel keys, voxel representative pick + point classification, list scatter.
 */
#pragma once
#include "sp_common.cuh"

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
#define K1_RW (K1_TW + 2 * K1_HALO)      /* 42 */
#define K1_RH (K1_TH + 2 * K1_HALO)
#define K1_FW (K1_TW + 8)                /* zero-flag region: halo 4 */
#define K1_FH (K1_TH + 8)
#define K1_GW (K1_TW + 4)                /* gradient region: halo 2 */
#define K1_GH (K1_TH + 4)
#define K1_THREADS 256

struct K1Smem {
    float sx[K1_RH * K1_RW], sy[K1_RH * K1_RW], sz[K1_RH * K1_RW];
    float gx[3][K1_GH * K1_GW], gy[3][K1_GH * K1_GW];
    unsigned char zero[K1_FH * K1_FW];
    unsigned char hd[K1_FH * K1_TW];
    unsigned char fx[K1_GH * K1_GW], fy[K1_GH * K1_GW];
    float onx[K1_TH * K1_TW], ony[K1_TH * K1_TW], onz[K1_TH * K1_TW];   /* normals of the tile, stored coalesced at the end */
    unsigned short list[K1_TH * K1_TW];   /* pixels that get a normal: 5 x 5 windows from the front, smaller ones from the back */
    int n5, nsmall;
    int red[6];
};

__device__ __forceinline__ bool k1_bad_pair(float za, float zb)
{
    /* depth-change test of the pair (a = upper/left element), integral_image_normal.hpp */
    const float thr = (0.02f * (fabsf(za) + 1.0f) * 2.0f);
    return (fabsf(za - zb) > thr) || !sp_finite(za) || !sp_finite(zb);
}

__global__ void __launch_bounds__(K1_THREADS) k_ingest_normals(SpDev d)
{
    extern __shared__ __align__(16) unsigned char k1_raw[];
    K1Smem &s = *reinterpret_cast<K1Smem *>(k1_raw);
    const int f = blockIdx.z, W = d.W, H = d.H, N = d.N;
    const int r0 = blockIdx.y * K1_TH, c0 = blockIdx.x * K1_TW;
    const int tid = threadIdx.x;
    const float4 *in = d.in + (size_t)f * N;
    float R[9];
    const bool rot = d.R != nullptr;
    if (rot) {
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(d.R + 9 * f + i);
    }
    const float qnan = __int_as_float(0x7fc00000);

    /* S0: load tile + halo, level, mask */
    for (int e = tid; e < K1_RH * K1_RW; e += K1_THREADS) {
        const int rr = e / K1_RW, cc = e - rr * K1_RW;
        const int gr = r0 - K1_HALO + rr, gc = c0 - K1_HALO + cc;
        float ox = qnan, oy = qnan, oz = qnan;
        if (gr >= 0 && gr < H && gc >= 0 && gc < W) {
            const float4 p = __ldg(in + (size_t)gr * W + gc);
            ox = p.x; oy = p.y; oz = p.z;
            if (rot) {
                if (!isnan(p.x)) {
                    ox = R[0] * p.x + R[1] * p.y + R[2] * p.z;
                    oy = R[3] * p.x + R[4] * p.y + R[5] * p.z;
                    oz = R[6] * p.x + R[7] * p.y + R[8] * p.z;
                } else { ox = oy = oz = qnan; }
            }
            const unsigned rgba = __float_as_uint(p.w);
            if ((rgba & 255u) == 0u || ((rgba >> 8) & 255u) == 0u || ((rgba >> 16) & 255u) == 0u) { ox = oy = oz = qnan; }
        }
        s.sx[e] = ox; s.sy[e] = oy; s.sz[e] = oz;
    }
    __syncthreads();

    /* owned pixels: write levelled xyz, crop predicate, bounding box, counts */
    {
        int mnx = 0x7fffffff, mny = 0x7fffffff, mnz = 0x7fffffff, mxx = (int)0x80000000, mxy = (int)0x80000000, mxz = (int)0x80000000;
        int nvalid = 0, ncrop = 0;
        for (int e = tid; e < K1_TH * K1_TW; e += K1_THREADS) {
            const int rr = e / K1_TW, cc = e - rr * K1_TW;
            const int gr = r0 + rr, gc = c0 + cc;
            if (gr < H && gc < W) {
                const int se = (rr + K1_HALO) * K1_RW + cc + K1_HALO;
                const float px = s.sx[se], py = s.sy[se], pz = s.sz[se];
                const size_t gi = (size_t)f * N + (size_t)gr * W + gc;
                d.x[gi] = px; d.y[gi] = py; d.z[gi] = pz;
                if (sp_finite3(px, py, pz)) {
                    ++nvalid;
                    if (!(px < d.c.x0 || px > d.c.x1) && !(py < d.c.y0 || py > d.c.y1) && !(pz < d.c.z0 || pz > d.c.z1)) {
                        ++ncrop;
                        const int ox = sp_f2ord(px), oy = sp_f2ord(py), oz = sp_f2ord(pz);
                        mnx = min(mnx, ox); mny = min(mny, oy); mnz = min(mnz, oz);
                        mxx = max(mxx, ox); mxy = max(mxy, oy); mxz = max(mxz, oz);
                    }
                }
            }
        }
        mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny); mnz = __reduce_min_sync(0xffffffffu, mnz);
        mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy); mxz = __reduce_max_sync(0xffffffffu, mxz);
        nvalid = __reduce_add_sync(0xffffffffu, nvalid); ncrop = __reduce_add_sync(0xffffffffu, ncrop);
        if ((tid & 31) == 0 && nvalid) {
            atomicAdd(&d.counts[f * SP_NCOUNT + 1], nvalid);
            if (ncrop) {
                atomicAdd(&d.counts[f * SP_NCOUNT + 2], ncrop);
                int *mm = d.mm + f * 8;
                atomicMin(mm + 0, mnx); atomicMin(mm + 1, mny); atomicMin(mm + 2, mnz);
                atomicMax(mm + 3, mxx); atomicMax(mm + 4, mxy); atomicMax(mm + 5, mxz);
            }
        }
    }

    /* S1: depth-change flags on the lateral coordinate z (reference quirk), halo 4 */
    for (int e = tid; e < K1_FH * K1_FW; e += K1_THREADS) {
        const int rr = e / K1_FW, cc = e - rr * K1_FW;
        const int se = (rr + 1) * K1_RW + cc + 1;
        const float zc = s.sz[se];
        const bool z0 = k1_bad_pair(zc, s.sz[se + 1]) || k1_bad_pair(s.sz[se - 1], zc) ||
                        k1_bad_pair(zc, s.sz[se + K1_RW]) || k1_bad_pair(s.sz[se - K1_RW], zc);
        s.zero[e] = z0 ? 1 : 0;
    }
    /* S4 (independent of S1): gradient fields, halo 2 */
    for (int e = tid; e < K1_GH * K1_GW; e += K1_THREADS) {
        const int rr = e / K1_GW, cc = e - rr * K1_GW;
        const int se = (rr + 3) * K1_RW + cc + 3;
        const float ax = s.sx[se + 1] - s.sx[se - 1], ay = s.sy[se + 1] - s.sy[se - 1], az = s.sz[se + 1] - s.sz[se - 1];
        const float bx = s.sx[se + K1_RW] - s.sx[se - K1_RW], by = s.sy[se + K1_RW] - s.sy[se - K1_RW], bz = s.sz[se + K1_RW] - s.sz[se - K1_RW];
        const bool fa = sp_finite(ax + (ay + az)), fb = sp_finite(bx + (by + bz));
        s.gx[0][e] = fa ? ax : 0.0f; s.gx[1][e] = fa ? ay : 0.0f; s.gx[2][e] = fa ? az : 0.0f;
        s.gy[0][e] = fb ? bx : 0.0f; s.gy[1][e] = fb ? by : 0.0f; s.gy[2][e] = fb ? bz : 0.0f;
        s.fx[e] = fa ? 1 : 0; s.fy[e] = fb ? 1 : 0;
    }
    __syncthreads();
    /* S2: horizontal distance to the nearest flagged pixel, capped at 5 */
    for (int e = tid; e < K1_FH * K1_TW; e += K1_THREADS) {
        const int rr = e / K1_TW, cc = e - rr * K1_TW;
        const unsigned char *row = &s.zero[rr * K1_FW + cc + 4];
        int h = 5;
        if (row[0]) h = 0;
        else if (row[-1] | row[1]) h = 1;
        else if (row[-2] | row[2]) h = 2;
        else if (row[-3] | row[3]) h = 3;
        else if (row[-4] | row[4]) h = 4;
        s.hd[e] = (unsigned char)h;
    }
    __syncthreads();

    /* S3: smoothing size per owned pixel.  Pixels without a normal (invalid, border, next to a depth discontinuity) are
     * about a third of the tile: the pixels that do get one are compacted into a list first, so that the box sums below run
     * with full warps.  kcat[a][h]: rect size implied by a flagged pixel at row distance a, column distance h
     * (chamfer 1.0/1.4: d = 1.4*min + (max-min); d<=2 -> 0 (NaN), (2,3) -> 2, [3,4) -> 3, [4,5) -> 4, >=5 -> 5) */
    if (tid == 0) { s.n5 = 0; s.nsmall = 0; }
    __syncthreads();
    const unsigned kcat0 = 0x543000u, kcat1 = 0x543200u, kcat2 = 0x543220u, kcat3 = 0x554333u, kcat4 = 0x555444u;
    for (int e = tid; e < K1_TH * K1_TW; e += K1_THREADS) {
        const int rr = e / K1_TW, cc = e - rr * K1_TW;
        const int gr = r0 + rr, gc = c0 + cc;
        s.onx[e] = qnan; s.ony[e] = qnan; s.onz[e] = qnan;
        int k = 0;
        if (gr < H && gc < W && gr >= 5 && gr < H - 5 && gc >= 5 && gc < W - 5 && sp_finite(s.sz[(rr + K1_HALO) * K1_RW + cc + K1_HALO])) {
            const unsigned char *col = &s.hd[(rr + 4) * K1_TW + cc];
            k = 5;
            k = min(k, (int)((kcat0 >> (4 * col[0])) & 15u));
            k = min(k, (int)((kcat1 >> (4 * col[-K1_TW])) & 15u));
            k = min(k, (int)((kcat1 >> (4 * col[K1_TW])) & 15u));
            k = min(k, (int)((kcat2 >> (4 * col[-2 * K1_TW])) & 15u));
            k = min(k, (int)((kcat2 >> (4 * col[2 * K1_TW])) & 15u));
            k = min(k, (int)((kcat3 >> (4 * col[-3 * K1_TW])) & 15u));
            k = min(k, (int)((kcat3 >> (4 * col[3 * K1_TW])) & 15u));
            k = min(k, (int)((kcat4 >> (4 * col[-4 * K1_TW])) & 15u));
            k = min(k, (int)((kcat4 >> (4 * col[4 * K1_TW])) & 15u));
        }
        /* warp-aggregated appends: one shared atomic per warp and list */
        const unsigned lanebit = 1u << (tid & 31), lt = lanebit - 1u;
        const unsigned b5 = __ballot_sync(0xffffffffu, k == 5), bs = __ballot_sync(0xffffffffu, k >= 2 && k < 5);
        if (b5) {
            int base = 0;
            if ((b5 & lt) == 0 && (b5 & lanebit)) base = atomicAdd(&s.n5, __popc(b5));
            base = __shfl_sync(0xffffffffu, base, __ffs(b5) - 1);
            if (k == 5) s.list[base + __popc(b5 & lt)] = (unsigned short)e;
        }
        if (bs) {
            int base = 0;
            if ((bs & lt) == 0 && (bs & lanebit)) base = atomicAdd(&s.nsmall, __popc(bs));
            base = __shfl_sync(0xffffffffu, base, __ffs(bs) - 1);
            if (k >= 2 && k < 5) s.list[K1_TH * K1_TW - 1 - (base + __popc(bs & lt))] = (unsigned short)(e | (k << 12));
        }
    }
    __syncthreads();
    /* S5: box sums + normal.  Sum order (the oracle's): rows top to bottom of (row sums left to right), all in double. */
    auto finish = [&](int e, double sgx0, double sgx1, double sgx2, double sgy0, double sgy1, double sgy2, int cntx, int cnty) {
        if (cntx != 0 && cnty != 0) {
            const double n0 = sgy1 * sgx2 - sgy2 * sgx1;
            const double n1 = sgy2 * sgx0 - sgy0 * sgx2;
            const double n2 = sgy0 * sgx1 - sgy1 * sgx0;
            const double L = n0 * n0 + (n1 * n1 + n2 * n2);
            if (L != 0.0) {
                const int rr = e / K1_TW, cc = e - rr * K1_TW;
                const int se = (rr + K1_HALO) * K1_RW + cc + K1_HALO;
                const float px = s.sx[se], py = s.sy[se], pz = s.sz[se];
                const double sq = sqrt(L);
                float fx = (float)(n0 / sq), fy = (float)(n1 / sq), fz = (float)(n2 / sq);
                const float vx = 0.0f - px, vy = 0.0f - py, vz = 0.0f - pz;
                const float cs = (vx * fx + vy * fy + vz * fz);
                if (cs < 0) { fx *= -1; fy *= -1; fz *= -1; }
                s.onx[e] = fx; s.ony[e] = fy; s.onz[e] = fz;
            }
        }
    };
    const int n5 = s.n5, nsmall = s.nsmall;
    for (int i = tid; i < n5; i += K1_THREADS) {                 /* the 5 x 5 windows: fixed trip counts */
        const int e = s.list[i];
        const int rr = e / K1_TW, cc = e - rr * K1_TW;
        const int g0 = rr * K1_GW + cc;                          /* (rr + 2 - 2, cc + 2 - 2) */
        double sgx0 = 0.0, sgx1 = 0.0, sgx2 = 0.0, sgy0 = 0.0, sgy1 = 0.0, sgy2 = 0.0;
        int cntx = 0, cnty = 0;
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            double rx0 = 0.0, rx1 = 0.0, rx2 = 0.0, ry0 = 0.0, ry1 = 0.0, ry2 = 0.0;
            const int gb = g0 + a * K1_GW;
#pragma unroll
            for (int b = 0; b < 5; ++b) {
                const int ge = gb + b;
                rx0 += (double)s.gx[0][ge]; rx1 += (double)s.gx[1][ge]; rx2 += (double)s.gx[2][ge];
                ry0 += (double)s.gy[0][ge]; ry1 += (double)s.gy[1][ge]; ry2 += (double)s.gy[2][ge];
                cntx += s.fx[ge]; cnty += s.fy[ge];
            }
            sgx0 += rx0; sgx1 += rx1; sgx2 += rx2; sgy0 += ry0; sgy1 += ry1; sgy2 += ry2;
        }
        finish(e, sgx0, sgx1, sgx2, sgy0, sgy1, sgy2, cntx, cnty);
    }
    for (int i = tid; i < nsmall; i += K1_THREADS) {             /* windows of 2, 3, 4 */
        const int code = s.list[K1_TH * K1_TW - 1 - i];
        const int e = code & 0xfff, k = code >> 12;
        const int rr = e / K1_TW, cc = e - rr * K1_TW;
        const int half = k >> 1;
        const int g0 = (rr + 2 - half) * K1_GW + (cc + 2 - half);
        double sgx0 = 0.0, sgx1 = 0.0, sgx2 = 0.0, sgy0 = 0.0, sgy1 = 0.0, sgy2 = 0.0;
        int cntx = 0, cnty = 0;
        for (int a = 0; a < k; ++a) {
            double rx0 = 0.0, rx1 = 0.0, rx2 = 0.0, ry0 = 0.0, ry1 = 0.0, ry2 = 0.0;
            const int gb = g0 + a * K1_GW;
            for (int b = 0; b < k; ++b) {
                const int ge = gb + b;
                rx0 += (double)s.gx[0][ge]; rx1 += (double)s.gx[1][ge]; rx2 += (double)s.gx[2][ge];
                ry0 += (double)s.gy[0][ge]; ry1 += (double)s.gy[1][ge]; ry2 += (double)s.gy[2][ge];
                cntx += s.fx[ge]; cnty += s.fy[ge];
            }
            sgx0 += rx0; sgx1 += rx1; sgx2 += rx2; sgy0 += ry0; sgy1 += ry1; sgy2 += ry2;
        }
        finish(e, sgx0, sgx1, sgx2, sgy0, sgy1, sgy2, cntx, cnty);
    }
    __syncthreads();
    for (int e = tid; e < K1_TH * K1_TW; e += K1_THREADS) {
        const int rr = e / K1_TW, cc = e - rr * K1_TW;
        const int gr = r0 + rr, gc = c0 + cc;
        if (gr >= H || gc >= W) continue;
        const size_t gi = (size_t)f * N + (size_t)gr * W + gc;
        d.nx[gi] = s.onx[e]; d.ny[gi] = s.ony[e]; d.nz[gi] = s.onz[e];
    }
}

/* ==================================================================================================== */
__global__ void k_init_frames(SpDev d, int nB)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nB * SP_NCOUNT) d.counts[i] = 0;
    if (i < nB * 8) { const int k = i & 7; d.mm[i] = (k < 3) ? 0x7fffffff : (int)0x80000000; }
    if (i < nB) d.flags[i] = 0;
    if (i < 2 * nB) { d.seg_n[i] = 0; d.n_clusters[i] = 0; d.n_cand[i] = 0; d.log_n[i] = 0; d.n_seg[i] = 0; d.n_merged[i] = 0; }
}

/* voxel grid of each frame: getMinMax3D result -> min_b, div_b, divb_mul (voxel_grid_indices.hpp:292-318)
 * plus the "<100 points" guard of process() (StairPerception.cpp:56). */
__global__ void k_frame_grid(SpDev d, int nB)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nB) return;
    SpFrameGrid g;
    g.status = 0; g.pad = 0;
    const int ncrop = d.counts[f * SP_NCOUNT + 2];
    for (int k = 0; k < 3; ++k) { g.minb[k] = 0; g.mul[k] = 0; }
    if (ncrop < 100) g.status = 1;
    else {
        const int *mm = d.mm + f * 8;
        const float mnx = sp_ord2f(mm[0]), mny = sp_ord2f(mm[1]), mnz = sp_ord2f(mm[2]);
        const float mxx = sp_ord2f(mm[3]), mxy = sp_ord2f(mm[4]), mxz = sp_ord2f(mm[5]);
        const long long dx = (long long)((mxx - mnx) * d.c.ivx) + 1, dy = (long long)((mxy - mny) * d.c.ivy) + 1,
                        dz = (long long)((mxz - mnz) * d.c.ivz) + 1;
        if (dx * dy * dz > 2147483647LL) g.status = 2;
        else {
            g.minb[0] = (int)floorf(mnx * d.c.ivx); g.minb[1] = (int)floorf(mny * d.c.ivy); g.minb[2] = (int)floorf(mnz * d.c.ivz);
            const int m0 = (int)floorf(mxx * d.c.ivx), m1 = (int)floorf(mxy * d.c.ivy);
            const int d0 = m0 - g.minb[0] + 1, d1 = m1 - g.minb[1] + 1;
            g.mul[0] = 1; g.mul[1] = d0; g.mul[2] = d0 * d1;
        }
    }
    d.grid[f] = g;
    d.counts[f * SP_NCOUNT + 0] = g.status;
}

/* voxel key per point (voxel_grid_indices.hpp:378-394); non-cropped points get the sentinel key */
__global__ void __launch_bounds__(256) k_voxel_key(SpDev d)
{
    const int f = blockIdx.y, N = d.N;
    const SpFrameGrid g = d.grid[f];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const size_t gi = (size_t)f * N + i;
        const float px = d.x[gi], py = d.y[gi], pz = d.z[gi];
        unsigned key = 0xFFFFFFFFu;
        if (g.status == 0 && sp_finite3(px, py, pz) && !(px < d.c.x0 || px > d.c.x1) && !(py < d.c.y0 || py > d.c.y1) &&
            !(pz < d.c.z0 || pz > d.c.z1)) {
            const int a = (int)(floorf(px * d.c.ivx) - (float)g.minb[0]);
            const int b = (int)(floorf(py * d.c.ivy) - (float)g.minb[1]);
            const int c = (int)(floorf(pz * d.c.ivz) - (float)g.minb[2]);
            key = (unsigned)(a * g.mul[0] + b * g.mul[1] + c * g.mul[2]);
        }
        d.key_a[gi] = ((unsigned long long)(unsigned)f << 32) | key;
        d.val_a[gi] = (unsigned)i;
    }
}

/* After the stable radix sort: one thread per sorted position.  The thread at the head of a run of equal keys
 * walks the run: sequential float centroid (pcl::CentroidPoint), first strictly-nearest point
 * (voxel_grid_indices.hpp:489-513), then the three point filters a6-a8 (StairPerception.cpp:598-713) on the
 * representative.  Per-1024-block class counts feed the list scatter. */
__global__ void __launch_bounds__(256) k_voxel_pick(SpDev d)
{
    __shared__ int cnt[5];
    const int f = blockIdx.y, N = d.N;
    const size_t base = (size_t)f * N;
    if (threadIdx.x < 5) cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long *keys = d.key_b + base;
    const unsigned *vals = d.val_b + base;
    const float *X = d.x + base, *Y = d.y + base, *Z = d.z + base;
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    for (int q = 0; q < 4; ++q) {
        const int j = blockIdx.x * 1024 + q * 256 + threadIdx.x;
        if (j >= N) break;
        const unsigned long long key = keys[j];
        unsigned char code = 0; int rep = -1;
        if ((unsigned)key != 0xFFFFFFFFu && (j == 0 || keys[j - 1] != key)) {
            float sx = 0.f, sy = 0.f, sz = 0.f; int n = 0;
            for (int t = j; t < N && keys[t] == key; ++t) { const unsigned p = vals[t]; sx += X[p]; sy += Y[p]; sz += Z[p]; ++n; }
            const float fn = (float)n;
            const float cx = sx / fn, cy = sy / fn, cz = sz / fn;
            float best = FLT_MAX;
            for (int t = j; t < j + n; ++t) {
                const unsigned p = vals[t];
                const float dx = cx - X[p], dy = cy - Y[p], dz = cz - Z[p];
                const float dis = dx * dx + dy * dy + dz * dz;
                if (best > dis) { best = dis; rep = (int)p; }
            }
            code = SPC_HEAD; ++c0;
            const float nx = d.nx[base + rep], ny = d.ny[base + rep], nz = d.nz[base + rep];
            if (sp_finite3(nx, ny, nz)) {
                code |= SPC_NONAN; ++c1;
                const float px = X[rep], py = Y[rep], pz = Z[rep];
                if (px * px + py * py + pz * pz > d.c.noleg_th2) {
                    code |= SPC_NOLEG; ++c2;
                    if (nx >= d.c.h_ge || nx <= d.c.h_le) { code |= SPC_H; ++c3; }
                    else if (nx >= d.c.v_lo && nx <= d.c.v_hi) { code |= SPC_V; ++c4; }
                }
            }
        }
        d.rep[base + j] = rep; d.code[base + j] = code;
    }
    c0 = __reduce_add_sync(0xffffffffu, c0); c1 = __reduce_add_sync(0xffffffffu, c1); c2 = __reduce_add_sync(0xffffffffu, c2);
    c3 = __reduce_add_sync(0xffffffffu, c3); c4 = __reduce_add_sync(0xffffffffu, c4);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&cnt[0], c0); atomicAdd(&cnt[1], c1); atomicAdd(&cnt[2], c2); atomicAdd(&cnt[3], c3); atomicAdd(&cnt[4], c4); }
    __syncthreads();
    if (threadIdx.x < 5) d.blk[((size_t)f * d.nblk + blockIdx.x) * 5 + threadIdx.x] = cnt[threadIdx.x];
}

/* exclusive scan of the per-block counts of one frame (one CTA per frame) + frame totals */
__global__ void __launch_bounds__(256) k_scan_blocks(SpDev d)
{
    __shared__ int carry[5];
    __shared__ int ws[5][8];
    const int f = blockIdx.x, nblk = d.nblk;
    int *blk = d.blk + (size_t)f * nblk * 5;
    if (threadIdx.x < 5) carry[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int b0 = 0; b0 < nblk; b0 += 256) {
        const int b = b0 + threadIdx.x;
        int v[5], inc[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) { v[k] = b < nblk ? blk[b * 5 + k] : 0; inc[k] = v[k]; }
#pragma unroll
        for (int k = 0; k < 5; ++k)
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc[k], o); if (lane >= o) inc[k] += t; }
        if (lane == 31) { for (int k = 0; k < 5; ++k) ws[k][wid] = inc[k]; }
        __syncthreads();
        int add[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) { int a = carry[k]; for (int w = 0; w < wid; ++w) a += ws[k][w]; add[k] = a; }
        if (b < nblk) { for (int k = 0; k < 5; ++k) blk[b * 5 + k] = add[k] + inc[k] - v[k]; }
        __syncthreads();
        if (threadIdx.x == 255) { for (int k = 0; k < 5; ++k) carry[k] = add[k] + inc[k]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int *c = d.counts + f * SP_NCOUNT;
        c[3] = carry[0]; c[4] = carry[1]; c[5] = carry[2]; c[6] = carry[3]; c[7] = carry[4];
        d.seg_n[2 * f] = carry[3]; d.seg_n[2 * f + 1] = carry[4];
    }
}

/* ordered scatter: voxel representative list (+ class code) and the horizontal / vertical point sets */
__global__ void __launch_bounds__(256) k_scatter_lists(SpDev d)
{
    __shared__ int wsum[33];
    const int f = blockIdx.y, N = d.N;
    const size_t base = (size_t)f * N;
    const int *blk = d.blk + ((size_t)f * d.nblk + blockIdx.x) * 5;
    int b_head = blk[0], b_h = blk[3], b_v = blk[4];
    for (int q = 0; q < 4; ++q) {
        const int j = blockIdx.x * 1024 + q * 256 + threadIdx.x;
        const unsigned char code = j < N ? d.code[base + j] : 0;
        const int rep = j < N ? d.rep[base + j] : -1;
        int tot;
        const int ph = sp_block_scan_flag((code & SPC_HEAD) != 0, wsum, &tot);
        if (code & SPC_HEAD) { d.voxel_idx[base + b_head + ph] = rep; d.voxel_code[base + b_head + ph] = code; }
        b_head += tot;
        const int pH = sp_block_scan_flag((code & SPC_H) != 0, wsum, &tot);
        if (code & SPC_H) d.pts[(size_t)(2 * f) * N + b_h + pH] = make_float4(d.x[base + rep], d.y[base + rep], d.z[base + rep], __int_as_float(rep));
        b_h += tot;
        const int pV = sp_block_scan_flag((code & SPC_V) != 0, wsum, &tot);
        if (code & SPC_V) d.pts[(size_t)(2 * f + 1) * N + b_v + pV] = make_float4(d.x[base + rep], d.y[base + rep], d.z[base + rep], __int_as_float(rep));
        b_v += tot;
    }
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
