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
    double2 CS[2][3][ST_T];               /* 5-row column sums of the six gradient channels, centre row i - 6; double-buffered by step */
    unsigned hist[2][ST_T];               /* per column: the vertical bit histories the window logic needs (see ST_H_*), by step parity */
    unsigned hedge[2][ST_NW + 1];         /* horizontal pair flag of each warp's last lane (its right neighbour is the next warp's lane 0) */
    int anyrisk;
    unsigned long long mbar[ST_RAW];
};
/* hist word of a column after the step that ingested row i: bits 0..8 depth-change flags of rows i - 2 .. i - 10, bits 9..13 / 14..18
 * finite dx / dy gradient elements of rows i - 4 .. i - 8, bits 19..28 inexact-risk coordinates of rows i .. i - 9 */
#define ST_H_FA 9
#define ST_H_FB 14
#define ST_H_RK 19

/* ordered K x K sums of the pixel in ring column t (k1_normal of the tile kernel, read from the gradient ring): the route of
 * pixels whose neighbourhood holds a coordinate without the exactness guarantee.  slot_r = ring slot of gradient row r. */
__device__ __forceinline__ int st_wrap6(int x) { return x >= ST_GR ? x - ST_GR : x; }
template <int K>
__device__ __forceinline__ void st_ordered_sums(const StSmem &s, int t, int slot_r, double *sg)
{
    const int half = K >> 1;
    double sgx0 = 0.0, sgx1 = 0.0, sgx2 = 0.0, sgy0 = 0.0, sgy1 = 0.0, sgy2 = 0.0;
#pragma unroll
    for (int a = 0; a < K; ++a) {
        double rx0 = 0.0, rx1 = 0.0, rx2 = 0.0, ry0 = 0.0, ry1 = 0.0, ry2 = 0.0;
        const int slot = st_wrap6(st_wrap6(slot_r + ST_GR - half) + a);       /* a < K <= 5: one wrap each */
#pragma unroll
        for (int b = 0; b < K; ++b) {
            const float4 ga = s.GA[slot][t - half + b];
            const float2 gb = s.GB[slot][t - half + b];
            rx0 += (double)ga.x; rx1 += (double)ga.y; rx2 += (double)ga.z;
            ry0 += (double)ga.w; ry1 += (double)gb.x; ry2 += (double)gb.y;
        }
        sgx0 += rx0; sgx1 += rx1; sgx2 += rx2; sgy0 += ry0; sgy1 += ry1; sgy2 += ry2;
    }
    sg[0] = sgx0; sg[1] = sgx1; sg[2] = sgx2; sg[3] = sgy0; sg[4] = sgy1; sg[5] = sgy2;
}

/* sg += sign * gradient element (slot, t) */
__device__ __forceinline__ void st_acc(const StSmem &s, int slot, int t, double sign, double *sg)
{
    const float4 ga = s.GA[slot][t];
    const float2 gb = s.GB[slot][t];
    sg[0] += sign * (double)ga.x; sg[1] += sign * (double)ga.y; sg[2] += sign * (double)ga.z;
    sg[3] += sign * (double)ga.w; sg[4] += sign * (double)gb.x; sg[5] += sign * (double)gb.y;
}
__device__ __forceinline__ void st_acc_cs(const StSmem &s, int buf, int t, double *sg)
{
    const double2 a = s.CS[buf][0][t], b = s.CS[buf][1][t], c = s.CS[buf][2][t];
    sg[0] += a.x; sg[1] += a.y; sg[2] += b.x; sg[3] += b.y; sg[4] += c.x; sg[5] += c.y;
}

__global__ void __launch_bounds__(ST_T, 2) k_ingest_normals_strip(SpDev d, const __grid_constant__ CUtensorMap tmap, int ns, int SW, int nb, int BR)
{
    extern __shared__ __align__(128) unsigned char st_smem_raw[];
    StSmem &s = *reinterpret_cast<StSmem *>(st_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
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
        const bool col_inner = c >= 5 && c < W - 5;                       /* PCL's border of 5: no normal outside */
        const int i0 = r0 - ST_HALO;                                      /* first row ingested */
        const int n_ingest = (r1 - r0) + 2 * ST_HALO;
        const int nsteps = (r1 - r0) + ST_HALO + ST_LAG;
        const long long fbase = (long long)f * N;
        float R[9];
        if (rot) {
#pragma unroll
            for (int k = 0; k < 9; ++k) R[k] = __ldg(d.R + 9 * f + k);
        }
        /* rings of the previous item: the gradient ring must read zero for rows above the band, the histories start empty */
#pragma unroll
        for (int k = 0; k < ST_GR; ++k) { s.GA[k][tid] = make_float4(0.f, 0.f, 0.f, 0.f); s.GB[k][tid] = make_float2(0.f, 0.f); }
        s.hist[0][tid] = 0u; s.hist[1][tid] = 0u;
        if (tid < 2 * (ST_NW + 1)) (&s.hedge[0][0])[tid] = 0u;
        if (tid == 0) s.anyrisk = 0;
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
        unsigned fh = 0, fah = 0, fbh = 0, rkh = 0, vmh = 0;              /* vertical bit histories of this column (bit k = k rows ago) */
        bool hp1 = false, vp1 = false;                                    /* pair flags hp(i - 2), vp(i - 3) of this column */
        int dirty = 0;                                                    /* risk mode: steps until the column sums are clean again */
        bool riskmode = false;
        float mnx = FLT_MAX, mny = FLT_MAX, mnz = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX, mxz = -FLT_MAX;
        int nvalid = 0, ncrop = 0;
        float pwx = 0.f, pwy = 0.f, pwz = 0.f;                            /* the levelled point of the pixel finished next (prefetched) */
        int p6 = 0;                                                       /* n % 6: ring slot of the row ingested in this step */
        int lslot = loads % ST_RAW, lpar = (loads / ST_RAW) & 1;          /* TMA slot / parity of the row ingested in this step */
        long long off_i = fbase + (long long)i0 * W + c;                  /* pixel offset of (row i, this column) */

        for (int n = 0; n < nsteps; ++n) {
            const int i = i0 + n;                                         /* row ingested in this step */
            if (tid == 0 && n + 2 < n_ingest) {                           /* slot of row i - 1 is free again: row i + 2 */
                const int slot = lslot + 2 >= ST_RAW ? lslot + 2 - ST_RAW : lslot + 2;
                st_mbar_expect(&s.mbar[slot], ST_T * 16);
                st_tma_row(&tmap, &s.raw[slot][0], &s.mbar[slot], s0 - ST_HALO, i + 2, f);
                st_tma_row(&tmap, &s.raw[slot][ST_BOX], &s.mbar[slot], s0 - ST_HALO + ST_BOX, i + 2, f);
            }
            riskmode = riskmode || (s.anyrisk != 0);                      /* CTA-uniform: written before the last barrier */
            const int pb = (n + 1) & 1, cb = n & 1;                       /* buffers written in the previous / in this step */

            /* ---------------------------------------------------------------- D: finish row r = i - 7 */
            const int r = i - ST_LAG;
            if (r >= r0 && r < r1 && col_owned) {
                /* smoothing window (PCL's 1.0 / 1.4 chamfer map capped at 5; kcat[a][h] of k_ingest_normals, S3) from the bit
                 * histories of the columns c - 4 .. c + 4: f_h = depth-change flags of rows r - 4 .. r + 4 (bit 4 - a) at column
                 * distance h, either side */
                const unsigned *hs = s.hist[pb] + tid;
                const unsigned h0 = hs[0], l1 = hs[-1], r1_ = hs[1], l2 = hs[-2], r2 = hs[2], l3 = hs[-3], r3 = hs[3], l4 = hs[-4], r4 = hs[4];
                const unsigned c1 = h0 | l1 | r1_, c2 = c1 | l2 | r2, c3 = c2 | l3 | r3, c4 = c3 | l4 | r4;   /* cumulative over |h| */
                const bool z0 = ((c2 & 0x010u) | (c1 & 0x028u) | (h0 & 0x044u)) != 0;
                const bool zz2 = (c2 & 0x07cu) != 0;
                const bool z3 = ((c3 & 0x07cu) | (c2 & 0x082u)) != 0;
                const bool z4 = ((c4 & 0x07cu) | (c3 & 0x082u) | (c2 & 0x101u)) != 0;
                /* one finite element in each gradient window (PCL's finite-element counts): rows r - 2 .. r + 2 = bits 4 .. 0 of the
                 * fa / fb fields; windows: 5 = rows -2..2, cols -2..2; 4 = rows -2..1, cols -2..1; 3 = rows / cols -1..1; 2 = -1..0 */
                const unsigned u2 = l1 | h0, u3 = u2 | r1_, u4 = u3 | l2, u5 = u4 | r2;
                const bool wn5 = (u5 & (0x1fu << ST_H_FA)) && (u5 & (0x1fu << ST_H_FB));
                const bool wn4 = (u4 & (0x1eu << ST_H_FA)) && (u4 & (0x1eu << ST_H_FB));
                const bool wn3 = (u3 & (0x0eu << ST_H_FA)) && (u3 & (0x0eu << ST_H_FB));
                const bool wn2 = (u2 & (0x0cu << ST_H_FA)) && (u2 & (0x0cu << ST_H_FB));
                const bool v = (vmh >> 6) & 1u;
                const int kk = !v ? 0 : (wn5 && !z4) ? 5 : (wn4 && z4 && !z3) ? 4 : (wn3 && z3 && !zz2) ? 3 : (wn2 && zz2 && !z0) ? 2 : 0;
                /* an inexact coordinate within three pixels: rows r - 3 .. r + 3 = bits 3 .. 9 of the risk field */
                const bool wrisk = riskmode && (((c3) & (0x3f8u << ST_H_RK)) != 0);
                const long long gi = off_i - (long long)ST_LAG * W;
                if (kk == 0) d.nrm[gi] = make_float4(qnan, qnan, qnan, 0.f);
                else {
                    double sg[6];
                    const int gslot = st_wrap6(p6 + 5);                   /* ring slot of gradient row r (n - 7 = n + 5 mod 6) */
                    if (wrisk) {
                        if (kk == 5) st_ordered_sums<5>(s, tid, gslot, sg);
                        else if (kk == 4) st_ordered_sums<4>(s, tid, gslot, sg);
                        else if (kk == 3) st_ordered_sums<3>(s, tid, gslot, sg);
                        else st_ordered_sums<2>(s, tid, gslot, sg);
                    } else if (kk == 5) {
                        const double2 a0 = s.CS[pb][0][tid - 2], a1 = s.CS[pb][0][tid - 1], a2 = s.CS[pb][0][tid + 1], a3 = s.CS[pb][0][tid + 2];
                        const double2 b0 = s.CS[pb][1][tid - 2], b1 = s.CS[pb][1][tid - 1], b2 = s.CS[pb][1][tid + 1], b3 = s.CS[pb][1][tid + 2];
                        const double2 e0 = s.CS[pb][2][tid - 2], e1 = s.CS[pb][2][tid - 1], e2 = s.CS[pb][2][tid + 1], e3 = s.CS[pb][2][tid + 2];
                        sg[0] = ((a0.x + a1.x) + w0) + (a2.x + a3.x); sg[1] = ((a0.y + a1.y) + w1) + (a2.y + a3.y);
                        sg[2] = ((b0.x + b1.x) + w2) + (b2.x + b3.x); sg[3] = ((b0.y + b1.y) + w3) + (b2.y + b3.y);
                        sg[4] = ((e0.x + e1.x) + w4) + (e2.x + e3.x); sg[5] = ((e0.y + e1.y) + w5) + (e2.y + e3.y);
                    } else {
                        /* windows of 4, 3, 2 with the exactness guarantee: every partial sum is exact, so the order is free.  The
                         * 5-row column sums minus the rows the window leaves out, or the 2 x 2 elements directly */
                        sg[0] = sg[1] = sg[2] = sg[3] = sg[4] = sg[5] = 0.0;
                        if (kk == 4) {                                    /* rows r - 2 .. r + 1, columns c - 2 .. c + 1 */
                            const int below = st_wrap6(gslot + 2);
                            sg[0] = w0; sg[1] = w1; sg[2] = w2; sg[3] = w3; sg[4] = w4; sg[5] = w5;
                            st_acc_cs(s, pb, tid - 2, sg); st_acc_cs(s, pb, tid - 1, sg); st_acc_cs(s, pb, tid + 1, sg);
                            st_acc(s, below, tid - 2, -1.0, sg); st_acc(s, below, tid - 1, -1.0, sg); st_acc(s, below, tid, -1.0, sg); st_acc(s, below, tid + 1, -1.0, sg);
                        } else if (kk == 3) {                             /* rows r - 1 .. r + 1, columns c - 1 .. c + 1 */
                            const int above = st_wrap6(gslot + ST_GR - 2), below = st_wrap6(gslot + 2);
                            sg[0] = w0; sg[1] = w1; sg[2] = w2; sg[3] = w3; sg[4] = w4; sg[5] = w5;
                            st_acc_cs(s, pb, tid - 1, sg); st_acc_cs(s, pb, tid + 1, sg);
                            for (int b = -1; b <= 1; ++b) { st_acc(s, above, tid + b, -1.0, sg); st_acc(s, below, tid + b, -1.0, sg); }
                        } else {                                          /* rows r - 1 .. r, columns c - 1 .. c */
                            const int up = st_wrap6(gslot + ST_GR - 1);
                            st_acc(s, up, tid - 1, 1.0, sg); st_acc(s, up, tid, 1.0, sg); st_acc(s, gslot, tid - 1, 1.0, sg); st_acc(s, gslot, tid, 1.0, sg);
                        }
                    }
                    st_finish(d, (size_t)gi, pwx, pwy, pwz, sg[0], sg[1], sg[2], sg[3], sg[4], sg[5]);
                }
            }

            /* ---------------------------------------------------------------- B: depth-change pair flags (on the levelled z: the reference's quirk) */
            {
                /* hp(i - 1): this column against its right neighbour; vp(i - 2): row i - 2 against row i - 1.  The flag of pixel
                 * (i - 2, c) is complete now: hp(i - 2) of this column and of the left neighbour, vp(i - 2), vp(i - 3) */
                const float zr = tid + 1 < ST_T ? s.P[st_wrap6(p6 + 5)][tid + 1].z : qnan;
                const bool hpn = n >= 1 && k1_bad_pair(z1, zr);
                const bool vpn = n >= 2 && k1_bad_pair(z2, z1);
                bool hl = __shfl_up_sync(0xffffffffu, hp1, 1);
                if (lane == 0) hl = s.hedge[pb][wid] != 0;                /* the left warp's last lane, one step ago */
                if (lane == 31) s.hedge[cb][wid + 1] = hpn ? 1u : 0u;
                fh = ((fh << 1) | ((hp1 | hl | vpn | vp1) ? 1u : 0u)) & 0x1ffu;
                hp1 = hpn; vp1 = vpn;
            }

            /* ---------------------------------------------------------------- A: ingest row i */
            float z0n = qnan;
            bool risk = false, vmn = false;
            if (n < n_ingest) {
                st_mbar_wait(&s.mbar[lslot], (unsigned)lpar);
                const float4 p = s.raw[lslot][tid];
                float ox = p.x, oy = p.y, oz = p.z;
                if (rot) {
                    if (!isnan(p.x)) {
                        ox = R[0] * p.x + R[1] * p.y + R[2] * p.z;
                        oy = R[3] * p.x + R[4] * p.y + R[5] * p.z;
                        oz = R[6] * p.x + R[7] * p.y + R[8] * p.z;
                    } else { ox = oy = oz = qnan; }
                }
                /* remove_nan_points: r, g or b == 0 marks an invalid pixel; pixels outside the image arrive as NaN records */
                if (__vcmpeq4(__float_as_uint(p.w) | 0xff000000u, 0u)) { ox = oy = oz = qnan; }
                bool crop = false;
                if (col_owned && i >= r0 && i < r1) {
                    d.xyzw[off_i] = make_float4(ox, oy, oz, 0.f);
                    if (sp_finite3(ox, oy, oz)) {
                        ++nvalid;
                        if (!(ox < d.c.x0 || ox > d.c.x1) && !(oy < d.c.y0 || oy > d.c.y1) && !(oz < d.c.z0 || oz > d.c.z1)) {
                            crop = true; ++ncrop;
                            mnx = fminf(mnx, ox); mny = fminf(mny, oy); mnz = fminf(mnz, oz);
                            mxx = fmaxf(mxx, ox); mxy = fmaxf(mxy, oy); mxz = fmaxf(mxz, oz);
                        }
                    }
                }
                s.P[p6][tid] = make_float4(ox, oy, oz, 0.f);
                z0n = oz;
                risk = k1_inexact_risk(ox) | k1_inexact_risk(oy) | k1_inexact_risk(oz);
                vmn = col_inner && i >= 5 && i < H - 5 && sp_finite(oz);
                if (risk) s.anyrisk = 1;
                const unsigned bc = __ballot_sync(0xffffffffu, crop);
                if (lane == 0 && bc) {                                    /* cropped pixels per 1024-pixel block, for the voxel stage */
                    const int p0 = i * W + c, left = 1024 - (p0 & 1023);
                    const unsigned lo = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
                    int *vc = d.vcnt + (size_t)f * d.nblk + (p0 >> 10);
                    if (bc & lo) atomicAdd(vc, __popc(bc & lo));
                    if (bc & ~lo) atomicAdd(vc + 1, __popc(bc & ~lo));
                }
            }
            rkh = ((rkh << 1) | (risk ? 1u : 0u)) & 0x3ffu;
            vmh = (vmh << 1) | (vmn ? 1u : 0u);
            z2 = z1; z1 = z0n;

            /* ---------------------------------------------------------------- C: gradients of row j = i - 4, column sums of centre row i - 6 */
            const int j = i - 4;
            bool fa = false, fb = false;
            if (j >= r0 - 2 && j <= r1 + 1) {                             /* CTA-uniform */
                const int pj = st_wrap6(p6 + 2), gj = pj;                 /* ring slots of row j (n - 4 = n + 2 mod 6) */
                const int tl = tid > 0 ? tid - 1 : 0, tr = tid + 1 < ST_T ? tid + 1 : ST_T - 1;
                const float4 L = s.P[pj][tl], Rr = s.P[pj][tr], U = s.P[st_wrap6(p6 + 1)][tid], Dn = s.P[st_wrap6(p6 + 3)][tid];
                const float ax = Rr.x - L.x, ay = Rr.y - L.y, az = Rr.z - L.z;
                const float bx = Dn.x - U.x, by = Dn.y - U.y, bz = Dn.z - U.z;
                fa = sp_finite(ax + (ay + az)); fb = sp_finite(bx + (by + bz));
                const float g0 = fa ? ax : 0.0f, g1 = fa ? ay : 0.0f, g2 = fa ? az : 0.0f, g3 = fb ? bx : 0.0f, g4 = fb ? by : 0.0f, g5 = fb ? bz : 0.0f;
                const float4 oa = s.GA[st_wrap6(gj + 1)][tid];            /* row j - 5 (zero before the band's first gradient row) */
                const float2 ob = s.GB[st_wrap6(gj + 1)][tid];
                s.GA[gj][tid] = make_float4(g0, g1, g2, g3);
                s.GB[gj][tid] = make_float2(g4, g5);
                bool rebuild = false;
                if (riskmode) {
                    /* an inexact element poisons this column's running sums until it has left the window: five steps later the
                     * sums are rebuilt from the ring (all clean again); pixels that would use a poisoned sum take the ordered sums
                     * (risk within three pixels), so nothing reads it in between.  Risk rows j - 1 .. j + 1 = bits 5 .. 3 of the
                     * risk histories (this step's included), columns c - 1 .. c + 1 */
                    const unsigned mine = rkh & 0x38u;
                    const unsigned lft = __shfl_up_sync(0xffffffffu, mine, 1), rgt = __shfl_down_sync(0xffffffffu, mine, 1);
                    const unsigned nl = (s.hist[pb][tl] >> ST_H_RK) & 0x1cu, nr_ = (s.hist[pb][tr] >> ST_H_RK) & 0x1cu;   /* one step old: bits 4 .. 2 */
                    const bool near = (mine | (lane > 0 ? lft : nl << 1) | (lane < 31 ? rgt : nr_ << 1)) != 0;
                    if (near) dirty = 6;
                    if (dirty > 0) { --dirty; rebuild = dirty == 0; }
                }
                if (!rebuild) {
                    w0 += (double)g0; w1 += (double)g1; w2 += (double)g2; w3 += (double)g3; w4 += (double)g4; w5 += (double)g5;
                    w0 -= (double)oa.x; w1 -= (double)oa.y; w2 -= (double)oa.z; w3 -= (double)oa.w; w4 -= (double)ob.x; w5 -= (double)ob.y;
                } else {
                    w0 = (double)g0; w1 = (double)g1; w2 = (double)g2; w3 = (double)g3; w4 = (double)g4; w5 = (double)g5;
                    for (int a = 1; a < 5; ++a) {                         /* rows j - 1 .. j - 4 (rows above the band read zero) */
                        const float4 qa = s.GA[st_wrap6(gj + ST_GR - a)][tid];
                        const float2 qb = s.GB[st_wrap6(gj + ST_GR - a)][tid];
                        w0 += (double)qa.x; w1 += (double)qa.y; w2 += (double)qa.z; w3 += (double)qa.w; w4 += (double)qb.x; w5 += (double)qb.y;
                    }
                }
                s.CS[cb][0][tid] = make_double2(w0, w1); s.CS[cb][1][tid] = make_double2(w2, w3); s.CS[cb][2][tid] = make_double2(w4, w5);
            }
            fah = ((fah << 1) | (fa ? 1u : 0u)) & 0x1fu;
            fbh = ((fbh << 1) | (fb ? 1u : 0u)) & 0x1fu;
            s.hist[cb][tid] = fh | (fah << ST_H_FA) | (fbh << ST_H_FB) | (rkh << ST_H_RK);
            /* the point of the pixel this thread finishes in the next step (row i - 6): it stored it six steps ago */
            if (col_owned && i - 6 >= r0 && i - 6 < r1) {
                const float4 pw = d.xyzw[off_i - 6LL * W];
                pwx = pw.x; pwy = pw.y; pwz = pw.z;
            }
            off_i += W;
            p6 = p6 + 1 == ST_PR ? 0 : p6 + 1;
            if (n < n_ingest) { if (++lslot == ST_RAW) { lslot = 0; lpar ^= 1; } }
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
