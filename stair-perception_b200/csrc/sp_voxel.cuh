/*
 * sp_voxel.cuh -- the voxel stage of one frame in ONE kernel, one CTA per frame, no host round trip:
 *
 *   voxelGridCloudANDNormal (StairPerception.cpp:489-517) -> pcl::VoxelGridIndices::applyFilterIndices
 *   (/root/reference/modules/stairperception/voxel_grid_indices.hpp:277-517) followed by removeNANNormalPoints (:598-616),
 *   removeClosetPoints (:619-639) and extractPerpendicularPlanePoints (:670-713).
 *
 *   1. grid of the frame from the crop box K1K2 left in d.mm (voxel_grid_indices.hpp:292-318) and the "< 100 points" guard
 *   2. voxel index per cropped point (:378-394), written compacted in pixel order (ordered block compaction), together with the
 *      digit histograms of every sort pass
 *   3. stable LSD radix sort of (voxel index, pixel) on 9-bit digits: ceil(bits of the grid / 9) passes (3 for the usual 2^26-cell
 *      grids), each a stable multi-split of 4096-element tiles: warp-striped keys, __match_any_sync ranks, per-warp digit counters
 *      in shared memory, one prefix over the warps per digit.  Stable = ascending pixel index inside a voxel: the oracle's
 *      documented tie rule for the reference's std::sort (:399)
 *   4. run heads of the sorted keys = voxels (:411-422).  Heads are compacted per tile so that every lane walks a run: sequential
 *      float centroid, first strictly-nearest point (:489-513), the three point filters on the representative.  The voxel list
 *      and the horizontal / vertical point sets come out in voxel order by block scans: no separate scatter pass
 *
 * Replaces k_frame_grid, k_voxel_offsets, k_frame_offsets, the host read of the offsets, k_voxel_key, cub::DeviceRadixSort,
 * k_voxel_pick, k_scan_blocks and k_scatter_lists of round 1.
 */
#pragma once
#include "sp_common.cuh"
#include "sp_async.cuh"

#define VX_T 512
#define VX_NW (VX_T / 32)
#define VX_E 8                           /* keys per thread and tile */
#define VX_TILE (VX_T * VX_E)
#define VX_RB 9
#define VX_BINS (1 << VX_RB)
#define VX_PE 4                          /* sorted positions per thread and tile of the run-head pass */
#define VX_OV 64                         /* positions staged past the end of a tile: a run that starts in the tile usually ends there */
static_assert(VX_T == VX_BINS, "one digit per thread in the bin scans");

#define VX_KT (VX_T * 4)                  /* pixels per tile of the key pass */
struct __align__(128) VxSmem {
    /* staging, by phase (the phases never overlap): raw tiles arrive here by bulk async copies (TMA, completion on mbar) */
    union {
        float4 kg[2][VX_KT];                                            /* key pass: levelled points of a tile, double-buffered */
        struct { unsigned k[2][VX_TILE], v[2][VX_TILE]; } so;           /* sort pass: (voxel index, pixel) of a tile, double-buffered */
        struct {
            float4 pt[VX_T * VX_PE + VX_OV];                            /* run pass: the tile's points {x, y, z, pixel} in sorted order */
            unsigned pk[VX_T * VX_PE + VX_OV + 1];                      /*           their voxel indices; [0] = the key before the tile */
            int heads[VX_T * VX_PE];                                    /*           run heads of the tile, in order */
        } pi;
    } u;
    unsigned wc[VX_NW][VX_BINS];         /* per-warp digit counters of a tile, then the warps' scatter offsets */
    unsigned hist[4][VX_BINS];           /* digit histograms of the (up to four) passes */
    unsigned binoff[VX_BINS];            /* running offset of every digit of the current pass */
    int wsum[VX_NW + 1];
    int wsum3[33];                       /* sp_block_scan_flags3 keeps its total in [32] */
    SpFrameGrid g;
    int vbits, npass;
    unsigned long long mbar[2];
};

/* contiguous global -> shared bulk copy (TMA, SASS UBLKCP); bytes a multiple of 16, both ends 16-byte aligned */
__device__ __forceinline__ void vx_bulk(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(st_sptr(dst)), "l"(src), "r"(bytes), "r"(st_sptr(bar)) : "memory");
}

__device__ __forceinline__ int vx_block_excl(int v, int *ws, int *total)       /* exclusive block scan of an int, VX_T threads */
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();
    if (lane == 31) ws[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const int x = lane < VX_NW ? ws[lane] : 0;
        int xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
        if (lane < VX_NW) ws[lane] = xi - x;
        if (lane == 31) ws[VX_NW] = xi;
    }
    __syncthreads();
    *total = ws[VX_NW];
    return ws[wid] + inc - v;
}

__global__ void __launch_bounds__(VX_T, 2) k_voxel_frame(SpDev d, int force_bits)
{
    extern __shared__ __align__(128) unsigned char vx_raw[];
    VxSmem &s = *reinterpret_cast<VxSmem *>(vx_raw);
    const int f = blockIdx.x, N = d.N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const size_t base = (size_t)f * N;
    const size_t sbase = (size_t)f * (size_t)((N + 3) & ~3);             /* the frame's slice of the sort arrays (16-byte aligned) */
    int *cnts = d.counts + f * SP_NCOUNT;
    if (tid == 0) {
        st_mbar_init(&s.mbar[0], 1); st_mbar_init(&s.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned ph = 0;                                                         /* bit b = parity the next wait on mbar[b] expects */

    /* ---- 1. the frame's grid (k_frame_grid of round 1) */
    if (tid == 0) {
        SpFrameGrid g;
        g.status = 0; g.cells = 0; g.sorted_in_b = 0;
        for (int k = 0; k < 3; ++k) { g.minb[k] = 0; g.mul[k] = 0; }
        const int ncrop = cnts[2];
        if (ncrop < 100) g.status = 1;                                       /* StairPerception.cpp:56 */
        else {
            const int *mm = d.mm + f * 8;
            const float mnx = sp_ord2f(mm[0]), mny = sp_ord2f(mm[1]), mnz = sp_ord2f(mm[2]);
            const float mxx = sp_ord2f(mm[3]), mxy = sp_ord2f(mm[4]), mxz = sp_ord2f(mm[5]);
            const long long dx = (long long)((mxx - mnx) * d.c.ivx) + 1, dy = (long long)((mxy - mny) * d.c.ivy) + 1,
                            dz = (long long)((mxz - mnz) * d.c.ivz) + 1;
            if (dx * dy * dz > 2147483647LL) g.status = 2;                   /* voxel_grid_indices.hpp:295-303 */
            else {
                g.minb[0] = (int)floorf(mnx * d.c.ivx); g.minb[1] = (int)floorf(mny * d.c.ivy); g.minb[2] = (int)floorf(mnz * d.c.ivz);
                const int m0 = (int)floorf(mxx * d.c.ivx), m1 = (int)floorf(mxy * d.c.ivy);
                const int d0 = m0 - g.minb[0] + 1, d1 = m1 - g.minb[1] + 1;
                g.mul[0] = 1; g.mul[1] = d0; g.mul[2] = d0 * d1;
                /* bound of the voxel indices formed below (floor differences, one more than dx, dy, dz at most); a product past
                 * int range means wrapped indices that need all 32 bits */
                const long long d2 = (long long)((int)floorf(mxz * d.c.ivz) - g.minb[2] + 1);
                const long long cells = (long long)d0 * d1 * d2;
                g.cells = cells > 2147483647LL ? -1 : (int)cells;
            }
        }
        int vbits = 1;
        while (vbits < 32 && (g.cells < 0 || (1LL << vbits) < (long long)g.cells)) ++vbits;
        if (force_bits > vbits) vbits = min(32, force_bits);
        const int npass = (vbits + VX_RB - 1) / VX_RB;
        g.sorted_in_b = npass & 1;
        s.g = g; s.vbits = vbits; s.npass = npass;
        d.grid[f] = g;
        cnts[0] = g.status;
    }
    for (int k = tid; k < 4 * VX_BINS; k += VX_T) (&s.hist[0][0])[k] = 0u;
    __syncthreads();
    const SpFrameGrid g = s.g;
    if (g.status != 0) {                                                     /* process() returns before the voxel filter */
        if (tid == 0) { cnts[3] = 0; cnts[4] = 0; cnts[5] = 0; cnts[6] = 0; cnts[7] = 0; d.seg_n[2 * f] = 0; d.seg_n[2 * f + 1] = 0; }
        return;
    }
    const int npass = s.npass;
    unsigned *ka = reinterpret_cast<unsigned *>(d.key_a) + sbase, *kb = reinterpret_cast<unsigned *>(d.key_b) + sbase;
    unsigned *va = d.val_a + sbase, *vb = d.val_b + sbase;
    const float4 *P = d.xyzw + base;

    /* ---- 2. voxel index per cropped point, compacted in pixel order (warp-striped tiles: position order = (warp, k, lane)).
     * The levelled points of tile i + 1 are in flight (bulk copy) while tile i is processed */
    int M = 0;
    {
        const int nkt = (N + VX_KT - 1) / VX_KT;
        if (tid == 0) {
            const unsigned bytes = (unsigned)min(VX_KT, N) * 16u;
            st_mbar_expect(&s.mbar[0], bytes);
            vx_bulk(&s.u.kg[0][0], P, bytes, &s.mbar[0]);
        }
        for (int it = 0; it < nkt; ++it) {
            const int t0 = it * VX_KT, b = it & 1;
            if (tid == 0 && it + 1 < nkt) {                                  /* buffer b ^ 1 was consumed before the last barrier */
                const unsigned bytes = (unsigned)min(VX_KT, N - (t0 + VX_KT)) * 16u;
                st_mbar_expect(&s.mbar[b ^ 1], bytes);
                vx_bulk(&s.u.kg[b ^ 1][0], P + t0 + VX_KT, bytes, &s.mbar[b ^ 1]);
            }
            st_mbar_wait(&s.mbar[b], (ph >> b) & 1u); ph ^= 1u << b;
            const int cl = wid * (32 * 4);
            unsigned key[4];
            unsigned okm = 0;
            int rank[4], cnt = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int q = cl + k * 32 + lane, i = t0 + q;
                bool ok = false; unsigned ky = 0;
                if (i < N) {
                    const float4 pw = s.u.kg[b][q];
                    const float px = pw.x, py = pw.y, pz = pw.z;
                    if (sp_finite3(px, py, pz) && !(px < d.c.x0 || px > d.c.x1) && !(py < d.c.y0 || py > d.c.y1) && !(pz < d.c.z0 || pz > d.c.z1)) {
                        const int a = (int)(floorf(px * d.c.ivx) - (float)g.minb[0]);
                        const int bq = (int)(floorf(py * d.c.ivy) - (float)g.minb[1]);
                        const int c = (int)(floorf(pz * d.c.ivz) - (float)g.minb[2]);
                        ky = (unsigned)(a * g.mul[0] + bq * g.mul[1] + c * g.mul[2]);
                        ok = true;
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                key[k] = ky; rank[k] = cnt + __popc(bal & lt); cnt += __popc(bal);
                if (ok) okm |= 1u << k;
            }
            int tot;
            const int wbase = vx_block_excl(lane == 0 ? cnt : 0, s.wsum, &tot);   /* lane 0 of each warp carries the warp's count */
            const int wb = __shfl_sync(0xffffffffu, wbase, 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (okm & (1u << k)) {
                    const int pos = M + wb + rank[k];
                    ka[pos] = key[k]; va[pos] = (unsigned)(t0 + cl + k * 32 + lane);
                    for (int p = 0; p < npass; ++p) atomicAdd(&s.hist[p][(key[k] >> (p * VX_RB)) & (VX_BINS - 1)], 1u);
                }
            }
            M += tot;
        }
    }
    asm volatile("fence.proxy.async;" ::: "memory");                        /* the pairs written above are read by bulk copies below */
    __syncthreads();

    /* ---- 3. stable LSD radix sort; tile i + 1 of the source is in flight while tile i is ranked and scattered */
    for (int p = 0; p < npass; ++p) {
        const unsigned *sk = (p & 1) ? kb : ka, *sv = (p & 1) ? vb : va;
        unsigned *dk = (p & 1) ? ka : kb, *dv = (p & 1) ? va : vb;
        const int shift = p * VX_RB;
        const int nt = (M + VX_TILE - 1) / VX_TILE;
        if (tid == 0) {
            const unsigned bytes = ((unsigned)min(VX_TILE, M) * 4u + 15u) & ~15u;
            st_mbar_expect(&s.mbar[0], 2 * bytes);
            vx_bulk(&s.u.so.k[0][0], sk, bytes, &s.mbar[0]);
            vx_bulk(&s.u.so.v[0][0], sv, bytes, &s.mbar[0]);
        }
        {
            int tot;
            const int e = vx_block_excl((int)s.hist[p][tid], s.wsum, &tot);       /* VX_T == VX_BINS: one digit per thread */
            s.binoff[tid] = (unsigned)e;
        }
        for (int k = lane; k < VX_BINS; k += 32) s.wc[wid][k] = 0u;
        __syncthreads();
        for (int it = 0; it < nt; ++it) {
            const int t0 = it * VX_TILE, b = it & 1;
            if (tid == 0 && it + 1 < nt) {
                const unsigned bytes = ((unsigned)min(VX_TILE, M - (t0 + VX_TILE)) * 4u + 15u) & ~15u;
                st_mbar_expect(&s.mbar[b ^ 1], 2 * bytes);
                vx_bulk(&s.u.so.k[b ^ 1][0], sk + t0 + VX_TILE, bytes, &s.mbar[b ^ 1]);
                vx_bulk(&s.u.so.v[b ^ 1][0], sv + t0 + VX_TILE, bytes, &s.mbar[b ^ 1]);
            }
            st_mbar_wait(&s.mbar[b], (ph >> b) & 1u); ph ^= 1u << b;
            const int cl = wid * (32 * VX_E), cb = t0 + cl;
            unsigned key[VX_E], val[VX_E], rank[VX_E];
#pragma unroll
            for (int k = 0; k < VX_E; ++k) { key[k] = s.u.so.k[b][cl + k * 32 + lane]; val[k] = s.u.so.v[b][cl + k * 32 + lane]; }
#pragma unroll
            for (int k = 0; k < VX_E; ++k) {
                const int j = cb + k * 32 + lane;
                const unsigned dg = (key[k] >> shift) & (VX_BINS - 1);
                const unsigned m = __match_any_sync(0xffffffffu, j < M ? dg : (unsigned)VX_BINS + lane);
                unsigned prev = 0;
                if (j < M) prev = s.wc[wid][dg];
                __syncwarp();
                if (j < M && (m & lt) == 0) s.wc[wid][dg] = prev + __popc(m);
                __syncwarp();
                rank[k] = prev + __popc(m & lt);
            }
            __syncthreads();
            {                                                                /* digit tid: where each warp's elements of this tile go */
                unsigned run = s.binoff[tid];
#pragma unroll
                for (int w = 0; w < VX_NW; ++w) { const unsigned c = s.wc[w][tid]; s.wc[w][tid] = run; run += c; }
                s.binoff[tid] = run;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < VX_E; ++k) {
                const int j = cb + k * 32 + lane;
                if (j < M) {
                    const unsigned dg = (key[k] >> shift) & (VX_BINS - 1);
                    const unsigned pos = s.wc[wid][dg] + rank[k];
                    dk[pos] = key[k]; dv[pos] = val[k];
                }
            }
            __syncwarp();
            for (int k = lane; k < VX_BINS; k += 32) s.wc[wid][k] = 0u;      /* the warp's own row, for the next tile */
        }
        asm volatile("fence.proxy.async;" ::: "memory");                    /* the next pass reads what this one wrote by bulk copies */
        __syncthreads();
    }

    /* ---- 4. voxels = runs of equal keys; representative + class per voxel; ordered lists */
    const unsigned *K = (npass & 1) ? kb : ka, *V = (npass & 1) ? vb : va;
    const size_t sH = (size_t)(2 * f) * N, sV = (size_t)(2 * f + 1) * N;
    int nvox = 0, nnonan = 0, nnoleg = 0, nH = 0, nV = 0;                    /* CTA-uniform running totals */
    for (int t0 = 0; t0 < M; t0 += VX_T * VX_PE) {
        /* stage the tile: one parallel gather per sorted position (the run walks below then read shared memory, not a chain of
         * dependent global loads) */
        const int nst = min(M - t0, VX_T * VX_PE + VX_OV);
        {
            unsigned vv[VX_PE + 1], kq[VX_PE + 1];
#pragma unroll
            for (int k = 0; k <= VX_PE; ++k) { const int q = tid + k * VX_T; vv[k] = 0u; kq[k] = 0u; if (q < nst) { vv[k] = V[t0 + q]; kq[k] = K[t0 + q]; } }
#pragma unroll
            for (int k = 0; k <= VX_PE; ++k) {
                const int q = tid + k * VX_T;
                if (q < nst) { float4 pw = P[vv[k]]; pw.w = __uint_as_float(vv[k]); s.u.pi.pt[q] = pw; s.u.pi.pk[q + 1] = kq[k]; }
            }
            if (tid == 0) s.u.pi.pk[0] = t0 > 0 ? K[t0 - 1] : 0u;
        }
        __syncthreads();
        /* heads of this tile, in order */
        int hcnt = 0;
        bool hd[VX_PE];
#pragma unroll
        for (int k = 0; k < VX_PE; ++k) {
            const int q = tid * VX_PE + k, j = t0 + q;
            hd[k] = j < M && (j == 0 || s.u.pi.pk[q] != s.u.pi.pk[q + 1]);
            hcnt += hd[k] ? 1 : 0;
        }
        int nh;
        int hp = vx_block_excl(hcnt, s.wsum, &nh);
#pragma unroll
        for (int k = 0; k < VX_PE; ++k) if (hd[k]) s.u.pi.heads[hp++] = tid * VX_PE + k;
        __syncthreads();
        for (int u0 = 0; u0 < nh; u0 += VX_T) {
            const int u = u0 + tid;
            bool fNN = false, fNL = false, fH = false, fV = false;
            int rep = -1; unsigned char code = 0;
            float4 rp = make_float4(0.f, 0.f, 0.f, 0.f);
            if (u < nh) {
                const int q0 = s.u.pi.heads[u];
                const unsigned key = s.u.pi.pk[q0 + 1];
                float sx = 0.f, sy = 0.f, sz = 0.f; int n = 0;
                int q = q0;
                for (; q < nst && s.u.pi.pk[q + 1] == key; ++q) { const float4 pw = s.u.pi.pt[q]; sx += pw.x; sy += pw.y; sz += pw.z; ++n; }
                if (q == nst) for (int t = t0 + q; t < M && K[t] == key; ++t) { const float4 pw = P[V[t]]; sx += pw.x; sy += pw.y; sz += pw.z; ++n; }   /* a run past the staged part */
                const float fn = (float)n;
                const float cx = sx / fn, cy = sy / fn, cz = sz / fn;
                float best = FLT_MAX;
                for (int e = 0; e < n; ++e) {
                    float4 pw;
                    if (q0 + e < nst) pw = s.u.pi.pt[q0 + e];
                    else { const unsigned p = V[t0 + q0 + e]; pw = P[p]; pw.w = __uint_as_float(p); }
                    const float dx = cx - pw.x, dy = cy - pw.y, dz = cz - pw.z;
                    const float dis = dx * dx + dy * dy + dz * dz;
                    if (best > dis) { best = dis; rp = pw; }
                }
                rep = (int)__float_as_uint(rp.w);
                code = SPC_HEAD;
                const float4 nw = d.nrm[base + rep];
                const float nx = nw.x, ny = nw.y, nz = nw.z;
                if (sp_finite3(nx, ny, nz)) {
                    code |= SPC_NONAN; fNN = true;
                    if (rp.x * rp.x + rp.y * rp.y + rp.z * rp.z > d.c.noleg_th2) {
                        code |= SPC_NOLEG; fNL = true;
                        if (nx >= d.c.h_ge || nx <= d.c.h_le) { code |= SPC_H; fH = true; }
                        else if (nx >= d.c.v_lo && nx <= d.c.v_hi) { code |= SPC_V; fV = true; }
                    }
                }
                d.voxel_idx[base + nvox + u] = rep; d.voxel_code[base + nvox + u] = code;
            }
            int tot3;
            const int pk3 = sp_block_scan_flags3(fH, fV, fNN, s.wsum3, &tot3);  /* 10-bit fields: VX_T = 512 < 1023 */
            const int cNL = __syncthreads_count(fNL);
            if (fH) d.pts[sH + nH + (pk3 & 1023)] = rp;                          /* {x, y, z, pixel} */
            if (fV) d.pts[sV + nV + ((pk3 >> 10) & 1023)] = rp;
            nH += tot3 & 1023; nV += (tot3 >> 10) & 1023; nnonan += tot3 >> 20; nnoleg += cNL;
        }
        nvox += nh;
        __syncthreads();                                                     /* heads[] and the staged tile are rewritten by the next tile */
    }
    if (tid == 0) {
        cnts[3] = nvox; cnts[4] = nnonan; cnts[5] = nnoleg; cnts[6] = nH; cnts[7] = nV;
        d.seg_n[2 * f] = nH; d.seg_n[2 * f + 1] = nV;
    }
}
