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

#define VX_T 512
#define VX_NW (VX_T / 32)
#define VX_E 8                           /* keys per thread and tile */
#define VX_TILE (VX_T * VX_E)
#define VX_RB 9
#define VX_BINS (1 << VX_RB)
#define VX_PE 4                          /* sorted positions per thread and tile of the run-head pass */
static_assert(VX_T == VX_BINS, "one digit per thread in the bin scans");

struct VxSmem {
    unsigned wc[VX_NW][VX_BINS];         /* per-warp digit counters of a tile, then the warps' scatter offsets */
    unsigned hist[4][VX_BINS];           /* digit histograms of the (up to four) passes */
    unsigned binoff[VX_BINS];            /* running offset of every digit of the current pass */
    int heads[VX_T * VX_PE];             /* run heads of a tile, in order */
    int wsum[VX_NW + 1];
    int wsum3[33];                       /* sp_block_scan_flags3 keeps its total in [32] */
    SpFrameGrid g;
    int vbits, npass;
    int carry[6];                        /* voxels, non-NaN, no-leg, H, V so far; [5] = M */
};

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
    extern __shared__ __align__(16) unsigned char vx_raw[];
    VxSmem &s = *reinterpret_cast<VxSmem *>(vx_raw);
    const int f = blockIdx.x, N = d.N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const size_t base = (size_t)f * N;
    int *cnts = d.counts + f * SP_NCOUNT;

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
        for (int k = 0; k < 6; ++k) s.carry[k] = 0;
    }
    for (int k = tid; k < 4 * VX_BINS; k += VX_T) (&s.hist[0][0])[k] = 0u;
    __syncthreads();
    const SpFrameGrid g = s.g;
    if (g.status != 0) {                                                     /* process() returns before the voxel filter */
        if (tid == 0) { cnts[3] = 0; cnts[4] = 0; cnts[5] = 0; cnts[6] = 0; cnts[7] = 0; d.seg_n[2 * f] = 0; d.seg_n[2 * f + 1] = 0; }
        return;
    }
    const int npass = s.npass;
    unsigned *ka = reinterpret_cast<unsigned *>(d.key_a) + base, *kb = reinterpret_cast<unsigned *>(d.key_b) + base;
    unsigned *va = d.val_a + base, *vb = d.val_b + base;
    const float4 *P = d.xyzw + base;

    /* ---- 2. voxel index per cropped point, compacted in pixel order (warp-striped tiles: position order = (warp, k, lane)) */
    int M = 0;
    for (int t0 = 0; t0 < N; t0 += VX_TILE) {
        const int cb = t0 + wid * (32 * VX_E);
        unsigned key[VX_E];
        unsigned okm = 0;
        int rank[VX_E], cnt = 0;
#pragma unroll
        for (int k = 0; k < VX_E; ++k) {
            const int i = cb + k * 32 + lane;
            bool ok = false; unsigned ky = 0;
            if (i < N) {
                const float4 pw = P[i];
                const float px = pw.x, py = pw.y, pz = pw.z;
                if (sp_finite3(px, py, pz) && !(px < d.c.x0 || px > d.c.x1) && !(py < d.c.y0 || py > d.c.y1) && !(pz < d.c.z0 || pz > d.c.z1)) {
                    const int a = (int)(floorf(px * d.c.ivx) - (float)g.minb[0]);
                    const int b = (int)(floorf(py * d.c.ivy) - (float)g.minb[1]);
                    const int c = (int)(floorf(pz * d.c.ivz) - (float)g.minb[2]);
                    ky = (unsigned)(a * g.mul[0] + b * g.mul[1] + c * g.mul[2]);
                    ok = true;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            key[k] = ky; rank[k] = cnt + __popc(bal & lt); cnt += __popc(bal);
            if (ok) okm |= 1u << k;
        }
        int tot;
        const int wbase = vx_block_excl(lane == 0 ? cnt : 0, s.wsum, &tot);       /* lane 0 of each warp carries the warp's count */
        const int wb = __shfl_sync(0xffffffffu, wbase, 0);
#pragma unroll
        for (int k = 0; k < VX_E; ++k) {
            if (okm & (1u << k)) {
                const int pos = M + wb + rank[k];
                ka[pos] = key[k]; va[pos] = (unsigned)(cb + k * 32 + lane);
                for (int p = 0; p < npass; ++p) atomicAdd(&s.hist[p][(key[k] >> (p * VX_RB)) & (VX_BINS - 1)], 1u);
            }
        }
        M += tot;
    }
    __syncthreads();

    /* ---- 3. stable LSD radix sort */
    for (int p = 0; p < npass; ++p) {
        const unsigned *sk = (p & 1) ? kb : ka, *sv = (p & 1) ? vb : va;
        unsigned *dk = (p & 1) ? ka : kb, *dv = (p & 1) ? va : vb;
        const int shift = p * VX_RB;
        {
            int tot;
            const int e = vx_block_excl((int)s.hist[p][tid], s.wsum, &tot);       /* VX_T == VX_BINS: one digit per thread */
            s.binoff[tid] = (unsigned)e;
        }
        for (int k = lane; k < VX_BINS; k += 32) s.wc[wid][k] = 0u;
        __syncthreads();
        for (int t0 = 0; t0 < M; t0 += VX_TILE) {
            const int cb = t0 + wid * (32 * VX_E);
            unsigned key[VX_E], val[VX_E], rank[VX_E];
#pragma unroll
            for (int k = 0; k < VX_E; ++k) {
                const int j = cb + k * 32 + lane;
                key[k] = j < M ? sk[j] : 0u; val[k] = j < M ? sv[j] : 0u;
            }
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
        __syncthreads();
    }

    /* ---- 4. voxels = runs of equal keys; representative + class per voxel; ordered lists */
    const unsigned *K = (npass & 1) ? kb : ka, *V = (npass & 1) ? vb : va;
    const size_t sH = (size_t)(2 * f) * N, sV = (size_t)(2 * f + 1) * N;
    int nvox = 0, nnonan = 0, nnoleg = 0, nH = 0, nV = 0;                    /* CTA-uniform running totals */
    for (int t0 = 0; t0 < M; t0 += VX_T * VX_PE) {
        /* heads of this tile, in order */
        int hcnt = 0;
        bool hd[VX_PE];
#pragma unroll
        for (int k = 0; k < VX_PE; ++k) {
            const int j = t0 + tid * VX_PE + k;
            hd[k] = j < M && (j == 0 || K[j - 1] != K[j]);
            hcnt += hd[k] ? 1 : 0;
        }
        int nh;
        int hp = vx_block_excl(hcnt, s.wsum, &nh);
#pragma unroll
        for (int k = 0; k < VX_PE; ++k) if (hd[k]) s.heads[hp++] = t0 + tid * VX_PE + k;
        __syncthreads();
        for (int u0 = 0; u0 < nh; u0 += VX_T) {
            const int u = u0 + tid;
            bool fNN = false, fNL = false, fH = false, fV = false;
            int rep = -1; unsigned char code = 0;
            float4 rp = make_float4(0.f, 0.f, 0.f, 0.f);
            if (u < nh) {
                const int j = s.heads[u];
                const unsigned key = K[j];
                float sx = 0.f, sy = 0.f, sz = 0.f; int n = 0;
                for (int t = j; t < M && K[t] == key; ++t) { const float4 pw = P[V[t]]; sx += pw.x; sy += pw.y; sz += pw.z; ++n; }
                const float fn = (float)n;
                const float cx = sx / fn, cy = sy / fn, cz = sz / fn;
                float best = FLT_MAX;
                for (int t = j; t < j + n; ++t) {
                    const unsigned p = V[t];
                    const float4 pw = P[p];
                    const float dx = cx - pw.x, dy = cy - pw.y, dz = cz - pw.z;
                    const float dis = dx * dx + dy * dy + dz * dz;
                    if (best > dis) { best = dis; rep = (int)p; rp = pw; }
                }
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
            const int pk = sp_block_scan_flags3(fH, fV, fNN, s.wsum3, &tot3);  /* 10-bit fields: VX_T = 512 < 1023 */
            const int cNL = __syncthreads_count(fNL);
            if (fH) { rp.w = __int_as_float(rep); d.pts[sH + nH + (pk & 1023)] = rp; }
            if (fV) { rp.w = __int_as_float(rep); d.pts[sV + nV + ((pk >> 10) & 1023)] = rp; }
            nH += tot3 & 1023; nV += (tot3 >> 10) & 1023; nnonan += tot3 >> 20; nnoleg += cNL;
        }
        nvox += nh;
        __syncthreads();                                                     /* heads[] is rewritten by the next tile */
    }
    if (tid == 0) {
        cnts[3] = nvox; cnts[4] = nnonan; cnts[5] = nnoleg; cnts[6] = nH; cnts[7] = nV;
        d.seg_n[2 * f] = nH; d.seg_n[2 * f + 1] = nV;
    }
}
