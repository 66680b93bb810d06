/*
 * sp_voxel.cuh -- the voxel stage: voxelGridCloudANDNormal (StairPerception.cpp:489-517) -> pcl::VoxelGridIndices::applyFilterIndices
 * (/root/reference/modules/stairperception/voxel_grid_indices.hpp:277-517), followed by removeNANNormalPoints (:598-616),
 * removeClosetPoints (:619-639) and extractPerpendicularPlanePoints (:670-713).
 *
 * Every frame owns a slice of the (voxel index, pixel) sort arrays and everything the host used to read back (how many
 * cropped points a frame has, how many bits its voxel grid needs, which buffer ends up holding the sorted pairs) stays on the
 * device in SpFrameGrid: the stage has no host round trip and can be captured in a CUDA graph.
 *
 *   k_vx_grid      grid of each frame from the crop box K1K2 left in d.mm (:292-318), the "< 100 points" guard, sort passes
 *   k_vx_offsets   K1K2's cropped pixels per 1024-pixel block -> offsets inside the frame's slice
 *   k_vx_key       voxel index per cropped point (:378-394), compacted in pixel order
 *   k_vx_hist / k_vx_scan / k_vx_scatter, once per pass
 *                  stable LSD radix sort on 9-bit digits, hand-written: tiles of 4096 pairs, (tile, digit) counts -> offsets by one
 *                  CTA per frame, then every tile ranks its keys (warp-striped, __match_any_sync groups, per-warp digit counters
 *                  in shared memory, one prefix over the warps per digit) and scatters.  A frame takes ceil(bits of its grid / 9)
 *                  passes: three for the usual 2^26-cell grids; the launches of a pass a frame does not need return at once.
 *                  Stable = ascending pixel index inside a voxel: the oracle's documented tie rule for the reference's std::sort (:399)
 *   k_vx_pick      runs of equal keys = voxels (:411-422): the head of a run walks it: sequential float centroid, first
 *                  strictly-nearest point (:489-513), the three point filters on the representative
 *   k_scan_blocks, k_scatter_lists   ordered voxel list and horizontal / vertical point sets
 * One CTA per frame for the whole stage (tried in this round) serialises a frame: 2.9 ms per 256 frames and twice
 * the single-frame latency; tiles keep every SM busy at any batch size.
 */
#pragma once
#include "sp_common.cuh"

#define VX_T 512
#define VX_NW (VX_T / 32)
#define VX_E 8                           /* keys per thread of a sort tile */
#define VX_TILE (VX_T * VX_E)
#define VX_RB 9
#define VX_BINS (1 << VX_RB)
#define VX_MAXPASS 4
static_assert(VX_T == VX_BINS, "one digit per thread in the digit scans");

__device__ __forceinline__ size_t vx_slice(const SpDev &d, int f) { return (size_t)f * (size_t)((d.N + 3) & ~3); }

/* voxel grid of each frame: getMinMax3D result -> min_b, div_b, divb_mul (voxel_grid_indices.hpp:292-318), the "<100 points" guard
 * of process() (StairPerception.cpp:56), and what the sort needs to know about the frame */
__global__ void k_vx_grid(SpDev d, int nB, int force_bits)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nB) return;
    int *cnts = d.counts + f * SP_NCOUNT;
    SpFrameGrid g;
    g.status = 0; g.cells = 0; g.sorted_in_b = 0; g.npass = 0; g.M = 0;
    for (int k = 0; k < 3; ++k) { g.minb[k] = 0; g.mul[k] = 0; }
    const int ncrop = cnts[2];
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
            /* bound of the voxel indices k_vx_key forms (floor differences, one more than dx, dy, dz at most); a product past
             * int range means wrapped indices that need all 32 bits */
            const long long d2 = (long long)((int)floorf(mxz * d.c.ivz) - g.minb[2] + 1);
            const long long cells = (long long)d0 * d1 * d2;
            g.cells = cells > 2147483647LL ? -1 : (int)cells;
        }
    }
    if (g.status == 0) {
        int vbits = 1;
        while (vbits < 32 && (g.cells < 0 || (1LL << vbits) < (long long)g.cells)) ++vbits;
        if (force_bits > vbits) vbits = min(32, force_bits);
        g.npass = (vbits + VX_RB - 1) / VX_RB;
        g.sorted_in_b = g.npass & 1;
        g.M = ncrop;
    }
    d.grid[f] = g;
    cnts[0] = g.status;
}

/* Only the cropped points go through the sort.  K1K2 left the number of cropped pixels per 1024-pixel block in vcnt; this
 * kernel turns them into offsets inside the frame's slice (one CTA per frame) */
__global__ void __launch_bounds__(256) k_vx_offsets(SpDev d)
{
    __shared__ int ws[8];
    __shared__ int carry;
    const int f = blockIdx.x, nblk = d.nblk;
    const bool live = d.grid[f].status == 0;
    const int *cnt = d.vcnt + (size_t)f * nblk;
    int *off = d.voff + (size_t)f * nblk;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int b0 = 0; b0 < nblk; b0 += 256) {
        const int b = b0 + threadIdx.x;
        const int v = (live && b < nblk) ? cnt[b] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) ws[wid] = inc;
        __syncthreads();
        int add = carry;
        for (int w = 0; w < wid; ++w) add += ws[w];
        if (b < nblk) off[b] = add + inc - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = add + inc;
        __syncthreads();
    }
}


/* voxel index per cropped point (voxel_grid_indices.hpp:378-394), written compacted in pixel order: one CTA per 1024 pixels */
__global__ void __launch_bounds__(256) k_vx_key(SpDev d)
{
    __shared__ int wsum[33];
    const int f = blockIdx.y, N = d.N;
    const SpFrameGrid g = d.grid[f];
    if (g.status != 0) return;
    unsigned *ka = reinterpret_cast<unsigned *>(d.key_a) + vx_slice(d, f);
    unsigned *va = d.val_a + vx_slice(d, f);
    int pos = d.voff[(size_t)f * d.nblk + blockIdx.x];
    for (int q = 0; q < 4; ++q) {
        const int i = blockIdx.x * 1024 + q * 256 + threadIdx.x;
        bool ok = false; unsigned key = 0;
        if (i < N) {
            const float4 pw = d.xyzw[(size_t)f * N + i];
            const float px = pw.x, py = pw.y, pz = pw.z;
            if (sp_finite3(px, py, pz) && !(px < d.c.x0 || px > d.c.x1) && !(py < d.c.y0 || py > d.c.y1) && !(pz < d.c.z0 || pz > d.c.z1)) {
                const int a = (int)(floorf(px * d.c.ivx) - (float)g.minb[0]);
                const int b = (int)(floorf(py * d.c.ivy) - (float)g.minb[1]);
                const int c = (int)(floorf(pz * d.c.ivz) - (float)g.minb[2]);
                key = (unsigned)(a * g.mul[0] + b * g.mul[1] + c * g.mul[2]);
                ok = true;
            }
        }
        int tot;
        const int r = sp_block_scan_flag(ok, wsum, &tot);
        if (ok) { ka[pos + r] = key; va[pos + r] = (unsigned)i; }
        pos += tot;
    }
}

/* ---- the sort: pass p reads buffer a (p even) or b (p odd) of the frame's slice and writes the other one ---------------------- */
/* digit histogram of one tile */
__global__ void __launch_bounds__(VX_T) k_vx_hist(SpDev d, int p)
{
    __shared__ unsigned h[VX_BINS];
    const int f = blockIdx.y, tile = blockIdx.x;
    const SpFrameGrid g = d.grid[f];
    if (p >= g.npass || tile * VX_TILE >= g.M) return;
    const unsigned *sk = reinterpret_cast<const unsigned *>((p & 1) ? d.key_b : d.key_a) + vx_slice(d, f);
    const int shift = p * VX_RB, j1 = min(g.M, (tile + 1) * VX_TILE);
    h[threadIdx.x] = 0u;
    __syncthreads();
    for (int j = tile * VX_TILE + threadIdx.x; j < j1; j += VX_T) atomicAdd(&h[(sk[j] >> shift) & (VX_BINS - 1)], 1u);
    __syncthreads();
    d.vth[((size_t)f * d.ntile + tile) * VX_BINS + threadIdx.x] = h[threadIdx.x];
}

/* (tile, digit) counts of a frame -> first destination of every tile's keys of every digit: digit-major, tile-minor (one CTA per frame) */
__global__ void __launch_bounds__(VX_T) k_vx_scan(SpDev d, int p)
{
    __shared__ int ws[VX_NW + 1];
    const int f = blockIdx.x;
    const SpFrameGrid g = d.grid[f];
    if (p >= g.npass) return;
    const int nt = (g.M + VX_TILE - 1) / VX_TILE, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned *th = d.vth + (size_t)f * d.ntile * VX_BINS;
    unsigned tot = 0;
    for (int t = 0; t < nt; ++t) tot += th[(size_t)t * VX_BINS + tid];
    int inc = (int)tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) ws[wid] = inc;
    __syncthreads();
    int add = 0;
    for (int w = 0; w < wid; ++w) add += ws[w];
    unsigned run = (unsigned)(add + inc) - tot;
    for (int t = 0; t < nt; ++t) { const unsigned c = th[(size_t)t * VX_BINS + tid]; th[(size_t)t * VX_BINS + tid] = run; run += c; }
}

/* one tile: stable rank of every key among the tile's keys of the same digit, then the scatter */
__global__ void __launch_bounds__(VX_T, 3) k_vx_scatter(SpDev d, int p)
{
    __shared__ __align__(16) unsigned wc[VX_NW][VX_BINS];                    /* per-warp digit counters, then the staged pairs */
    __shared__ unsigned off[VX_BINS], lst[VX_BINS], wsum[VX_NW];
    const int f = blockIdx.y, tile = blockIdx.x;
    const SpFrameGrid g = d.grid[f];
    if (p >= g.npass || tile * VX_TILE >= g.M) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, M = g.M;
    const unsigned lt = (1u << lane) - 1u;
    const size_t sl = vx_slice(d, f);
    const unsigned *sk = reinterpret_cast<const unsigned *>((p & 1) ? d.key_b : d.key_a) + sl, *sv = ((p & 1) ? d.val_b : d.val_a) + sl;
    unsigned *dk = reinterpret_cast<unsigned *>((p & 1) ? d.key_a : d.key_b) + sl, *dv = ((p & 1) ? d.val_a : d.val_b) + sl;
    const int shift = p * VX_RB;
    for (int k = lane; k < VX_BINS / 4; k += 32) reinterpret_cast<uint4 *>(wc[wid])[k] = make_uint4(0u, 0u, 0u, 0u);
    const int cb = tile * VX_TILE + wid * (32 * VX_E);                       /* warp-striped: position order = (warp, k, lane) */
    unsigned key[VX_E], val[VX_E], rank[VX_E], grp[VX_E];
#pragma unroll
    for (int k = 0; k < VX_E; ++k) { const int j = cb + k * 32 + lane; key[k] = j < M ? sk[j] : 0u; val[k] = j < M ? sv[j] : 0u; }
#pragma unroll
    for (int k = 0; k < VX_E; ++k) {
        const int j = cb + k * 32 + lane;
        grp[k] = __match_any_sync(0xffffffffu, j < M ? (key[k] >> shift) & (VX_BINS - 1) : (unsigned)VX_BINS + lane);
    }
    __syncwarp();
    /* the first lane of every digit group takes the group's slots from the warp's counter: one shared-memory atomic per group,
     * the eight rows back to back (a warp's atomics on its own counters complete in program order), instead of eight dependent
     * read - modify - write round trips */
#pragma unroll
    for (int k = 0; k < VX_E; ++k) {
        const int j = cb + k * 32 + lane;
        rank[k] = 0u;
        if (j < M && (grp[k] & lt) == 0u) rank[k] = atomicAdd(&wc[wid][(key[k] >> shift) & (VX_BINS - 1)], (unsigned)__popc(grp[k]));
        __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < VX_E; ++k) rank[k] = __shfl_sync(0xffffffffu, rank[k], __ffs(grp[k]) - 1) + __popc(grp[k] & lt);
    __syncthreads();
    {   /* digit tid: the tile's keys of this digit, warp by warp -> position of each (warp, digit) group in the tile sorted by digit */
        unsigned tot = 0u;
#pragma unroll
        for (int w = 0; w < VX_NW; ++w) { const unsigned c = wc[w][tid]; wc[w][tid] = tot; tot += c; }
        unsigned inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        unsigned before = 0u;
        for (int w = 0; w < wid; ++w) before += wsum[w];
        const unsigned ls = before + inc - tot;                              /* tile-local start of the digit */
        off[tid] = d.vth[((size_t)f * d.ntile + tile) * VX_BINS + tid] - ls; /* local position -> destination */
        lst[tid] = ls;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < VX_E; ++k) { const unsigned dg = (key[k] >> shift) & (VX_BINS - 1); rank[k] += wc[wid][dg] + lst[dg]; }   /* tile-local sorted position */
    __syncthreads();
    /* the pairs go through shared memory (over the counters) in digit order, so that neighbouring threads write neighbouring
     * destinations: a digit's keys of a tile are a run of ~8 consecutive words = whole 32-byte sectors, where the direct scatter
     * sent 32 partial sectors per store instruction to L2 */
    uint2 *stage = reinterpret_cast<uint2 *>(&wc[0][0]);
#pragma unroll
    for (int k = 0; k < VX_E; ++k) { const int j = cb + k * 32 + lane; if (j < M) stage[rank[k]] = make_uint2(key[k], val[k]); }
    __syncthreads();
    const int cnt = min(VX_TILE, M - tile * VX_TILE);
#pragma unroll
    for (int k = 0; k < VX_E; ++k) {
        const int i = k * VX_T + tid;
        if (i < cnt) {
            const uint2 e = stage[i];
            const unsigned pos = off[(e.x >> shift) & (VX_BINS - 1)] + (unsigned)i;
            dk[pos] = e.x; dv[pos] = e.y;
        }
    }
}

/* After the stable radix sort: runs of equal keys = voxels (voxel_grid_indices.hpp:411-422).  One CTA per 1024 sorted positions.
 * The points of the tile (+ 64 positions past its end) are gathered ONCE, in parallel, into shared memory; the run heads are
 * compacted so that every lane of a warp walks a run (a thread per position leaves two thirds of the lanes idle: runs are three
 * points long on average): sequential float centroid (pcl::CentroidPoint), first strictly-nearest point (:489-513), then the three
 * point filters a6-a8 (StairPerception.cpp:598-713) on the representative.  Per-1024-block class counts feed the list scatter. */
#define VP_T 256
#define VP_PE 4
#define VP_TILE (VP_T * VP_PE)
#define VP_OV 64
__global__ void __launch_bounds__(VP_T, 6) k_vx_pick(SpDev d)
{
    __shared__ float4 pt[VP_TILE + VP_OV];                               /* {x, y, z, pixel} in sorted order */
    __shared__ unsigned pk[VP_TILE + VP_OV + 1];                         /* voxel indices; [0] = the key before the tile */
    __shared__ unsigned short heads[VP_TILE];
    __shared__ int wsum[33];
    __shared__ int cnt[5];
    const int f = blockIdx.y, N = d.N, tid = threadIdx.x;
    const size_t base = (size_t)f * N;
    const SpFrameGrid g = d.grid[f];
    const int M = g.M, t0 = blockIdx.x * VP_TILE;                        /* M = 0 when process() stopped before the voxel filter */
    if (tid < 5) cnt[tid] = 0;
    if (t0 >= M) {                                                       /* the list scatter still reads this block's counts */
        __syncthreads();
        if (tid < 5) d.blk[((size_t)f * d.nblk + blockIdx.x) * 5 + tid] = 0;
        return;
    }
    const unsigned *K = reinterpret_cast<const unsigned *>(g.sorted_in_b ? d.key_b : d.key_a) + vx_slice(d, f);
    const unsigned *V = (g.sorted_in_b ? d.val_b : d.val_a) + vx_slice(d, f);
    const float4 *P = d.xyzw + base;
    const int nst = min(M - t0, VP_TILE + VP_OV);
    {
        unsigned vv[VP_PE + 1], kq[VP_PE + 1];
#pragma unroll
        for (int k = 0; k <= VP_PE; ++k) { const int q = tid + k * VP_T; vv[k] = 0u; kq[k] = 0u; if (q < nst) { vv[k] = V[t0 + q]; kq[k] = K[t0 + q]; } }
#pragma unroll
        for (int k = 0; k <= VP_PE; ++k) {
            const int q = tid + k * VP_T;
            if (q < nst) { float4 pw = P[vv[k]]; pw.w = __uint_as_float(vv[k]); pt[q] = pw; pk[q + 1] = kq[k]; }
        }
        if (tid == 0) pk[0] = t0 > 0 ? K[t0 - 1] : 0u;
    }
    __syncthreads();
    /* run heads of the tile, in order; a position that is no head gets its (empty) record here */
    int hcnt = 0;
    bool hd[VP_PE];
#pragma unroll
    for (int k = 0; k < VP_PE; ++k) {
        const int q = tid * VP_PE + k, j = t0 + q;
        hd[k] = j < M && (j == 0 || pk[q] != pk[q + 1]);
        hcnt += hd[k] ? 1 : 0;
        if (j < M && !hd[k]) { d.rep[base + j] = -1; d.code[base + j] = 0; }
    }
    int nh, hp;
    {
        const int lane = tid & 31, wid = tid >> 5;
        int inc = hcnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        int add = 0, all = 0;
        for (int w = 0; w < VP_T / 32; ++w) { const int x = wsum[w]; if (w < wid) add += x; all += x; }
        hp = add + inc - hcnt; nh = all;
    }
#pragma unroll
    for (int k = 0; k < VP_PE; ++k) if (hd[k]) heads[hp++] = (unsigned short)(tid * VP_PE + k);
    __syncthreads();
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    for (int u = tid; u < nh; u += VP_T) {
        const int q0 = heads[u];
        const unsigned key = pk[q0 + 1];
        float sx = 0.f, sy = 0.f, sz = 0.f; int n = 0;
        int q = q0;
        for (; q < nst && pk[q + 1] == key; ++q) { const float4 pw = pt[q]; sx += pw.x; sy += pw.y; sz += pw.z; ++n; }
        if (q == nst) for (int t = t0 + q; t < M && K[t] == key; ++t) { const float4 pw = P[V[t]]; sx += pw.x; sy += pw.y; sz += pw.z; ++n; }   /* a run past the staged part */
        const float fn = (float)n;
        const float cx = sx / fn, cy = sy / fn, cz = sz / fn;
        float best = FLT_MAX;
        float4 rp = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = 0; e < n; ++e) {
            float4 pw;
            if (q0 + e < nst) pw = pt[q0 + e];
            else { const unsigned pp = V[t0 + q0 + e]; pw = P[pp]; pw.w = __uint_as_float(pp); }
            const float dx = cx - pw.x, dy = cy - pw.y, dz = cz - pw.z;
            const float dis = dx * dx + dy * dy + dz * dz;
            if (best > dis) { best = dis; rp = pw; }
        }
        const int rep = (int)__float_as_uint(rp.w);
        unsigned char code = SPC_HEAD; ++c0;
        const float4 nw = d.nrm[base + rep];
        const float nx = nw.x, ny = nw.y, nz = nw.z;
        if (sp_finite3(nx, ny, nz)) {
            code |= SPC_NONAN; ++c1;
            if (rp.x * rp.x + rp.y * rp.y + rp.z * rp.z > d.c.noleg_th2) {
                code |= SPC_NOLEG; ++c2;
                if (nx >= d.c.h_ge || nx <= d.c.h_le) { code |= SPC_H; ++c3; }
                else if (nx >= d.c.v_lo && nx <= d.c.v_hi) { code |= SPC_V; ++c4; }
            }
        }
        d.rep[base + t0 + q0] = rep; d.code[base + t0 + q0] = code;
    }
    c0 = __reduce_add_sync(0xffffffffu, c0); c1 = __reduce_add_sync(0xffffffffu, c1); c2 = __reduce_add_sync(0xffffffffu, c2);
    c3 = __reduce_add_sync(0xffffffffu, c3); c4 = __reduce_add_sync(0xffffffffu, c4);
    if ((tid & 31) == 0) { atomicAdd(&cnt[0], c0); atomicAdd(&cnt[1], c1); atomicAdd(&cnt[2], c2); atomicAdd(&cnt[3], c3); atomicAdd(&cnt[4], c4); }
    __syncthreads();
    if (tid < 5) d.blk[((size_t)f * d.nblk + blockIdx.x) * 5 + tid] = cnt[tid];
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
    const int M = d.grid[f].M;                                            /* sorted positions of this frame */
    if (blockIdx.x * 1024 >= M) return;
    const int *blk = d.blk + ((size_t)f * d.nblk + blockIdx.x) * 5;
    int b_head = blk[0], b_h = blk[3], b_v = blk[4];
    const int nH = d.seg_n[2 * f];                                        /* k_scan_blocks has the frame totals */
    for (int q = 0; q < 4; ++q) {
        const int j = blockIdx.x * 1024 + q * 256 + threadIdx.x;
        const unsigned char code = j < M ? d.code[base + j] : 0;
        const int rep = j < M ? d.rep[base + j] : -1;
        int tot;
        const int pk = sp_block_scan_flags3((code & SPC_HEAD) != 0, (code & SPC_H) != 0, (code & SPC_V) != 0, wsum, &tot);
        const int ph = pk & 1023, pH = (pk >> 10) & 1023, pV = pk >> 20;
        if (code & SPC_HEAD) { d.voxel_idx[base + b_head + ph] = rep; d.voxel_code[base + b_head + ph] = code; }
        if (code & SPC_H) { float4 pw = d.xyzw[base + rep]; pw.w = __int_as_float(rep); d.pts[base + b_h + pH] = pw; }
        if (code & SPC_V) { float4 pw = d.xyzw[base + rep]; pw.w = __int_as_float(rep); d.pts[base + nH + b_v + pV] = pw; }   /* the vertical set follows the horizontal one */
        b_head += tot & 1023; b_h += (tot >> 10) & 1023; b_v += tot >> 20;
    }
}
