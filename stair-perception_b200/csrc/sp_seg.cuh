/*
 * sp_seg.cuh -- per-segment kernels: radius-graph connected components (pcl::EuclideanClusterExtraction),
 * RANSAC plane extraction (pcl::SACSegmentation, SACMODEL_PLANE / SACMODEL_PARALLEL_PLANE), plane merge,
 * per-plane statistics, plane sort and the per-frame result record.
 * A "segment" is the horizontal (s = 2f) or vertical (s = 2f+1) point set of frame f.
 */
#pragma once
#include "sp_common.cuh"

#define SEG_GRID_X 16

/* ==================================================================================================== */
/* Euclidean clustering, StairPerception.cpp:935-963.  Edge iff the sequential-float squared distance of
 * flann::L2_Simple is < r2 (RadiusResultSet's strict compare).  Components by lock-free union-find where the
 * smaller index always becomes the root, so a component's root is its smallest member position. */
__device__ __forceinline__ unsigned sp_cell_hash(int a, int b, int c)
{
    return ((unsigned)a * 73856093u) ^ ((unsigned)b * 19349663u) ^ ((unsigned)c * 83492791u);
}

/* Cell grid of one segment: cell edge = cluster tolerance (x 1.0001), cells hashed into 1 << SP_HASH_BITS buckets.
 * Three passes lay the points out bucket-major (a counting sort): count, scan, scatter.  Points of different
 * cells that share a bucket are told apart by the distance test itself. */
__device__ __forceinline__ unsigned sp_point_bucket(const float4 &p, float inv_cell, int &a, int &b, int &c)
{
    /* any consistent binning with cells no narrower than the tolerance is valid: the 1e-4 margin of the edge covers the rounding
     * of the product for |coordinate| < 800 cells */
    a = (int)floorf(p.x * inv_cell); b = (int)floorf(p.y * inv_cell); c = (int)floorf(p.z * inv_cell);
    return sp_cell_hash(a, b, c) & ((1u << SP_HASH_BITS) - 1u);
}

__global__ void __launch_bounds__(256) k_cl_count(SpDev d)
{
    const int s = blockIdx.y, n = d.seg_n[s];
    const size_t base = sp_seg_base(d, s);
    int *cnt = d.ccnt + ((size_t)s << SP_HASH_BITS);
    const float cell = d.c.inv_cell;
    const int lane = threadIdx.x & 31;
    for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + lane;
        unsigned h = 0x80000000u + lane;                     /* lanes past the end: a group of their own */
        if (i < n) {
            const float4 p = d.pts[base + i];
            int a, b, c;
            h = sp_point_bucket(p, cell, a, b, c);
            d.parent[base + i] = i;
            d.csize[base + i] = 0;
        }
        const unsigned m = __match_any_sync(0xffffffffu, h);  /* list neighbours often share a cell: one atomic per cell and warp */
        if (i < n && lane == __ffs(m) - 1) atomicAdd(&cnt[h], __popc(m));
    }
}

/* exclusive scan of one segment's bucket counts: one CTA of 1024 threads, a warp per 2048 consecutive buckets, read and
 * written as coalesced int4 rows of 128 buckets (pass 1: warp totals; pass 2, from L2: running scan row by row) */
__global__ void __launch_bounds__(1024) k_cl_scan(SpDev d)
{
    __shared__ int wsum[32];
    const int s = blockIdx.x;
    if (d.seg_n[s] == 0) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int per_warp = (1 << SP_HASH_BITS) / 32, rows = per_warp / 128;
    const int4 *cnt = reinterpret_cast<const int4 *>(d.ccnt + ((size_t)s << SP_HASH_BITS) + (size_t)wid * per_warp) + lane;
    int *start = d.cstart + (size_t)s * ((1 << SP_HASH_BITS) + 1) + (size_t)wid * per_warp;
    int tot = 0;
#pragma unroll 4
    for (int q = 0; q < rows; ++q) { const int4 v = cnt[q * 32]; tot += (v.x + v.y) + (v.z + v.w); }
    tot = __reduce_add_sync(0xffffffffu, tot);
    if (lane == 0) wsum[wid] = tot;
    __syncthreads();
    if (wid == 0) {
        int w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        wsum[lane] = wi - w;
        if (lane == 31) d.cstart[(size_t)s * ((1 << SP_HASH_BITS) + 1) + (1 << SP_HASH_BITS)] = wi;
    }
    __syncthreads();
    int run = wsum[wid];                                                  /* buckets before this row */
    for (int q = 0; q < rows; ++q) {
        const int4 v = cnt[q * 32];
        const int mine = (v.x + v.y) + (v.z + v.w);
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        int r = run + inc - mine;
        int *o = start + q * 128 + lane * 4;                             /* cstart rows start at an odd multiple of 4 bytes: scalar stores */
        o[0] = r; r += v.x; o[1] = r; r += v.y; o[2] = r; r += v.z; o[3] = r;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
}

__global__ void __launch_bounds__(256) k_cl_scatter(SpDev d)
{
    const int s = blockIdx.y, n = d.seg_n[s];
    const size_t base = sp_seg_base(d, s);
    int *cnt = d.ccnt + ((size_t)s << SP_HASH_BITS);
    const int *start = d.cstart + (size_t)s * ((1 << SP_HASH_BITS) + 1);
    const float cell = d.c.inv_cell;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + lane;
        unsigned h = 0x80000000u + lane;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n) {
            p = d.pts[base + i];
            int a, b, c;
            h = sp_point_bucket(p, cell, a, b, c);
        }
        const unsigned m = __match_any_sync(0xffffffffu, h);
        const int leader = __ffs(m) - 1;
        int top = 0;
        if (i < n && lane == leader) top = atomicSub(&cnt[h], __popc(m));   /* drains the counters back to zero */
        top = __shfl_sync(0xffffffffu, top, leader);
        if (i < n) {
            const int pos = start[h] + top - 1 - __popc(m & lt);
            d.spts[base + pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
        }
    }
}

__device__ __forceinline__ int sp_uf_find(int *parent, int i)
{
    int p = parent[i];
    while (p != i) { const int g = parent[p]; if (g != p) parent[i] = g; i = p; p = g; }
    return i;
}

__device__ __forceinline__ void sp_uf_unite(int *parent, int a, int b)
{
    for (;;) {
        a = sp_uf_find(parent, a); b = sp_uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }          /* a > b: hook the larger root under the smaller */
        const int old = atomicCAS(&parent[a], a, b);
        if (old == a) return;
        a = old;
    }
}

/* Runs: the segment's points are in voxel order, so list neighbours are mostly spatial neighbours (a tread is walked forward
 * column by column, a riser downwards).  Consecutive points closer than the tolerance form a run, and a run is a ready-made
 * depth-1 tree: parent = first position of the run (a prefix maximum over the run starts, one CTA per segment, no atomics).
 * Every tree edge is an edge of the radius graph, so the components k_cl_union ends with are unchanged -- but it starts from a
 * few hundred shallow trees per surface instead of one node per point, and its "same parent?" filter is exact inside a run. */
__global__ void __launch_bounds__(1024) k_cl_runs(SpDev d)
{
    __shared__ int wmax[32];
    __shared__ int carry;
    const int s = blockIdx.x, n = d.seg_n[s];
    const size_t base = sp_seg_base(d, s);
    const float4 *pts = d.pts + base;
    int *parent = d.parent + base;
    const float r2 = d.c.r2;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        int head = 0;                                                    /* start of the run if this point starts one, else 0 */
        if (i < n && i > 0) {
            const float4 p = pts[i], q = pts[i - 1];
            const float d0 = p.x - q.x, d1 = p.y - q.y, d2 = p.z - q.z;
            float dist = 0.f; dist += d0 * d0; dist += d1 * d1; dist += d2 * d2;   /* flann::L2_Simple order */
            if (!(dist < r2)) head = i;
        }
        int inc = head;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = max(inc, u); }
        if (lane == 31) wmax[wid] = inc;
        __syncthreads();
        int pre = carry;
        for (int w = 0; w < wid; ++w) pre = max(pre, wmax[w]);
        inc = max(inc, pre);
        if (i < n) parent[i] = inc;
        __syncthreads();
        if (threadIdx.x == 1023) carry = inc;
        __syncthreads();
    }
}

/* One warp per cell.  A warp takes a window of 32 positions of the bucket-major point array and handles the buckets that START in
 * it.  For a cell (the points of a bucket with equal cell coordinates; a bucket shared by two cells is split) lanes 0..13 look up
 * the own bucket and the 13 "forward" neighbour buckets, and the concatenation of those ranges is the CANDIDATE list: 32 candidates
 * at a time sit one per lane -- loaded once per cell, not once per member point -- and the cell's points are broadcast from shared
 * memory one after the other, so every distance test runs with a full warp and every unordered pair of points in adjacent cells is
 * tested once (twice inside the own cell).
 * Edges: lanes keep the set of candidate lanes known to be connected already (start: equal parent, i.e. same tree); a member that
 * reaches several of those sets emits ONE edge per set (none to the set that holds the member itself) and merges them, so a cell
 * hands a spanning forest of its little graph to the union-find instead of every close pair.  Edges go to a per-warp queue and are
 * united 32 at a time (lock-free union-find on list positions, smaller index wins: a component's root is its smallest member
 * whatever the visiting order). */
__global__ void __launch_bounds__(256, 8) k_cl_union(SpDev d)
{
    __shared__ int2 queue[8][64];
    __shared__ float4 memb[8][32];
    __shared__ int rstart[8][16], roff[8][16];
    const int s = blockIdx.y, n = d.seg_n[s];
    const size_t base = sp_seg_base(d, s);
    const int *start = d.cstart + (size_t)s * ((1 << SP_HASH_BITS) + 1);
    int *parent = d.parent + base;
    const float4 *sp = d.spts + base;
    float r2 = d.c.r2;
    asm volatile("" : "+l"(parent), "+l"(sp), "+l"(start), "+f"(r2));    /* keep the segment's bases and r2 in registers: one IMAD.WIDE per address */
    __builtin_assume(__isGlobal(parent)); __builtin_assume(__isGlobal(sp)); __builtin_assume(__isGlobal(start));
    const float cell = d.c.inv_cell;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    int2 *q = queue[wid];
    float4 *mb = memb[wid];
    int *rs_ = rstart[wid], *ro_ = roff[wid];
    const int kk = (lane < 14 ? lane : 0) + 13;                          /* lanes 0..13: cell offset (0,0,0), then the 13 forward ones */
    const int da = kk / 9 - 1, db = (kk / 3) % 3 - 1, dc = kk % 3 - 1;
    int qn = 0;                                                          /* warp-uniform */
    for (int t0 = (blockIdx.x * 8 + wid) * 32; t0 < n; t0 += gridDim.x * 256) {
        const int t = t0 + lane;
        bool first = false; unsigned h = 0u; int hnext = 0;
        if (t < n) {
            const float4 p = sp[t];
            int a, b, c;
            h = sp_point_bucket(p, cell, a, b, c);
            first = start[h] == t; hnext = start[h + 1];
        }
        unsigned firsts = __ballot_sync(0xffffffffu, first);
        while (firsts) {                                                 /* buckets that start in this window */
            const int fl = __ffs(firsts) - 1;
            firsts &= firsts - 1;
            const int j0 = t0 + fl, j1 = __shfl_sync(0xffffffffu, hnext, fl);
            for (int m0 = j0; m0 < j1; m0 += 32) {                       /* the bucket's points, 32 at a time */
                const int mi = m0 + lane;
                const bool ism = mi < j1;
                int am = 0, bm = 0, cm = 0;
                if (ism) { const float4 pm = sp[mi]; sp_point_bucket(pm, cell, am, bm, cm); mb[lane] = pm; }
                unsigned todo = __ballot_sync(0xffffffffu, ism);
                while (todo) {                                           /* cells of the bucket (one, unless two cells share it) */
                    const int gl = __ffs(todo) - 1;
                    const int ga = __shfl_sync(0xffffffffu, am, gl), gb = __shfl_sync(0xffffffffu, bm, gl), gc = __shfl_sync(0xffffffffu, cm, gl);
                    unsigned gm = __ballot_sync(0xffffffffu, ism && am == ga && bm == gb && cm == gc);
                    todo &= ~gm;
                    int rs = 0, rc = 0;
                    if (lane < 14) {
                        const unsigned hk = sp_cell_hash(ga + da, gb + db, gc + dc) & ((1u << SP_HASH_BITS) - 1u);
                        rs = start[hk]; rc = start[hk + 1] - rs;
                    }
                    int inc = rc;
#pragma unroll
                    for (int o = 1; o < 16; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                    if (lane < 16) { rs_[lane] = rs; ro_[lane] = inc - rc; }
                    const int C = __shfl_sync(0xffffffffu, inc, 15);
                    __syncwarp();
                    for (int c0 = 0; c0 < C; c0 += 32) {                 /* candidates, 32 at a time */
                        const int ci = c0 + lane;
                        const bool isc = ci < C;
                        int k = 0;                                       /* the range ci falls in: largest k with ro_[k] <= ci */
                        if (ro_[k + 8] <= ci) k += 8;
                        if (ro_[k + 4] <= ci) k += 4;
                        if (ro_[k + 2] <= ci) k += 2;
                        if (ro_[k + 1] <= ci) k += 1;
                        float4 qq = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);        /* no candidate: NaN fails every test */
                        int oj = -1, pc = -1 - lane;
                        if (isc) { qq = sp[(unsigned)(rs_[k] + (ci - ro_[k]))]; oj = __float_as_int(qq.w); pc = parent[(unsigned)oj]; }
                        unsigned comp = __match_any_sync(0xffffffffu, pc);       /* candidates of one tree */
                        /* range 0 is the own bucket, in member order: member lane ml of this chunk is candidate m0 - j0 + ml */
                        const int self0 = m0 - j0 - c0;
                        for (unsigned g = gm; g; g &= g - 1) {
                            const int ml = __ffs(g) - 1;
                            const float4 pm = mb[ml];
                            const float d0 = pm.x - qq.x, d1 = pm.y - qq.y, d2 = pm.z - qq.z;
                            float dist = d0 * d0; dist += d1 * d1; dist += d2 * d2;   /* flann::L2_Simple order (0 + d0 * d0 == d0 * d0) */
                            const bool hit = dist < r2;
                            const unsigned self = (unsigned)(self0 + ml) < 32u ? 1u << (self0 + ml) : 0u;
                            if (__any_sync(0xffffffffu, hit && (comp & self) == 0u)) {          /* reaches a set it is not known to be in */
                                const int oi = __float_as_int(pm.w);
                                const unsigned H = __ballot_sync(0xffffffffu, hit);
                                const unsigned U = __reduce_or_sync(0xffffffffu, hit ? comp : 0u);
                                const bool lead = hit && lane == __ffs(comp & H) - 1 && (comp & self) == 0u;
                                const unsigned L = __ballot_sync(0xffffffffu, lead);
                                if (lead) q[qn + __popc(L & lt)] = make_int2(oi, oj);
                                qn += __popc(L);
                                if ((U >> lane) & 1u) comp = U;
                                __syncwarp();
                                while (qn >= 32) {
                                    const int2 e = q[qn - 32 + lane];
                                    qn -= 32;
                                    __syncwarp();
                                    sp_uf_unite(parent, e.x, e.y);
                                }
                            }
                        }
                    }
                    __syncwarp();                                        /* ro_ / rs_ are rewritten by the next cell */
                }
                __syncwarp();                                            /* mb is rewritten by the next chunk */
            }
        }
    }
    __syncwarp();
    if (lane < qn) { const int2 e = q[lane]; sp_uf_unite(parent, e.x, e.y); }
}

__global__ void __launch_bounds__(256) k_cl_flatten(SpDev d)
{
    const int s = blockIdx.y, n = d.seg_n[s];
    const size_t base = sp_seg_base(d, s);
    int *parent = d.parent + base;
    const int lane = threadIdx.x & 31;
    for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + lane;
        int r = -1 - lane;                                   /* lanes past the end: a group of their own */
        if (i < n) {
            r = i;
            while (parent[r] != r) r = parent[r];
            d.root[base + i] = r;
        }
        /* list neighbours mostly share their root: one atomic per root and warp instead of one per point (a 5 000-point surface
         * used to send 5 000 atomics to one address) */
        const unsigned m = __match_any_sync(0xffffffffu, r);
        if (i < n && lane == __ffs(m) - 1) atomicAdd(&d.csize[base + r], __popc(m));
    }
}

/* One CTA per segment: count clusters within [min,max]; collect those that can yield a plane (size >= rest),
 * rank them like PCL (size descending; ties: smaller first member), lay out their workspaces and plane slots; then
 * lay the candidate clusters' points out contiguously, cluster by cluster in rank order, members in ascending
 * position order (what pointCluster's per-cluster clouds hold, StairPerception.cpp:951-962) -- a stable multi-way
 * partition: per-warp contiguous chunks, per-(warp, cluster) counts via __match_any_sync, one prefix pass. */
#define RK_THREADS 1024
#define RK_WARPS (RK_THREADS / 32)
__global__ void __launch_bounds__(RK_THREADS) k_cl_rank(SpDev d)
{
    __shared__ int c_root[1024], c_size[1024];
    __shared__ int cnt[RK_WARPS][SP_CMAX + 1];
    __shared__ int nc, nvalid;
    const int s = blockIdx.x, n = d.seg_n[s];
    const size_t base = sp_seg_base(d, s);
    int *crank = d.parent + base;                       /* union-find storage is free again: cluster rank per root position */
    if (threadIdx.x == 0) { nc = 0; nvalid = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (d.root[base + i] != i) continue;
        crank[i] = -1;
        const int sz = d.csize[base + i];
        if (sz < d.c.min_c || sz > d.c.max_c) continue;
        atomicAdd(&nvalid, 1);
        if (sz >= d.c.rest) { const int slot = atomicAdd(&nc, 1); if (slot < 1024) { c_root[slot] = i; c_size[slot] = sz; } }
    }
    __syncthreads();
    const int C = min(nc, 1024);
    if (threadIdx.x == 0) {
        d.n_clusters[s] = nvalid;
        d.n_cand[s] = min(C, SP_CMAX);
        if (nc > SP_CMAX) atomicOr(&d.flags[s >> 1], 1);
    }
    for (int a = threadIdx.x; a < C; a += blockDim.x) {
        int rank = 0, off = 0, slot = 0;
        for (int b = 0; b < C; ++b) {
            const bool before = (c_size[b] > c_size[a]) || (c_size[b] == c_size[a] && c_root[b] < c_root[a]);
            if (before) { ++rank; off += c_size[b]; slot += c_size[b] / d.c.rest; }
        }
        if (rank < SP_CMAX) {
            d.cand_root[s * SP_CMAX + rank] = c_root[a]; d.cand_size[s * SP_CMAX + rank] = c_size[a];
            d.cand_off[s * SP_CMAX + rank] = off; d.cand_slot[s * SP_CMAX + rank] = slot;
            crank[c_root[a]] = rank;
        }
    }
    for (int i = threadIdx.x; i < RK_WARPS * (SP_CMAX + 1); i += blockDim.x) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const int CC = min(C, SP_CMAX);
    if (CC == 0) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int L = ((n + RK_THREADS - 1) / RK_THREADS) * 32;              /* chunk per warp, a multiple of 32 */
    const unsigned lt = (1u << lane) - 1u;
    for (int g = 0; g < L; g += 32) {
        const int i = w * L + g + lane;
        const int key = i < n ? crank[d.root[base + i]] : -1;
        const unsigned m = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && (m & lt) == 0) cnt[w][key] += __popc(m);
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < CC) {
        int run = d.cand_off[s * SP_CMAX + threadIdx.x];
        for (int q = 0; q < RK_WARPS; ++q) { const int t = cnt[q][threadIdx.x]; cnt[q][threadIdx.x] = run; run += t; }
    }
    __syncthreads();
    for (int g = 0; g < L; g += 32) {
        const int i = w * L + g + lane;
        const int key = i < n ? crank[d.root[base + i]] : -1;
        const unsigned m = __match_any_sync(0xffffffffu, key);
        if (key >= 0) {
            const int pos = cnt[w][key] + __popc(m & lt);
            const float4 q = d.pts[base + i];
            d.wx[base + pos] = q.x; d.wy[base + pos] = q.y; d.wz[base + pos] = q.z; d.wpix[base + pos] = __float_as_int(q.w);
        }
        __syncwarp();
        if (key >= 0 && (m & lt) == 0) cnt[w][key] += __popc(m);
        __syncwarp();
    }
}

/* ==================================================================================================== */
/* RANSAC plane extraction: one CTA per (segment, cluster).  Restates pcl::SACSegmentation::segment with
 * RandomSampleConsensus (ransac.hpp), the index-shuffle sampler of SampleConsensusModel (sac_model.h) driven by
 * boost::mt19937(12345), SampleConsensusModelPlane / ParallelPlane and the callers' extract-until-small loops
 * (StairPerception.cpp:773-808, :853-884).
 *   - the cluster's points arrive contiguous (k_cl_rank) in wx/wy/wz/wpix; every extraction writes the remainder
 *     to the twin buffers (ping-pong), so no pass is ordered by a barrier chain;
 *   - hypotheses are drawn by one thread, scored SP_ROUND at a time by the whole CTA (register counters ->
 *     __reduce_add_sync -> shared atomics), then the sequential "first strictly better wins / adaptive k" rule is
 *     replayed exactly;
 *   - the refit's nine single-pass float sums (computeMeanAndCovarianceMatrix) must be added in index order: warp 0
 *     is the "summer" (lane q owns accumulator q, one FADD per element on the critical path), the other warps stage
 *     the nine products per point, double-buffered, with +0.0f for non-inliers (adding +0 never changes a float sum
 *     that started at +0). */
#define RS_THREADS 256
#define RS_TILE (RS_THREADS - 32)
#define RS_NW (RS_THREADS / 32)
struct RsSmem {
    int map_key[1024], map_val[1024];
    float4 model[SP_ROUND];
    int tri[SP_ROUND][3];
    int mvalid[SP_ROUND];
    int counts[SP_ROUND];
    int n_round, done, have, mt_pos, map_n, overflow;
    float best_m[4], refined[4];
    int best_tri[3], best_count, scored;
    float accu[9];
    int n_inl, slot;
    int wcnt[RS_NW], wbase[RS_NW];
    float4 stage[2][RS_TILE][3];
};

__device__ __forceinline__ int rs_map_get(RsSmem &S, int k)
{
    unsigned h = ((unsigned)k * 2654435761u) >> 22;
    for (;;) { const int kk = S.map_key[h]; if (kk == k) return S.map_val[h]; if (kk < 0) return k; h = (h + 1) & 1023u; }
}
__device__ __forceinline__ void rs_map_set(RsSmem &S, int k, int v)
{
    unsigned h = ((unsigned)k * 2654435761u) >> 22;
    for (;;) {
        const int kk = S.map_key[h];
        if (kk == k) { S.map_val[h] = v; return; }
        if (kk < 0) { if (S.map_n >= 960) { S.overflow = 1; return; } S.map_key[h] = k; S.map_val[h] = v; ++S.map_n; return; }
        h = (h + 1) & 1023u;
    }
}

__device__ __forceinline__ bool rs_degenerate(const float *X, const float *Y, const float *Z, int i0, int i1, int i2)
{
    const float q0 = (X[i1] - X[i0]) / (X[i2] - X[i0]);
    const float q1 = (Y[i1] - Y[i0]) / (Y[i2] - Y[i0]);
    const float q2 = (Z[i1] - Z[i0]) / (Z[i2] - Z[i0]);
    return (q0 == q1) && (q2 == q1);
}

__device__ __forceinline__ bool rs_model_valid(const SpConst &c, bool parallel, const float *m)
{
    if (!parallel) return true;
    const float nrm = sqrtf((m[0] * m[0] + m[1] * m[1]) + (m[2] * m[2] + 0.0f));
    const float c0 = m[0] / nrm;
    return !((double)fabsf(c0) > c.sin_angle);
}

__global__ void __launch_bounds__(RS_THREADS, 4) k_ransac(SpDev d)
{
    __shared__ RsSmem S;
    const int s = blockIdx.y, ci = blockIdx.x;
    if (ci >= d.n_cand[s]) return;
    const bool vertical = (s & 1) != 0;
    const size_t base = sp_seg_base(d, s);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int off = d.cand_off[s * SP_CMAX + ci];
    /* current / other point buffers of this cluster (ping-pong across extractions) */
    float *X = d.wx + base + off, *Y = d.wy + base + off, *Z = d.wz + base + off;
    int *PIX = d.wpix + base + off;
    float *X2 = d.vx + base + off, *Y2 = d.vy + base + off, *Z2 = d.vz + base + off;
    int *PIX2 = d.vpix + base + off;
    int *OUT = d.plane_pts + base + off;
    float *OX = d.plane_x + base + off, *OY = d.plane_y + base + off, *OZ = d.plane_z + base + off;
    int n = d.cand_size[s * SP_CMAX + ci], out_n = 0, call = 0;
    const int rest = d.c.rest;
    const float thr = d.c.seg_thr;

    for (;;) {
        if (!vertical && n < rest) break;
        /* ---------------- segment() ---------------- */
        if (tid == 0) {
            S.done = (n < 3) ? 1 : 0; S.have = 0; S.mt_pos = 0; S.map_n = 0; S.overflow = 0; S.best_count = 0; S.scored = 0;
            S.best_tri[0] = S.best_tri[1] = S.best_tri[2] = -1; S.n_inl = 0;
        }
        for (int i = tid; i < 1024; i += blockDim.x) S.map_key[i] = -1;
        __syncthreads();
        int iterations = 0, best = -2147483647; double k = 1.0; unsigned skipped = 0;   /* thread 0 only */
        const unsigned max_skip = (unsigned)d.c.max_iters * 10u;
        while (!S.done) {
            if (tid == 0) {
                /* draw up to SP_ROUND samples (SampleConsensusModel::getSamples / drawIndexSample); no more than the
                 * adaptive rule can still consume */
                int want = SP_ROUND;
                { const double left = k - (double)iterations; if (left < (double)SP_ROUND) want = max(1, (int)ceil(left)); }
                int nr = 0;
                for (; nr < want; ++nr) {
                    bool good = false; int t0 = 0, t1 = 0, t2 = 0;
                    for (int it = 0; it < 1000 && !good; ++it) {
                        if (S.mt_pos + 3 > SP_MT_LEN || S.overflow) { S.overflow = 1; break; }
                        for (int i = 0; i < 3; ++i) {
                            const unsigned r = (unsigned)d.mt[S.mt_pos++];
                            const int j = i + (int)(r % (unsigned)(n - i));
                            const int vi = rs_map_get(S, i), vj = rs_map_get(S, j);
                            rs_map_set(S, i, vj); rs_map_set(S, j, vi);
                        }
                        t0 = rs_map_get(S, 0); t1 = rs_map_get(S, 1); t2 = rs_map_get(S, 2);
                        good = !rs_degenerate(X, Y, Z, t0, t1, t2);
                    }
                    if (!good) { S.tri[nr][0] = -1; ++nr; break; }       /* empty selection: computeModel breaks */
                    S.tri[nr][0] = t0; S.tri[nr][1] = t1; S.tri[nr][2] = t2;
                    const float ax = X[t1] - X[t0], ay = Y[t1] - Y[t0], az = Z[t1] - Z[t0];
                    const float bx = X[t2] - X[t0], by = Y[t2] - Y[t0], bz = Z[t2] - Z[t0];
                    float n0 = ay * bz - az * by, n1 = az * bx - ax * bz, n2 = ax * by - ay * bx;
                    const float nrm = sqrtf((n0 * n0 + n1 * n1) + (n2 * n2 + 0.0f));
                    n0 = n0 / nrm; n1 = n1 / nrm; n2 = n2 / nrm;
                    float m[4] = { n0, n1, n2, 0.f };
                    m[3] = -1 * sp_dot4(n0, n1, n2, 0.0f, X[t0], Y[t0], Z[t0], 1.0f);
                    S.model[nr] = make_float4(m[0], m[1], m[2], m[3]);
                    S.mvalid[nr] = rs_model_valid(d.c, vertical, m) ? 1 : 0;
                }
                S.n_round = nr;
                for (int h = 0; h < SP_ROUND; ++h) S.counts[h] = 0;
            }
            __syncthreads();
            /* score the round: countWithinDistance for every hypothesis in one sweep over the cluster */
            {
                const int nr = S.n_round;
                int cnt[SP_ROUND];
#pragma unroll
                for (int h = 0; h < SP_ROUND; ++h) cnt[h] = 0;
                if (nr > 2) {
                    for (int i = tid; i < n; i += blockDim.x) {
                        const float px = X[i], py = Y[i], pz = Z[i];
#pragma unroll
                        for (int h = 0; h < SP_ROUND; ++h) {
                            const float4 m = S.model[h];
                            const float dd = fabsf(sp_dot4(m.x, m.y, m.z, m.w, px, py, pz, 1.0f));
                            cnt[h] += (h < nr && dd < thr) ? 1 : 0;
                        }
                    }
                } else {
                    for (int i = tid; i < n; i += blockDim.x) {
                        const float px = X[i], py = Y[i], pz = Z[i];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float4 m = S.model[h];
                            const float dd = fabsf(sp_dot4(m.x, m.y, m.z, m.w, px, py, pz, 1.0f));
                            cnt[h] += (h < nr && dd < thr) ? 1 : 0;
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < SP_ROUND; ++h) {
                    const int c = __reduce_add_sync(0xffffffffu, cnt[h]);
                    if (lane == 0 && c) atomicAdd(&S.counts[h], c);
                }
            }
            __syncthreads();
            if (tid == 0) {
                /* replay RandomSampleConsensus::computeModel over the scored round */
                bool done = false;
                for (int h = 0; h < S.n_round && !done; ++h) {
                    if (!(iterations < k && skipped < max_skip)) { done = true; break; }
                    if (S.tri[h][0] < 0) { done = true; break; }
                    const int cnt = S.mvalid[h] ? S.counts[h] : 0;
                    ++S.scored;
                    if (cnt > best) {
                        best = cnt; S.have = 1;
                        S.best_m[0] = S.model[h].x; S.best_m[1] = S.model[h].y; S.best_m[2] = S.model[h].z; S.best_m[3] = S.model[h].w;
                        S.best_tri[0] = S.tri[h][0]; S.best_tri[1] = S.tri[h][1]; S.best_tri[2] = S.tri[h][2]; S.best_count = cnt;
                        const double w = (double)best * (1.0 / (double)n);
                        double p_no = 1.0 - w * w * w;
                        p_no = fmax(2.220446049250313e-16, p_no);
                        p_no = fmin(1.0 - 2.220446049250313e-16, p_no);
                        k = d.c.log_prob / det_log(p_no);
                    }
                    ++iterations;
                    if (iterations > d.c.max_iters) { done = true; break; }
                }
                if (!(iterations < k && skipped < max_skip)) done = true;
                if (S.overflow) { done = true; atomicOr(&d.flags[s >> 1], 2); }
                S.done = done ? 1 : 0;
            }
            __syncthreads();
        }
        int n_inl = 0;
        if (S.have) {
            /* optimizeModelCoefficients: single-pass float mean + covariance over the inliers of the best model, in
             * index order (computeMeanAndCovarianceMatrix) */
            const float m0 = S.best_m[0], m1 = S.best_m[1], m2 = S.best_m[2], m3 = S.best_m[3];
            const bool bvalid = rs_model_valid(d.c, vertical, S.best_m);
            if (bvalid) {
                const int ntiles = (n + RS_TILE - 1) / RS_TILE;
                float acc = 0.f; int mycnt = 0;
                for (int t = 0; t <= ntiles; ++t) {
                    if (wid > 0) {
                        const int e = t * RS_TILE + (tid - 32);
                        if (t < ntiles && e < n) {
                            const float px = X[e], py = Y[e], pz = Z[e];
                            const bool inl = fabsf(sp_dot4(m0, m1, m2, m3, px, py, pz, 1.0f)) < thr;
                            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
                            if (inl) { a = make_float4(px * px, px * py, px * pz, py * py); b = make_float4(py * pz, pz * pz, px, py); c.x = pz; ++mycnt; }
                            float4 *dst = S.stage[t & 1][tid - 32];
                            dst[0] = a; dst[1] = b; dst[2] = c;
                        }
                    } else if (t > 0 && lane < 9) {
                        const int cnt = min(RS_TILE, n - (t - 1) * RS_TILE);
                        const float *src = reinterpret_cast<const float *>(S.stage[(t - 1) & 1]) + lane;
                        int e = 0;
                        for (; e + 8 <= cnt; e += 8) {
                            const float v0 = src[(e + 0) * 12], v1 = src[(e + 1) * 12], v2 = src[(e + 2) * 12], v3 = src[(e + 3) * 12];
                            const float v4 = src[(e + 4) * 12], v5 = src[(e + 5) * 12], v6 = src[(e + 6) * 12], v7 = src[(e + 7) * 12];
                            acc += v0; acc += v1; acc += v2; acc += v3; acc += v4; acc += v5; acc += v6; acc += v7;
                        }
                        for (; e < cnt; ++e) acc += src[e * 12];
                    }
                    __syncthreads();
                }
                if (wid == 0 && lane < 9) S.accu[lane] = acc;
                mycnt = __reduce_add_sync(0xffffffffu, mycnt);
                if (lane == 0 && mycnt) atomicAdd(&S.n_inl, mycnt);
            }
            __syncthreads();
            if (tid == 0) {
                float refined[4] = { m0, m1, m2, m3 };
                if (bvalid && S.n_inl >= 4) {
                    const float cnt = (float)S.n_inl;
                    float a[9];
                    for (int q = 0; q < 9; ++q) a[q] = S.accu[q] / cnt;
                    float cov[9];
                    cov[0] = a[0] - a[6] * a[6]; cov[1] = a[1] - a[6] * a[7]; cov[2] = a[2] - a[6] * a[8];
                    cov[4] = a[3] - a[7] * a[7]; cov[5] = a[4] - a[7] * a[8]; cov[8] = a[5] - a[8] * a[8];
                    cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
                    float ev, evec[3];
                    det_eigen33f<0>(cov, &ev, evec);
                    float r[4] = { evec[0], evec[1], evec[2], 0.f };
                    r[3] = -1 * sp_dot4(evec[0], evec[1], evec[2], 0.0f, a[6], a[7], a[8], 1.0f);
                    if (rs_model_valid(d.c, vertical, r)) { refined[0] = r[0]; refined[1] = r[1]; refined[2] = r[2]; refined[3] = r[3]; }
                }
                S.refined[0] = refined[0]; S.refined[1] = refined[1]; S.refined[2] = refined[2]; S.refined[3] = refined[3];
                S.mvalid[0] = rs_model_valid(d.c, vertical, refined) ? 1 : 0;
            }
            __syncthreads();
        }
        /* selectWithinDistance on the refined model: every warp owns a contiguous chunk; count first */
        const float f0 = S.refined[0], f1 = S.refined[1], f2 = S.refined[2], f3 = S.refined[3];
        const bool rvalid = S.have && S.mvalid[0] != 0;
        const int L = ((n + RS_THREADS - 1) / RS_THREADS) * 32;
        {
            int cnt = 0;
            if (rvalid)
                for (int g = 0; g < L; g += 32) {
                    const int i = wid * L + g + lane;
                    cnt += (i < n && fabsf(sp_dot4(f0, f1, f2, f3, X[i], Y[i], Z[i], 1.0f)) < thr) ? 1 : 0;
                }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0) S.wcnt[wid] = cnt;
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int q = 0; q < RS_NW; ++q) { S.wbase[q] = run; run += S.wcnt[q]; }
            n_inl = run;
            /* plane slots come from a per-segment pool; k_merge orders them by (cluster rank, call) */
            int slot = -1;
            if (S.have && n_inl >= rest) {
                slot = atomicAdd(&d.n_seg[s], 1);
                if (slot >= SP_PSEG) { slot = -1; atomicOr(&d.flags[s >> 1], 4); }
            }
            S.slot = slot; S.n_inl = n_inl;
            const int row = atomicAdd(&d.log_n[s], 1);
            if (row < SP_LOGCAP) {
                int *Lg = d.logs + ((size_t)s * SP_LOGCAP + row) * 10;
                Lg[0] = ci; Lg[1] = call; Lg[2] = n; Lg[3] = S.scored; Lg[4] = S.best_tri[0]; Lg[5] = S.best_tri[1]; Lg[6] = S.best_tri[2];
                Lg[7] = S.best_count; Lg[8] = S.have ? n_inl : 0; Lg[9] = slot >= 0 ? 1 : 0;
            }
        }
        __syncthreads();
        const int slot = S.slot;
        n_inl = S.n_inl;
        ++call;
        if (slot < 0) break;
        {
            /* split the cluster preserving order: inliers -> the plane's point list, the rest -> the twin buffers */
            const unsigned lt = (1u << lane) - 1u;
            int w_in = S.wbase[wid], w_out = wid * L - S.wbase[wid];
            const bool keep_rest = (n - n_inl) >= rest;
            for (int g = 0; g < L; g += 32) {
                const int i = wid * L + g + lane;
                float px = 0.f, py = 0.f, pz = 0.f; int pix = 0; bool inl = false;
                const bool act = i < n;
                if (act) { px = X[i]; py = Y[i]; pz = Z[i]; pix = PIX[i]; inl = fabsf(sp_dot4(f0, f1, f2, f3, px, py, pz, 1.0f)) < thr; }
                const unsigned bi = __ballot_sync(0xffffffffu, act && inl), bo = __ballot_sync(0xffffffffu, act && !inl);
                if (act && inl) { const int q = out_n + w_in + __popc(bi & lt); OUT[q] = pix; OX[q] = px; OY[q] = py; OZ[q] = pz; }
                if (act && !inl && keep_rest) { const int q = w_out + __popc(bo & lt); X2[q] = px; Y2[q] = py; Z2[q] = pz; PIX2[q] = pix; }
                w_in += __popc(bi); w_out += __popc(bo);
            }
            if (tid == 0) {
                SpSegPlane pl;
                pl.coef[0] = f0; pl.coef[1] = f1; pl.coef[2] = f2; pl.coef[3] = f3;
                pl.count = n_inl; pl.off = off + out_n; pl.cluster = ci; pl.call = call - 1;
                d.seg_planes[s * SP_PSEG + slot] = pl;
            }
            out_n += n_inl; n -= n_inl;
            { float *t; t = X; X = X2; X2 = t; t = Y; Y = Y2; Y2 = t; t = Z; Z = Z2; Z2 = t; int *u = PIX; PIX = PIX2; PIX2 = u; }
            __syncthreads();
        }
        if (n < rest) break;
    }
}

/* zero the plane slots before k_ransac */
__global__ void k_clear_planes(SpDev d, int nseg)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nseg * SP_PSEG) d.seg_planes[i].count = 0;
}

/* ==================================================================================================== */
/* mergeSeperatedPlanes, StairPerception.cpp:965-1090: one warp per frame replays the sequential algorithm on the (few)
 * planes; the lanes share the work inside a step (the 20 sample fetches of a plane, the 40 point-plane tests of a pair).
 * libc rand() is replaced by a per-frame restart of glibc's stream from seed 1 (the GUI's srand(0) per redraw,
 * modules/gui/gui.cpp:303, rewinds it the same way); horizontal planes draw first. */
#define MG_WARPS 2
struct MgWarp {
    float sx[SP_PSEG][20], sy[SP_PSEG][20], sz[SP_PSEG][20];
    int tail[SP_PSEG];
};

__device__ __forceinline__ int mg_point(const SpDev &d, int s, const SpSegPlane *chunks, const int *nxt, int head, int k)
{
    int c = head;
    while (c >= 0 && k >= chunks[c].count) { k -= chunks[c].count; c = nxt[c]; }
    if (c < 0) return -1;
    return d.plane_pts[sp_seg_base(d, s) + chunks[c].off + k];
}

__global__ void __launch_bounds__(MG_WARPS * 32) k_merge(SpDev d, int nB)
{
    __shared__ MgWarp SW[MG_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int f = blockIdx.x * MG_WARPS + wid;
    if (f >= nB) return;
    MgWarp &S = SW[wid];
    int rpos = 0;                                             /* warp-uniform position in the rand() stream */
    const float4 *PT = d.xyzw + (size_t)f * d.N;
    for (int hv = 0; hv < 2; ++hv) {
        const int s = 2 * f + hv;
        SpSegPlane *chunks = d.chunks + s * SP_PSEG;
        int *nxt = d.chunk_next + s * SP_PSEG;
        SpMergedPlane *mp = d.merged + s * SP_PSEG;
        /* compact the plane slots into (cluster rank, call) order: rank sort, keys are unique */
        int P = min(d.n_seg[s], SP_PSEG);
        for (int i = lane; i < P; i += 32) {
            const SpSegPlane pl = d.seg_planes[s * SP_PSEG + i];
            int rank = 0;
            for (int j = 0; j < P; ++j) {
                const SpSegPlane &o = d.seg_planes[s * SP_PSEG + j];
                rank += (o.cluster < pl.cluster || (o.cluster == pl.cluster && o.call < pl.call)) ? 1 : 0;
            }
            chunks[rank] = pl;
            nxt[rank] = -1;
            S.tail[rank] = rank;
            SpMergedPlane m0;
            m0.coef[0] = pl.coef[0]; m0.coef[1] = pl.coef[1]; m0.coef[2] = pl.coef[2]; m0.coef[3] = pl.coef[3];
            m0.count = pl.count; m0.head = rank; m0.pad0 = m0.pad1 = 0;
            mp[rank] = m0;
        }
        if (lane == 0) d.n_seg[s] = P;
        __syncwarp();
        /* samples: 20 points per plane, index = rand() % (number of planes) */
        auto resample = [&](int i, int np) {
            if (lane < 20) {
                const int r = d.rnd[(rpos + lane) & (SP_RAND_LEN - 1)];
                const int pi = mg_point(d, s, chunks, nxt, mp[i].head, r % np);
                /* the reference indexes points[rand() % number-of-planes] (StairPerception.cpp:985): past the end of a plane with
                 * fewer points than there are planes (only possible when seg_rest_point < plane count) it reads out of bounds.
                 * Here: a NaN sample, which never counts as near, and a capacity flag on the frame. */
                float4 pw = make_float4(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000), __int_as_float(0x7fc00000), 0.f);
                if (pi >= 0) pw = PT[pi]; else atomicOr(&d.flags[f], 16);
                S.sx[i][lane] = pw.x; S.sy[i][lane] = pw.y; S.sz[i][lane] = pw.z;
            }
            rpos += 20;
            __syncwarp();
        };
        for (int i = 0; i < P; ++i) resample(i, P);
        for (int i = 0; i < P; ++i) {
            for (int j = i + 1; j < P; ++j) {
                const float a1 = mp[i].coef[0], b1 = mp[i].coef[1], c1 = mp[i].coef[2], d1 = mp[i].coef[3];
                const float a2 = mp[j].coef[0], b2 = mp[j].coef[1], c2 = mp[j].coef[2], d2 = mp[j].coef[3];
                const float dot = a1 * a2 + (b1 * b2 + c1 * c2);
                const float angle = det_acosf(dot);
                if (((double)angle < d.c.merge_ath) || (3.14159265358979323846 - (double)angle < d.c.merge_ath)) {
                    bool near = false;
                    if (lane < 20) near = fabsf(a2 * S.sx[i][lane] + b2 * S.sy[i][lane] + c2 * S.sz[i][lane] + d2) <= d.c.merge_thr;
                    int count = __popc(__ballot_sync(0xffffffffu, near));
                    near = false;
                    if (lane < 20) near = fabsf(a1 * S.sx[j][lane] + b1 * S.sy[j][lane] + c1 * S.sz[j][lane] + d1) <= d.c.merge_thr;
                    count += __popc(__ballot_sync(0xffffffffu, near));
                    if (count > 20) {
                        const int n1 = mp[i].count, n2 = mp[j].count, nn = n1 + n2;
                        float a, b, c, dd;
                        if (dot > 0) { a = (a1 * n1 + a2 * n2) / nn; b = (b1 * n1 + b2 * n2) / nn; c = (c1 * n1 + c2 * n2) / nn; dd = (d1 * n1 + d2 * n2) / nn; }
                        else { a = (a1 * n1 - a2 * n2) / nn; b = (b1 * n1 - b2 * n2) / nn; c = (c1 * n1 - c2 * n2) / nn; dd = (d1 * n1 - d2 * n2) / nn; }
                        __syncwarp();
                        if (lane == 0) {
                            nxt[S.tail[i]] = mp[j].head; S.tail[i] = S.tail[j];
                            mp[i].coef[0] = a; mp[i].coef[1] = b; mp[i].coef[2] = c; mp[i].coef[3] = dd; mp[i].count = nn;
                            for (int t = j; t + 1 < P; ++t) { mp[t] = mp[t + 1]; S.tail[t] = S.tail[t + 1]; }
                        }
                        __syncwarp();
                        for (int t = j; t + 1 < P; ++t) {
                            float vx = 0.f, vy = 0.f, vz = 0.f;
                            if (lane < 20) { vx = S.sx[t + 1][lane]; vy = S.sy[t + 1][lane]; vz = S.sz[t + 1][lane]; }
                            __syncwarp();
                            if (lane < 20) { S.sx[t][lane] = vx; S.sy[t][lane] = vy; S.sz[t][lane] = vz; }
                            __syncwarp();
                        }
                        --P;
                        resample(i, P);
                        break;
                    }
                }
            }
        }
        if (lane == 0) d.n_merged[s] = P;
        __syncwarp();
    }
}

/* ==================================================================================================== */
/* computeVectorPlaneInfo, StairPerception.cpp:1196-1325: one CTA per merged plane, points streamed (coalesced) from
 * the plane point lists k_ransac wrote, in chunk-chain order.
 *   centre, covariance: sequential float sums in point order (the reference's loops) -- warp 0 is the "summer"
 *     (lane q owns accumulator q: one dependent FADD per point), warps 1.. stage the addends, double-buffered;
 *   extents: the reference's `if (pj > max) .. else if (pj < min)` scan (:1279-1310) in its exact parallel form:
 *     max = first occurrence of the largest projection; min = first occurrence of the smallest projection among the
 *     points that did NOT set a new running maximum (exclusive prefix-max scan + masked arg-min). */
#define PS_THREADS 256
#define PS_TILE (PS_THREADS - 32)
#define PS_E 4                      /* consecutive points per thread in the extent scan */
struct PsSmem {
    int ch_start[SP_PSEG + 1], ch_off[SP_PSEG];
    int nch;
    float4 stage[2][PS_TILE][2];
    float acc[6], cen[3], evs[9], evl[3];
    float wmax[3][PS_THREADS / 32];
    float carry[3];
    float rv[6][PS_THREADS / 32]; int ri[6][PS_THREADS / 32];
    float emn[3], emx[3]; int imn[3], imx[3];
};

__device__ __forceinline__ int ps_locate(const PsSmem &S, int i, int &c)
{
    while (c + 1 < S.nch && i >= S.ch_start[c + 1]) ++c;
    return S.ch_off[c] + (i - S.ch_start[c]);
}

__global__ void __launch_bounds__(PS_THREADS) k_plane_stats(SpDev d)
{
    __shared__ PsSmem S;
    const int s = blockIdx.y, m = blockIdx.x;
    if (m >= d.n_merged[s]) return;
    const int f = s >> 1, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const SpMergedPlane mp = d.merged[s * SP_PSEG + m];
    const SpSegPlane *chunks = d.chunks + s * SP_PSEG;
    const int *nxt = d.chunk_next + s * SP_PSEG;
    const size_t base = sp_seg_base(d, s);
    const float *PX = d.plane_x + base, *PY = d.plane_y + base, *PZ = d.plane_z + base;
    const int *PP = d.plane_pts + base;
    if (tid == 0) {
        int c = mp.head, k = 0, st = 0;
        while (c >= 0 && k < SP_PSEG) { S.ch_start[k] = st; S.ch_off[k] = chunks[c].off; st += chunks[c].count; c = nxt[c]; ++k; }
        S.ch_start[k] = st; S.nch = k;
    }
    __syncthreads();
    const int n = mp.count;
    const int ntiles = (n + PS_TILE - 1) / PS_TILE;
    /* 1. centre, 2. covariance: two summer passes */
    float cx = 0.f, cy = 0.f, cz = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
        float acc = 0.f;
        const int nacc = pass == 0 ? 3 : 6;
        int cc = 0;
        for (int t = 0; t <= ntiles; ++t) {
            if (wid > 0) {
                const int e = t * PS_TILE + (tid - 32);
                if (t < ntiles && e < n) {
                    const int q = ps_locate(S, e, cc);
                    const float x = PX[q], y = PY[q], z = PZ[q];
                    float4 *dst = S.stage[t & 1][tid - 32];
                    if (pass == 0) dst[0] = make_float4(x, y, z, 0.f);
                    else {
                        const float px = x - cx, py = y - cy, pz = z - cz;
                        dst[0] = make_float4(px * px, px * py, px * pz, py * py);
                        dst[1] = make_float4(py * pz, pz * pz, 0.f, 0.f);
                    }
                }
            } else if (t > 0 && lane < nacc) {
                const int cnt = min(PS_TILE, n - (t - 1) * PS_TILE);
                const float *src = reinterpret_cast<const float *>(S.stage[(t - 1) & 1]) + lane;
                int e = 0;
                for (; e + 8 <= cnt; e += 8) {
                    const float v0 = src[(e + 0) * 8], v1 = src[(e + 1) * 8], v2 = src[(e + 2) * 8], v3 = src[(e + 3) * 8];
                    const float v4 = src[(e + 4) * 8], v5 = src[(e + 5) * 8], v6 = src[(e + 6) * 8], v7 = src[(e + 7) * 8];
                    acc += v0; acc += v1; acc += v2; acc += v3; acc += v4; acc += v5; acc += v6; acc += v7;
                }
                for (; e < cnt; ++e) acc += src[e * 8];
            }
            __syncthreads();
        }
        if (wid == 0 && lane < nacc) { if (pass == 0) S.cen[lane] = acc / (float)n; else S.acc[lane] = acc / (float)n; }
        __syncthreads();
        if (pass == 0) { cx = S.cen[0]; cy = S.cen[1]; cz = S.cen[2]; }
    }
    if (tid == 0) {
        const float cov[9] = { S.acc[0], S.acc[1], S.acc[2], S.acc[1], S.acc[3], S.acc[4], S.acc[2], S.acc[4], S.acc[5] };
        float ev[3], evc[9];
        det_jacobi3f(cov, ev, evc);
        for (int j = 0; j < 3; ++j) { S.evl[j] = ev[j]; for (int k = 0; k < 3; ++k) S.evs[3 * j + k] = evc[k * 3 + j]; }
        S.carry[0] = S.carry[1] = S.carry[2] = -FLT_MAX;
    }
    __syncthreads();
    /* 3. extents along the eigenvectors */
    float ev[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) ev[k] = S.evs[k];
    float bmx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX }, bmn[3] = { FLT_MAX, FLT_MAX, FLT_MAX };
    int imx[3] = { -1, -1, -1 }, imn[3] = { -1, -1, -1 };
    {
        int cc = 0;
        for (int t0 = 0; t0 < n; t0 += PS_THREADS * PS_E) {
            float pj[3][PS_E]; float lm[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
            const int e0 = t0 + tid * PS_E;
#pragma unroll
            for (int u = 0; u < PS_E; ++u) {
                const int e = e0 + u;
                if (e < n) {
                    const int q = ps_locate(S, e, cc);
                    const float px = PX[q] - cx, py = PY[q] - cy, pz = PZ[q] - cz;
#pragma unroll
                    for (int a = 0; a < 3; ++a) { pj[a][u] = px * ev[3 * a] + (py * ev[3 * a + 1] + pz * ev[3 * a + 2]); lm[a] = fmaxf(lm[a], pj[a][u]); }
                } else {
#pragma unroll
                    for (int a = 0; a < 3; ++a) pj[a][u] = __int_as_float(0x7fc00000);
                }
            }
            /* exclusive prefix max over threads (+ carry of the earlier tiles) */
            float ex[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float inc = lm[a];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = fmaxf(inc, v); }
                if (lane == 31) S.wmax[a][wid] = inc;
                float e1 = __shfl_up_sync(0xffffffffu, inc, 1);
                ex[a] = lane == 0 ? -FLT_MAX : e1;
            }
            __syncthreads();
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float pre = S.carry[a];
                for (int q = 0; q < wid; ++q) pre = fmaxf(pre, S.wmax[a][q]);
                float run = fmaxf(pre, ex[a]);
#pragma unroll
                for (int u = 0; u < PS_E; ++u) {
                    const float v = pj[a][u];
                    const int e = e0 + u;
                    if (v > run) { run = v; if (v > bmx[a]) { bmx[a] = v; imx[a] = e; } }
                    else if (v < bmn[a]) { bmn[a] = v; imn[a] = e; }
                }
            }
            __syncthreads();
            if (tid == 0) {
#pragma unroll
                for (int a = 0; a < 3; ++a) { float c = S.carry[a]; for (int q = 0; q < PS_THREADS / 32; ++q) c = fmaxf(c, S.wmax[a][q]); S.carry[a] = c; }
            }
            __syncthreads();
        }
    }
    /* block reduce: (max value, smallest ordinal) and (min value, smallest ordinal) per axis */
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float v = bmx[a]; int i = imx[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_down_sync(0xffffffffu, v, o); const int i2 = __shfl_down_sync(0xffffffffu, i, o);
            if (i2 >= 0 && (i < 0 || v2 > v || (v2 == v && i2 < i))) { v = v2; i = i2; }
        }
        if (lane == 0) { S.rv[a][wid] = v; S.ri[a][wid] = i; }
        v = bmn[a]; i = imn[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_down_sync(0xffffffffu, v, o); const int i2 = __shfl_down_sync(0xffffffffu, i, o);
            if (i2 >= 0 && (i < 0 || v2 < v || (v2 == v && i2 < i))) { v = v2; i = i2; }
        }
        if (lane == 0) { S.rv[3 + a][wid] = v; S.ri[3 + a][wid] = i; }
    }
    __syncthreads();
    if (tid < 6) {
        const bool is_max = tid < 3;
        float v = is_max ? -FLT_MAX : FLT_MAX; int i = -1;
        for (int q = 0; q < PS_THREADS / 32; ++q) {
            const float v2 = S.rv[tid][q]; const int i2 = S.ri[tid][q];
            if (i2 >= 0 && (i < 0 || (is_max ? v2 > v : v2 < v) || (v2 == v && i2 < i))) { v = v2; i = i2; }
        }
        int cc = 0;
        const int pix = i >= 0 ? PP[ps_locate(S, i, cc)] : -1;
        if (is_max) { S.emx[tid] = v; S.imx[tid] = pix; } else { S.emn[tid - 3] = v; S.imn[tid - 3] = pix; }
    }
    __syncthreads();
    if (tid == 0) {
        const float4 *P = d.xyzw + (size_t)f * d.N;
        sp_plane_record r;
        r.type = (s & 1) ? 0 : 1; r.ptype = 3; r.count = n; r.first = 0;
        const float qnan = __int_as_float(0x7fc00000);
        for (int k = 0; k < 4; ++k) r.coef[k] = mp.coef[k];
        for (int k = 0; k < 3; ++k) { r.center[k] = S.cen[k]; r.eval[k] = S.evl[k]; r.pmin[k] = S.emn[k]; r.pmax[k] = S.emx[k]; r.idx_min[k] = S.imn[k]; r.idx_max[k] = S.imx[k]; }
        for (int k = 0; k < 9; ++k) r.evec[k] = S.evs[k];
        for (int a2 = 0; a2 < 3; ++a2) {
            for (int k = 0; k < 3; ++k) { r.pt_min[a2][k] = qnan; r.pt_max[a2][k] = qnan; }
            if (S.imn[a2] >= 0) { const float4 pw = P[S.imn[a2]]; r.pt_min[a2][0] = pw.x; r.pt_min[a2][1] = pw.y; r.pt_min[a2][2] = pw.z; }
            if (S.imx[a2] >= 0) { const float4 pw = P[S.imx[a2]]; r.pt_max[a2][0] = pw.x; r.pt_max[a2][1] = pw.y; r.pt_max[a2][2] = pw.z; }
        }
        for (int k = 0; k < 3; ++k) r.pcenter[k] = (r.pt_max[0][k] + r.pt_min[0][k]) / 2;
        d.info[s * SP_PSEG + m] = r;
    }
}

/* sortPlanesByXValues (StairPerception.cpp:1383-1441) + the per-frame result record; one thread per frame.
 * order[f][i] = (segment-local merged plane index) | hv << 16 is kept for the plane point gather. */
__global__ void __launch_bounds__(64) k_finalize(SpDev d, int nB, int *order, int with_planes, int show_mode)
{
    /* one CTA per frame: thread 0 replays the selection by centre.x (a few tens of planes), all threads copy the records */
    __shared__ int s_src[SP_MAX_PLANES], s_first[SP_MAX_PLANES], s_np;
    const int f = blockIdx.x;
    if (f >= nB) return;
    sp_frame_result &R = d.results[f];
    if (threadIdx.x == 0) {
        const int *c = d.counts + f * SP_NCOUNT;
        R.status = c[0] | (d.flags[f] ? (0x100 | (d.flags[f] << 9)) : 0);
        R.n_valid = c[1]; R.n_crop = c[2]; R.n_voxel = c[3]; R.n_nonan = c[4]; R.n_noleg = c[5]; R.n_h = c[6]; R.n_v = c[7];
        R.n_clusters_h = d.n_clusters[2 * f]; R.n_clusters_v = d.n_clusters[2 * f + 1];
        R.n_seg_h = d.n_seg[2 * f]; R.n_seg_v = d.n_seg[2 * f + 1];
        int nh = d.n_merged[2 * f], nv = d.n_merged[2 * f + 1];
        R.n_planes_h = nh; R.n_planes_v = nv;
        /* process() returns at the ShowModeDef tap (StairPerception.cpp:49-176): nothing computed past it is reported, even
         * where a fused kernel happened to produce it (H and V sets are classified, clustered and segmented together) */
        if (show_mode < SP_SHOW_VOXEL) { R.n_voxel = 0; R.n_nonan = 0; }
        if (show_mode < SP_SHOW_NOLEG) R.n_noleg = 0;
        if (show_mode < SP_SHOW_HORIZONTAL_CLOUD) { R.n_h = 0; R.n_v = 0; }
        if (show_mode < SP_SHOW_SEG_HORIZONTAL_PLANE) { R.n_clusters_h = 0; R.n_seg_h = 0; }
        if (show_mode < SP_SHOW_SEG_VERTICAL_PLANE) { R.n_clusters_v = 0; R.n_seg_v = 0; }
        if (show_mode < SP_SHOW_MERGE_HORIZONTAL_PLANE) R.n_planes_h = 0;
        if (show_mode < SP_SHOW_MERGE_VERTICAL_PLANE) R.n_planes_v = 0;
        if (!with_planes) { nh = 0; nv = 0; }
        float chx[SP_PSEG], cvx[SP_PSEG];
        for (int i = 0; i < nh; ++i) chx[i] = d.info[(2 * f) * SP_PSEG + i].center[0];
        for (int i = 0; i < nv; ++i) cvx[i] = d.info[(2 * f + 1) * SP_PSEG + i].center[0];
        int total = nh + nv, np = 0, first = 0;
        if (total > SP_MAX_PLANES) { total = SP_MAX_PLANES; R.status |= 0x100 | (8 << 9); }
        for (int i = 0; i < total; ++i) {
            float xmax = -FLT_MAX; int hi = 0, vi = 0; bool in_h = false;
            for (int h = 0; h < nh; ++h) if (chx[h] > xmax) { xmax = chx[h]; hi = h; in_h = true; }
            for (int v = 0; v < nv; ++v) if (cvx[v] > xmax) { xmax = cvx[v]; vi = v; in_h = false; }
            if (!in_h && nv == 0) break;
            const int s = in_h ? 2 * f : 2 * f + 1, m = in_h ? hi : vi;
            if (in_h) chx[hi] = -FLT_MAX; else cvx[vi] = -FLT_MAX;
            s_src[np] = s * SP_PSEG + m; s_first[np] = first;
            first += d.info[s * SP_PSEG + m].count;
            order[f * SP_MAX_PLANES + np] = m | ((in_h ? 0 : 1) << 16);
            ++np;
        }
        s_np = np;
        R.n_planes = np; R.n_plane_points = first;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < s_np; i += blockDim.x) {
        sp_plane_record r = d.info[s_src[i]];
        r.first = s_first[i];
        R.planes[i] = r;
    }
}

/* final plane point lists in the reference's order: one CTA per (plane, frame) */
__global__ void __launch_bounds__(256) k_gather_plane_points(SpDev d, const int *order)
{
    const int f = blockIdx.y, pl = blockIdx.x;
    const sp_frame_result &R = d.results[f];
    if (pl >= R.n_planes) return;
    const int o = order[f * SP_MAX_PLANES + pl];
    const int s = 2 * f + (o >> 16), m = o & 0xffff;
    const SpSegPlane *chunks = d.chunks + s * SP_PSEG;
    const int *nxt = d.chunk_next + s * SP_PSEG;
    int *dst = d.plane_idx + (size_t)f * d.N + R.planes[pl].first;
    const int *PP = d.plane_pts + sp_seg_base(d, s);
    int c = d.merged[s * SP_PSEG + m].head, w = 0;
    while (c >= 0) {
        const int cnt = chunks[c].count, off = chunks[c].off;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[w + i] = PP[off + i];
        w += cnt; c = nxt[c];
    }
}

/* ==================================================================================================== */
/* findMinMaxProjPoint, StairPerception.cpp:2780-2808: the two points of a plane whose projection of (p - centre) on a
 * direction is largest / smallest (strict compares: the first such point wins).  One CTA over the plane's final point
 * list; out = { ordinal of max, ordinal of min } (-1 when the plane is empty or no projection is comparable). */
__global__ void __launch_bounds__(256) k_minmax_proj(SpDev d, int frame, int first, int n, float cx, float cy, float cz,
                                                     float vx, float vy, float vz, int *out)
{
    __shared__ float smx[8], smn[8]; __shared__ int imx[8], imn[8];
    const size_t base = (size_t)frame * d.N;
    const int *idx = d.plane_idx + base + first;
    float bmx = -FLT_MAX, bmn = FLT_MAX; int amx = -1, amn = -1;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int pix = idx[i];
        const float4 pw = d.xyzw[base + pix];
        const float px = pw.x - cx, py = pw.y - cy, pz = pw.z - cz;
        const float proj = vx * px + (vy * py + vz * pz);            /* Eigen's unrolled 3-term reduction: c0 + (c1 + c2) */
        if (proj > bmx) { bmx = proj; amx = i; }
        if (proj < bmn) { bmn = proj; amn = i; }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float v2 = __shfl_down_sync(0xffffffffu, bmx, o); const int i2 = __shfl_down_sync(0xffffffffu, amx, o);
        if (i2 >= 0 && (amx < 0 || v2 > bmx || (v2 == bmx && i2 < amx))) { bmx = v2; amx = i2; }
        const float w2 = __shfl_down_sync(0xffffffffu, bmn, o); const int j2 = __shfl_down_sync(0xffffffffu, amn, o);
        if (j2 >= 0 && (amn < 0 || w2 < bmn || (w2 == bmn && j2 < amn))) { bmn = w2; amn = j2; }
    }
    if (lane == 0) { smx[wid] = bmx; imx[wid] = amx; smn[wid] = bmn; imn[wid] = amn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) {
            if (imx[q] >= 0 && (amx < 0 || smx[q] > bmx || (smx[q] == bmx && imx[q] < amx))) { bmx = smx[q]; amx = imx[q]; }
            if (imn[q] >= 0 && (amn < 0 || smn[q] < bmn || (smn[q] == bmn && imn[q] < amn))) { bmn = smn[q]; amn = imn[q]; }
        }
        out[0] = amx; out[1] = amn;
    }
}

/* ==================================================================================================== */
/* Compact records: what a multi-GPU run exchanges and what a consumer of many frames reads.  Frame f of a batch becomes
 * { 16 int32 = the leading counters of sp_frame_result, n_planes x sp_plane_record }, frames back to back.
 * k_compact_offsets: byte offset of every frame (one CTA, exclusive scan); k_compact_copy: one CTA per frame. */
#define SP_COMPACT_HEADER 64
__global__ void __launch_bounds__(1024) k_compact_offsets(const sp_frame_result *res, int nframes, long long *off)
{
    __shared__ long long ws[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int f0 = 0; f0 < nframes; f0 += 1024) {
        const int f = f0 + threadIdx.x;
        const long long v = f < nframes ? (long long)SP_COMPACT_HEADER + (long long)sizeof(sp_plane_record) * res[f].n_planes : 0;
        long long inc = v;
        for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) ws[wid] = inc;
        __syncthreads();
        long long add = carry;
        for (int w = 0; w < wid; ++w) add += ws[w];
        if (f < nframes) off[f] = add + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = add + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[nframes] = carry;
}

__global__ void __launch_bounds__(128) k_compact_copy(const sp_frame_result *res, int nframes, const long long *off, unsigned char *out, long long cap)
{
    const int f = blockIdx.x;
    if (f >= nframes || off[f + 1] > cap) return;             /* a buffer too small is reported by the host from off[nframes] */
    const int words = (int)((off[f + 1] - off[f]) / 4);
    const int *src = reinterpret_cast<const int *>(res + f);
    int *dst = reinterpret_cast<int *>(out + off[f]);
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
}

/* ==================================================================================================== */
/* The consumer's view of a chunk: the point clouds of every hand-off plane of every frame, packed {x, y, z} in the reference's
 * point order, frames back to back (what StairDetection::stairModel reads through Plane::cloud, StairPerception.cpp:1621-2773) */
__global__ void __launch_bounds__(1024) k_plane_cloud_offsets(const sp_frame_result *res, int nframes, long long *off)
{
    __shared__ long long ws[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int f0 = 0; f0 < nframes; f0 += 1024) {
        const int f = f0 + threadIdx.x;
        const long long v = f < nframes ? (long long)res[f].n_plane_points : 0;
        long long inc = v;
        for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) ws[wid] = inc;
        __syncthreads();
        long long add = carry;
        for (int w = 0; w < wid; ++w) add += ws[w];
        if (f < nframes) off[f] = add + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = add + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[nframes] = carry;
}

__global__ void __launch_bounds__(256) k_plane_cloud_pack(SpDev d, const sp_frame_result *res, const long long *off, float *out, long long cap_points)
{
    const int f = blockIdx.y;
    const int n = res[f].n_plane_points;
    if (off[f] + n > cap_points) return;
    const size_t base = (size_t)f * d.N;
    float *o = out + 3 * off[f];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 pw = d.xyzw[base + d.plane_idx[base + i]];
        o[3 * i] = pw.x; o[3 * i + 1] = pw.y; o[3 * i + 2] = pw.z;
    }
}
