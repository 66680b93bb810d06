/*
 * sp_knn.cuh -- a3': k-nearest-neighbour covariance normals for unorganised clouds
 * (StairDetection::normalEstimationOMP, /root/reference/modules/stairperception/StairPerception.cpp:524-541 ->
 * pcl::NormalEstimationOMP: kNN search, single-pass float mean + covariance of the k neighbours in neighbour order,
 * pcl::eigen33 smallest eigenvector, flip towards the sensor origin).
 *
 * Query = search surface = every finite point of the frame.  Exact search on a hashed uniform grid: points are
 * sorted by cell bucket (CUB radix sort), one warp per query walks the cells shell by shell (Chebyshev rings around the
 * query's cell), keeps the best KNN_CAP candidates as a sorted list of 64-bit keys (distance bits << 32 | point index:
 * ascending distance, ties by ascending index -- the oracle's rule) in shared memory, merging 128 candidates at a time
 * with a warp-synchronous bitonic network, and stops as soon as the k-th distance lies inside the radius the scanned
 * shells guarantee.
 */
#pragma once
#include "sp_common.cuh"
#include "sp_seg.cuh"

#define KNN_BITS 21
#define KNN_CAP 128                  /* k <= 128 */
#define KNN_WARPS 8
#define KNN_MAX_SHELL 12

struct SpKnnGrid { float h; int nvalid; int occupied; int pad1; };   /* occupied = hashed cells that hold points (density probe) */

/* levelling + masking + crop predicate / bounding box / counts (what k_ingest_normals does besides the normals) for a
 * cloud without image structure; also the bounding box of ALL finite points (kNN search surface) */
__global__ void __launch_bounds__(256) k_ingest_plain(SpDev d, int *knn_mm)
{
    const int f = blockIdx.y, N = d.N;
    const float4 *in = d.in + (size_t)f * N;
    float R[9];
    const bool rot = d.R != nullptr;
    if (rot) {
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(d.R + 9 * f + i);
    }
    const float qnan = __int_as_float(0x7fc00000);
    int mn[3] = { 0x7fffffff, 0x7fffffff, 0x7fffffff }, mx[3] = { (int)0x80000000, (int)0x80000000, (int)0x80000000 };
    int an[3] = { 0x7fffffff, 0x7fffffff, 0x7fffffff }, ax[3] = { (int)0x80000000, (int)0x80000000, (int)0x80000000 };
    int nvalid = 0, ncrop = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(in + i);
        float ox = p.x, oy = p.y, oz = p.z;
        if (rot) {
            if (!isnan(p.x)) {
                ox = R[0] * p.x + R[1] * p.y + R[2] * p.z;
                oy = R[3] * p.x + R[4] * p.y + R[5] * p.z;
                oz = R[6] * p.x + R[7] * p.y + R[8] * p.z;
            } else { ox = oy = oz = qnan; }
        }
        const unsigned rgba = __float_as_uint(p.w);
        if ((rgba & 255u) == 0u || ((rgba >> 8) & 255u) == 0u || ((rgba >> 16) & 255u) == 0u) { ox = oy = oz = qnan; }
        const size_t gi = (size_t)f * N + i;
        d.xyzw[gi] = make_float4(ox, oy, oz, 0.f);
        d.nrm[gi] = make_float4(qnan, qnan, qnan, 0.f);
        bool crop = false;
        if (sp_finite3(ox, oy, oz)) {
            ++nvalid;
            const int o0 = sp_f2ord(ox), o1 = sp_f2ord(oy), o2 = sp_f2ord(oz);
            an[0] = min(an[0], o0); an[1] = min(an[1], o1); an[2] = min(an[2], o2);
            ax[0] = max(ax[0], o0); ax[1] = max(ax[1], o1); ax[2] = max(ax[2], o2);
            if (!(ox < d.c.x0 || ox > d.c.x1) && !(oy < d.c.y0 || oy > d.c.y1) && !(oz < d.c.z0 || oz > d.c.z1)) {
                ++ncrop;
                crop = true;
                mn[0] = min(mn[0], o0); mn[1] = min(mn[1], o1); mn[2] = min(mn[2], o2);
                mx[0] = max(mx[0], o0); mx[1] = max(mx[1], o1); mx[2] = max(mx[2], o2);
            }
        }
        {   /* cropped points per 1024-point block (the voxel stage compacts by them): a warp's 32 points share a block */
            const unsigned am = __activemask();
            const unsigned bm = __ballot_sync(am, crop);
            if (bm && (threadIdx.x & 31) == __ffs(am) - 1) atomicAdd(&d.vcnt[(size_t)f * d.nblk + (i >> 10)], __popc(bm));
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        mn[k] = __reduce_min_sync(0xffffffffu, mn[k]); mx[k] = __reduce_max_sync(0xffffffffu, mx[k]);
        an[k] = __reduce_min_sync(0xffffffffu, an[k]); ax[k] = __reduce_max_sync(0xffffffffu, ax[k]);
    }
    nvalid = __reduce_add_sync(0xffffffffu, nvalid); ncrop = __reduce_add_sync(0xffffffffu, ncrop);
    if ((threadIdx.x & 31) == 0 && nvalid) {
        atomicAdd(&d.counts[f * SP_NCOUNT + 1], nvalid);
        int *am = knn_mm + f * 8;
        for (int k = 0; k < 3; ++k) { atomicMin(am + k, an[k]); atomicMax(am + 3 + k, ax[k]); }
        if (ncrop) {
            atomicAdd(&d.counts[f * SP_NCOUNT + 2], ncrop);
            int *mm = d.mm + f * 8;
            for (int k = 0; k < 3; ++k) { atomicMin(mm + k, mn[k]); atomicMax(mm + 3 + k, mx[k]); }
        }
    }
}

__global__ void k_knn_init(int *knn_mm, int nB)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nB * 8) { const int k = i & 7; knn_mm[i] = (k < 3) ? 0x7fffffff : (int)0x80000000; }
}

/* cell edge per frame: the radius that holds k neighbours on a surface sampled at the cloud's mean density */
__global__ void k_knn_grid(SpDev d, const int *knn_mm, SpKnnGrid *grid, int k, int nB)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nB) return;
    SpKnnGrid g; g.occupied = 0; g.pad1 = 0;
    g.nvalid = d.counts[f * SP_NCOUNT + 1];
    g.h = 1.0f;
    if (g.nvalid > 0) {
        const int *m = knn_mm + f * 8;
        const float ex = sp_ord2f(m[3]) - sp_ord2f(m[0]), ey = sp_ord2f(m[4]) - sp_ord2f(m[1]), ez = sp_ord2f(m[5]) - sp_ord2f(m[2]);
        const float area = ex * ey + ey * ez + ex * ez;
        float h = sqrtf(fmaxf(1e-12f, area * (float)k / (3.14159265f * (float)g.nvalid)));
        g.h = fminf(fmaxf(h, 1e-4f), 1e6f);
    }
    grid[f] = g;
}

/* second look at the cell edge: the first grid told how many points share an occupied cell (m = surface density x h^2), and
 * k neighbours need pi h^2 density >= k, so h <- h sqrt(k / (pi m)).  The bounding-box guess above can be off by 2x on scenes
 * that are mostly empty space, which quadruples the candidates every query has to rank. */
__global__ void k_knn_refine(SpKnnGrid *grid, int k, int nB)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nB) return;
    SpKnnGrid g = grid[f];
    if (g.nvalid > 0 && g.occupied > 0) {
        const float m = (float)g.nvalid / (float)g.occupied;
        const float h = g.h * sqrtf((float)k / (3.14159265f * m)) * 1.05f;
        g.h = fminf(fmaxf(h, 1e-4f), 1e6f);
    }
    g.occupied = 0;
    grid[f] = g;
}

/* cell of a point: a multiplication by 1 / h instead of a division (the division cost 15 % of k_knn_normals' instructions).  The
 * grid only prunes the search: any partition works as long as binning, query and candidate check agree, and the 0.999 slack of the
 * reach test below covers the rounding of the product at a cell boundary */
__device__ __forceinline__ void sp_knn_cell(float x, float y, float z, float h, int &a, int &b, int &c)
{
    const float ih = 1.0f / h;
    a = (int)floorf(x * ih); b = (int)floorf(y * ih); c = (int)floorf(z * ih);
}
__device__ __forceinline__ unsigned sp_knn_bucket(int a, int b, int c) { return sp_cell_hash(a, b, c) & ((1u << KNN_BITS) - 1u); }

__global__ void __launch_bounds__(256) k_knn_key(SpDev d, const SpKnnGrid *grid)
{
    const int f = blockIdx.y, N = d.N;
    const float h = grid[f].h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const size_t gi = (size_t)f * N + i;
        const float4 pw = d.xyzw[gi];
        const float x = pw.x, y = pw.y, z = pw.z;
        unsigned key = 0xFFFFFFFFu;
        if (sp_finite3(x, y, z)) { int a, b, c; sp_knn_cell(x, y, z, h, a, b, c); key = sp_knn_bucket(a, b, c); }
        d.key_a[gi] = ((unsigned long long)(unsigned)f << 32) | key;
        d.val_a[gi] = (unsigned)i;
    }
}

/* after the sort: bucket ranges + the points in bucket order {x, y, z, point index} */
__global__ void __launch_bounds__(256) k_knn_starts(SpDev d, int *bstart, int *bend, float4 *kpts, SpKnnGrid *grid)
{
    const int f = blockIdx.y, N = d.N;
    const size_t base = (size_t)f * N;
    int *bs = bstart + ((size_t)f << KNN_BITS), *be = bend + ((size_t)f << KNN_BITS);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        const unsigned key = (unsigned)d.key_b[base + j];
        if (key == 0xFFFFFFFFu) continue;
        const unsigned pix = d.val_b[base + j];
        { float4 pw = d.xyzw[base + pix]; pw.w = __int_as_float((int)pix); kpts[base + j] = pw; }
        if (j == 0 || (unsigned)d.key_b[base + j - 1] != key) { bs[key] = j; atomicAdd(&grid[f].occupied, 1); }
        if (j == N - 1 || (unsigned)d.key_b[base + j + 1] != key) be[key] = j + 1;
    }
}

/* warp-synchronous bitonic machinery on 128 u64 keys in shared memory */
/* ascending sort of the first n keys (n = 32, 64 or 128: a batch of waiting candidates is padded to the next of them, a second
 * batch of a query is usually half the size of the first); the keys past n must already be INF */
/* The sort runs in REGISTERS: E = n / 32 consecutive keys per lane (key index = lane * E + r), strides below E are compare-exchanges
 * between a lane's own registers, the others exchange with lane ^ (stride / E) by shuffle.  In shared memory the same network
 * cost four 64-bit accesses per compare-exchange with two-way bank conflicts and made the kernel LSU-bound (l1tex 93 %). */
template <int E>
__device__ __forceinline__ void knn_sort_regs(unsigned long long *A, int lane)
{
    unsigned long long v[E];
#pragma unroll
    for (int r = 0; r < E; ++r) v[r] = A[lane * E + r];
#pragma unroll
    for (int size = 2; size <= 32 * E; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride < E) {
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    if ((r & stride) == 0) {
                        const bool asc = size < E ? ((r & size) == 0) : (((lane * E) & size) == 0);
                        const unsigned long long x = v[r], y = v[r | stride];
                        const bool sw = (x > y) == asc;
                        v[r] = sw ? y : x; v[r | stride] = sw ? x : y;
                    }
                }
            } else {
                const int dl = stride / E;
                const bool lower = (lane & dl) == 0;
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const bool asc = size < E ? ((r & size) == 0) : (((lane * E) & size) == 0);
                    const unsigned long long x = v[r];
                    const unsigned long long y = ((unsigned long long)__shfl_xor_sync(0xffffffffu, (unsigned)(x >> 32), dl) << 32) |
                                                 __shfl_xor_sync(0xffffffffu, (unsigned)x, dl);
                    v[r] = ((x < y) == (lower == asc)) ? x : y;          /* the lower index keeps the minimum of an ascending pair */
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < E; ++r) A[lane * E + r] = v[r];
    __syncwarp();
}
__device__ __forceinline__ void knn_sort128(unsigned long long *A, int lane, int n = 128)
{
    __syncwarp();
    if (n <= 32) knn_sort_regs<1>(A, lane);
    else if (n <= 64) knn_sort_regs<2>(A, lane);
    else knn_sort_regs<4>(A, lane);
}
__device__ __forceinline__ void knn_merge128(unsigned long long *L, const unsigned long long *B, int lane)
{
    /* L, B ascending: keep the 128 smallest of both in L, ascending */
#pragma unroll
    for (int q = lane; q < 128; q += 32) { const unsigned long long a = L[q], b = B[127 - q]; L[q] = a < b ? a : b; }
    __syncwarp();
    for (int stride = 64; stride > 0; stride >>= 1) {
#pragma unroll
        for (int q = lane; q < 64; q += 32) {
            const int i = 2 * q - (q & (stride - 1)), j = i + stride;
            const unsigned long long a = L[i], b = L[j];
            if (a > b) { L[i] = b; L[j] = a; }
        }
        __syncwarp();
    }
}

struct KnnWarp {
    unsigned long long L[KNN_CAP], B[KNN_CAP];
    float4 g[KNN_CAP];                   /* the neighbours' coordinates, in list order (one 16-byte element: the three factor reads of a step hit one bank row) */
    int rofs[32], rst[32];               /* fast path: exclusive candidate offset and first position of each of the 27 cell ranges */
    float pend[32][10];                  /* queries waiting for their eigen-solve: the nine sums + the neighbour count */
    float4 pq[32];                       /*                                        the query point {x, y, z, point index} */
};

/* solvePlaneParameters + flipNormalTowardsViewpoint for the pending queries of a warp, one per lane: pcl::eigen33 is a long scalar
 * routine (double-precision series), run for one query at a time it kept one lane of the warp busy for a sixth of the kernel */
__device__ __forceinline__ void knn_solve_pending(SpDev &d, KnnWarp &W, size_t base, int npend, int lane)
{
    __syncwarp();
    if (lane < npend) {
        const float fc = W.pend[lane][9];
        const float4 q = W.pq[lane];
        float ac[9];
        for (int i = 0; i < 9; ++i) ac[i] = W.pend[lane][i] / fc;
        float cov[9];
        cov[0] = ac[0] - ac[6] * ac[6]; cov[1] = ac[1] - ac[6] * ac[7]; cov[2] = ac[2] - ac[6] * ac[8];
        cov[4] = ac[3] - ac[7] * ac[7]; cov[5] = ac[4] - ac[7] * ac[8]; cov[8] = ac[5] - ac[8] * ac[8];
        cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
        float ev, evec[3];
        det_eigen33f<0>(cov, &ev, evec);
        float fx = evec[0], fy = evec[1], fz = evec[2];
        const float vx = 0.0f - q.x, vy = 0.0f - q.y, vz = 0.0f - q.z;
        const float cs = (vx * fx + vy * fy + vz * fz);
        if (cs < 0) { fx *= -1; fy *= -1; fz *= -1; }
        d.nrm[base + __float_as_int(q.w)] = make_float4(fx, fy, fz, 0.f);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(KNN_WARPS * 32, 4) k_knn_normals(SpDev d, const SpKnnGrid *grid, const int *bstart, const int *bend,
                                                                const float4 *kpts, int k)
{
    __shared__ KnnWarp S[KNN_WARPS];
    const int f = blockIdx.y, N = d.N;
    const size_t base = (size_t)f * N;
    const SpKnnGrid g = grid[f];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    KnnWarp &W = S[wid];
    const int *bs = bstart + ((size_t)f << KNN_BITS), *be = bend + ((size_t)f << KNN_BITS);
    const float4 *P = kpts + base;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned long long INF = 0xFFFFFFFFFFFFFFFFull;
    const int want = min(min(k, KNN_CAP), g.nvalid);
    const float h = g.h;
    int npend = 0;                                                      /* queries of this warp waiting for their eigen-solve (warp-uniform) */
    float pred = 0.f;                                                   /* k-th squared distance expected for the next query (0: none yet) */
    for (int t = blockIdx.x * KNN_WARPS + wid; t < g.nvalid; t += gridDim.x * KNN_WARPS) {
        const float4 q = P[t];
        const int qi = __float_as_int(q.w);
        int a, b, c;
        sp_knn_cell(q.x, q.y, q.z, h, a, b, c);
        /* Fast path (almost every query of a surface scan): the 27 cells around the query are looked up by 27 lanes at once and
         * their ranges concatenated, so that the candidates fill whole warps (a cell holds ~8 points: one cell at a time used a
         * quarter of the lanes); candidates beyond a threshold predicted from the earlier queries of this warp (the k-th distance
         * of a uniformly sampled surface hardly varies) are dropped at once, and ONE sort of at most 128 keys replaces the
         * sort - merge rounds.  Exactness: the filter keeps every key <= threshold, so if at least `want` pass, the `want` smallest
         * of all candidates are among them; the shell-1 guarantee (k-th distance < h) is checked as in the general path below,
         * which still handles everything else (first query, sparse regions, more than 128 survivors). */
        bool fast = false;
        if (pred > 0.f) {
            int js = 0, je = 0;
            unsigned hb = 0xffffffffu - (unsigned)lane;
            if (lane < 27) hb = sp_knn_bucket(a + lane % 3 - 1, b + (lane / 3) % 3 - 1, c + lane / 9 - 1);
            const unsigned same = __match_any_sync(0xffffffffu, hb);     /* two of the 27 cells in one bucket: its range is taken once */
            if (lane < 27 && (same & lt) == 0u) { js = bs[hb]; je = be[hb]; }
            const int cc = max(je - js, 0);
            int inc = cc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
            const int C = __shfl_sync(0xffffffffu, inc, 31);
            W.rofs[lane] = inc - cc; W.rst[lane] = js;
            __syncwarp();
            float tscale = 1.3f;
            for (int attempt = 0; attempt < 3 && !fast; ++attempt) {
                const unsigned long long thr0 = ((unsigned long long)__float_as_uint(pred * tscale) << 32) | 0xFFFFFFFFull;
                int bn = 0;
                for (int c0 = 0; c0 < C; c0 += 32) {
                    const int ci = c0 + lane;
                    bool ok = false; unsigned long long key = INF;
                    if (ci < C) {
                        int kk = 0;                                      /* range of candidate ci: largest kk with rofs[kk] <= ci */
                        if (W.rofs[kk + 16] <= ci) kk += 16;
                        if (W.rofs[kk + 8] <= ci) kk += 8;
                        if (W.rofs[kk + 4] <= ci) kk += 4;
                        if (W.rofs[kk + 2] <= ci) kk += 2;
                        if (W.rofs[kk + 1] <= ci) kk += 1;
                        const float4 p = P[W.rst[kk] + (ci - W.rofs[kk])];
                        int pa, pb, pc;
                        sp_knn_cell(p.x, p.y, p.z, h, pa, pb, pc);
                        if ((unsigned)(pa - a + 1) <= 2u && (unsigned)(pb - b + 1) <= 2u && (unsigned)(pc - c + 1) <= 2u) {   /* buckets may mix cells */
                            const float d0 = q.x - p.x, d1 = q.y - p.y, d2 = q.z - p.z;
                            float dist = 0.f; dist += d0 * d0; dist += d1 * d1; dist += d2 * d2;      /* flann::L2_Simple */
                            key = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned)__float_as_int(p.w);
                            ok = key <= thr0;
                        }
                    }
                    const unsigned bm = __ballot_sync(0xffffffffu, ok);
                    const int at = bn + __popc(bm & lt);
                    if (ok && at < KNN_CAP) W.B[at] = key;
                    bn += __popc(bm);
                }
                if (bn > KNN_CAP || bn < want) {
                    /* too many survivors for one sort, or too few (border of a surface): on a surface the count grows with the
                     * squared radius, so aim the next threshold at 1.2 * want survivors */
                    tscale *= 1.2f * (float)want / (float)max(bn, 4);
                    continue;
                }
                for (int i = bn + lane; i < KNN_CAP; i += 32) W.B[i] = INF;
                __syncwarp();
                knn_sort128(W.B, lane, bn <= 32 ? 32 : (bn <= 64 ? 64 : 128));
                const float kth = __uint_as_float((unsigned)(W.B[want - 1] >> 32)), reach = h * 0.999f;
                fast = want > 0 && kth < reach * reach;                  /* every point nearer than h has been seen */
                if (!fast) break;
            }
        }
        const unsigned long long *res = W.B;
        if (!fast) {
            res = W.L;
            for (int i = lane; i < KNN_CAP; i += 32) W.L[i] = INF;
            __syncwarp();
            int bn = 0;                                  /* candidates waiting in W.B (warp-uniform) */
            unsigned long long thr = INF;                /* current want-th best */
            bool done = false;
            for (int shell = 0; shell <= KNN_MAX_SHELL && !done; ++shell) {
                const int side = 2 * shell + 1, ncell = side * side * side;
                for (int c0 = 0; c0 < ncell; c0 += 32) {
                    /* lane -> one cell of the cube; only the ring at Chebyshev distance == shell is new */
                    const int ci = c0 + lane;
                    int js = 0, je = 0, ta = 0, tb = 0, tc = 0;
                    if (ci < ncell) {
                        const int da = ci % side - shell, db = (ci / side) % side - shell, dc = ci / (side * side) - shell;
                        if (max(abs(da), max(abs(db), abs(dc))) == shell) {
                            ta = a + da; tb = b + db; tc = c + dc;
                            const unsigned hb = sp_knn_bucket(ta, tb, tc);
                            js = bs[hb]; je = be[hb];
                        }
                    }
                    unsigned todo = __ballot_sync(0xffffffffu, je > js);
                    while (todo) {
                        const int src = __ffs(todo) - 1; todo &= todo - 1;
                        const int s0 = __shfl_sync(0xffffffffu, js, src), e0 = __shfl_sync(0xffffffffu, je, src);
                        const int xa = __shfl_sync(0xffffffffu, ta, src), xb = __shfl_sync(0xffffffffu, tb, src), xc = __shfl_sync(0xffffffffu, tc, src);
                        for (int j0 = s0; j0 < e0; j0 += 32) {
                            const int j = j0 + lane;
                            bool ok = false; unsigned long long key = INF;
                            if (j < e0) {
                                const float4 p = P[j];
                                int pa, pb, pc;
                                sp_knn_cell(p.x, p.y, p.z, h, pa, pb, pc);
                                if (pa == xa && pb == xb && pc == xc) {          /* buckets may mix cells: take this cell's points only */
                                    const float d0 = q.x - p.x, d1 = q.y - p.y, d2 = q.z - p.z;
                                    float dist = 0.f; dist += d0 * d0; dist += d1 * d1; dist += d2 * d2;      /* flann::L2_Simple */
                                    key = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned)__float_as_int(p.w);
                                    ok = key < thr;
                                }
                            }
                            const unsigned bm = __ballot_sync(0xffffffffu, ok);
                            if (bm) {
                                if (ok) W.B[bn + __popc(bm & lt)] = key;
                                bn += __popc(bm);
                                if (bn > KNN_CAP - 32) {
                                    for (int i = bn + lane; i < KNN_CAP; i += 32) W.B[i] = INF;
                                    __syncwarp();
                                    knn_sort128(W.B, lane, bn <= 32 ? 32 : (bn <= 64 ? 64 : 128));
                                    knn_merge128(W.L, W.B, lane);
                                    bn = 0;
                                    thr = want > 0 ? W.L[want - 1] : INF;
                                }
                            }
                        }
                    }
                }
                if (bn) {
                    for (int i = bn + lane; i < KNN_CAP; i += 32) W.B[i] = INF;
                    __syncwarp();
                    knn_sort128(W.B, lane, bn <= 32 ? 32 : (bn <= 64 ? 64 : 128));
                    knn_merge128(W.L, W.B, lane);
                    bn = 0;
                    thr = want > 0 ? W.L[want - 1] : INF;
                }
                /* every point nearer than shell * h has been seen */
                if (thr != INF && shell > 0) {
                    const float reach = (float)shell * h * 0.999f;
                    done = __uint_as_float((unsigned)(thr >> 32)) < reach * reach;
                }
            }
            if (!done) {
                /* isolated query: scan every point outside the cube already covered */
                for (int j0 = 0; j0 < g.nvalid; j0 += 32) {
                    const int j = j0 + lane;
                    bool ok = false; unsigned long long key = INF;
                    if (j < g.nvalid) {
                        const float4 p = P[j];
                        int pa, pb, pc;
                        sp_knn_cell(p.x, p.y, p.z, h, pa, pb, pc);
                        if (max(abs(pa - a), max(abs(pb - b), abs(pc - c))) > KNN_MAX_SHELL) {
                            const float d0 = q.x - p.x, d1 = q.y - p.y, d2 = q.z - p.z;
                            float dist = 0.f; dist += d0 * d0; dist += d1 * d1; dist += d2 * d2;
                            key = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned)__float_as_int(p.w);
                            ok = key < thr;
                        }
                    }
                    const unsigned bm = __ballot_sync(0xffffffffu, ok);
                    if (bm) {
                        if (ok) W.B[bn + __popc(bm & lt)] = key;
                        bn += __popc(bm);
                        if (bn > KNN_CAP - 32) {
                            for (int i = bn + lane; i < KNN_CAP; i += 32) W.B[i] = INF;
                            __syncwarp();
                            knn_sort128(W.B, lane, bn <= 32 ? 32 : (bn <= 64 ? 64 : 128));
                            knn_merge128(W.L, W.B, lane);
                            bn = 0;
                            thr = want > 0 ? W.L[want - 1] : INF;
                        }
                    }
                }
                if (bn) {
                    for (int i = bn + lane; i < KNN_CAP; i += 32) W.B[i] = INF;
                    __syncwarp();
                    knn_sort128(W.B, lane, bn <= 32 ? 32 : (bn <= 64 ? 64 : 128));
                    knn_merge128(W.L, W.B, lane);
                    bn = 0;
                }
            }
        }
        if (want > 0) {                                                  /* running estimate of the k-th squared distance, biased towards the small (interior) values */
            const float kth = __uint_as_float((unsigned)(res[want - 1] >> 32));
            if (kth < 3.0e38f) pred = pred <= 0.f ? kth : (kth < pred ? 0.5f * (pred + kth) : 0.9f * pred + 0.1f * kth);
        }
        int cnt = 0;
        for (int i = lane; i < KNN_CAP; i += 32) {
            const unsigned long long key = res[i];
            const bool have = i < want && key != INF;
            if (have) {
                const unsigned pix = (unsigned)key;
                const float4 pw = d.xyzw[base + pix];
                W.g[i] = pw;
            }
            cnt += have ? 1 : 0;
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        __syncwarp();
        /* computeMeanAndCovarianceMatrix: nine float sums in neighbour order, lane q owns accumulator q */
        if (lane < 9) {
            /* accumulators xx xy xz yy yz zz x y z: the two factors of lane q are chosen once, outside the serial loop */
            const float *g0 = reinterpret_cast<const float *>(W.g);
            const float *pu = g0 + (lane < 3 ? 0 : (lane < 5 ? 1 : (lane == 5 ? 2 : lane - 6)));
            const float *pv = g0 + (lane == 0 ? 0 : ((lane == 1 || lane == 3) ? 1 : 2));
            const bool plain = lane >= 6;                               /* sums of x, y, z: factor 1.0f (exact) */
            float acc = 0.f;
            for (int i = 0; i < cnt; ++i) {
                const float u = pu[4 * i], v = plain ? 1.0f : pv[4 * i];
                acc += u * v;
            }
            if (cnt > 0) W.pend[npend][lane] = acc;
        }
        if (cnt > 0) {                                                  /* warp-uniform */
            if (lane == 9) { W.pend[npend][9] = (float)cnt; W.pq[npend] = make_float4(q.x, q.y, q.z, __int_as_float(qi)); }
            if (++npend == 32) { knn_solve_pending(d, W, base, npend, lane); npend = 0; }
        }
        __syncwarp();
    }
    if (npend) knn_solve_pending(d, W, base, npend, lane);
}
