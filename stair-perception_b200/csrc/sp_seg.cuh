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
__device__ __forceinline__ unsigned sp_point_bucket(const float4 &p, float cell, int &a, int &b, int &c)
{
    a = (int)floorf(p.x / cell); b = (int)floorf(p.y / cell); c = (int)floorf(p.z / cell);
    return sp_cell_hash(a, b, c) & ((1u << SP_HASH_BITS) - 1u);
}

__global__ void __launch_bounds__(256) k_cl_count(SpDev d)
{
    const int s = blockIdx.y, n = d.seg_n[s];
    const size_t base = (size_t)s * d.N;
    int *cnt = d.ccnt + ((size_t)s << SP_HASH_BITS);
    const float cell = d.c.cell;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = d.pts[base + i];
        int a, b, c;
        const unsigned h = sp_point_bucket(p, cell, a, b, c);
        atomicAdd(&cnt[h], 1);
        d.parent[base + i] = i;
        d.csize[base + i] = 0;
    }
}

/* exclusive scan of one segment's bucket counts: one CTA of 1024 threads, 64 buckets per thread */
__global__ void __launch_bounds__(1024) k_cl_scan(SpDev d)
{
    __shared__ int wsum[32];
    const int s = blockIdx.x;
    if (d.seg_n[s] == 0) return;
    const int4 *cnt = reinterpret_cast<const int4 *>(d.ccnt + ((size_t)s << SP_HASH_BITS)) + threadIdx.x * 16;
    int *start = d.cstart + (size_t)s * ((1 << SP_HASH_BITS) + 1);
    int4 v[16];
    int tot = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) { v[q] = cnt[q]; tot += v[q].x + v[q].y + v[q].z + v[q].w; }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    int run = wsum[wid] + inc - tot;
    int *o = start + threadIdx.x * 64;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        o[4 * q] = run; run += v[q].x; o[4 * q + 1] = run; run += v[q].y;
        o[4 * q + 2] = run; run += v[q].z; o[4 * q + 3] = run; run += v[q].w;
    }
    if (threadIdx.x == 1023) start[1 << SP_HASH_BITS] = run;
}

__global__ void __launch_bounds__(256) k_cl_scatter(SpDev d)
{
    const int s = blockIdx.y, n = d.seg_n[s];
    const size_t base = (size_t)s * d.N;
    int *cnt = d.ccnt + ((size_t)s << SP_HASH_BITS);
    const int *start = d.cstart + (size_t)s * ((1 << SP_HASH_BITS) + 1);
    const float cell = d.c.cell;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = d.pts[base + i];
        int a, b, c;
        const unsigned h = sp_point_bucket(p, cell, a, b, c);
        const int pos = start[h] + atomicSub(&cnt[h], 1) - 1;          /* drains the counters back to zero */
        d.spts[base + pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
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

/* One thread per point in bucket-major order.  Every unordered pair of points in adjacent cells is tested exactly
 * once: own cell against members with a smaller list position, plus the 13 "forward" neighbour cells.  (Two cells
 * sharing a bucket only cause repeated, idempotent unions.)  Union-find is indexed by list position, so a
 * component's root is its smallest member position whatever the visiting order. */
__global__ void __launch_bounds__(256) k_cl_union(SpDev d)
{
    const int s = blockIdx.y, n = d.seg_n[s];
    const size_t base = (size_t)s * d.N;
    const int *start = d.cstart + (size_t)s * ((1 << SP_HASH_BITS) + 1);
    int *parent = d.parent + base;
    const float4 *sp = d.spts + base;
    const float cell = d.c.cell, r2 = d.c.r2;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const float4 p = sp[t];
        const int oi = __float_as_int(p.w);
        int a, b, c;
        const unsigned h0 = sp_point_bucket(p, cell, a, b, c);
        for (int k = 13; k < 27; ++k) {                                  /* offsets (0,0,0) then the 13 forward ones */
            const int da = k / 9 - 1, db = (k / 3) % 3 - 1, dc = k % 3 - 1;
            const unsigned h = k == 13 ? h0 : (sp_cell_hash(a + da, b + db, c + dc) & ((1u << SP_HASH_BITS) - 1u));
            const int j1 = start[h + 1];
            for (int j = start[h]; j < j1; ++j) {
                const float4 q = sp[j];
                const int oj = __float_as_int(q.w);
                if (k == 13 && oj >= oi) continue;
                const float d0 = p.x - q.x, d1 = p.y - q.y, d2 = p.z - q.z;
                float dist = 0.f; dist += d0 * d0; dist += d1 * d1; dist += d2 * d2;   /* flann::L2_Simple order */
                if (dist < r2 && parent[oi] != parent[oj]) sp_uf_unite(parent, oi, oj);
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_cl_flatten(SpDev d)
{
    const int s = blockIdx.y, n = d.seg_n[s];
    const size_t base = (size_t)s * d.N;
    int *parent = d.parent + base;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int r = i;
        while (parent[r] != r) r = parent[r];
        d.root[base + i] = r;
        atomicAdd(&d.csize[base + r], 1);
    }
}

/* One CTA per segment: count clusters within [min,max]; collect those that can yield a plane (size >= rest),
 * rank them like PCL (size descending; ties: smaller first member), lay out their workspaces and plane slots. */
__global__ void __launch_bounds__(256) k_cl_rank(SpDev d)
{
    __shared__ int c_root[1024], c_size[1024];
    __shared__ int nc, nvalid;
    const int s = blockIdx.x, n = d.seg_n[s];
    const size_t base = (size_t)s * d.N;
    if (threadIdx.x == 0) { nc = 0; nvalid = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (d.root[base + i] != i) continue;
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
        }
    }
}

/* ==================================================================================================== */
/* RANSAC plane extraction: one CTA per (segment, cluster).  Restates pcl::SACSegmentation::segment with
 * RandomSampleConsensus (ransac.hpp), the index-shuffle sampler of SampleConsensusModel (sac_model.h) driven by
 * boost::mt19937(12345), SampleConsensusModelPlane / ParallelPlane and the callers' extract-until-small loops
 * (StairPerception.cpp:773-808, :853-884).  Hypotheses are drawn by one thread, scored SP_ROUND at a time by
 * the whole CTA (register counters -> __reduce_add_sync -> shared atomics), then the sequential
 * "first strictly better wins / adaptive k" rule is replayed exactly. */
struct RsSmem {
    int map_key[1024], map_val[1024];
    float4 model[SP_ROUND];
    int tri[SP_ROUND][3];
    int mvalid[SP_ROUND];
    int counts[SP_ROUND];
    int wsum[33];
    int n_round, done, have, mt_pos, map_n, overflow;
    float best_m[4], refined[4];
    int best_tri[3], best_count, scored;
    float accu[9];
    int n_inl;
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

__global__ void __launch_bounds__(256) k_ransac(SpDev d)
{
    __shared__ RsSmem S;
    const int s = blockIdx.y, ci = blockIdx.x;
    if (ci >= d.n_cand[s]) return;
    const bool vertical = (s & 1) != 0;
    const size_t base = (size_t)s * d.N;
    const int tid = threadIdx.x;
    const int my_root = d.cand_root[s * SP_CMAX + ci];
    const int off = d.cand_off[s * SP_CMAX + ci];
    const int size0 = d.cand_size[s * SP_CMAX + ci];
    float *X = d.wx + base + off, *Y = d.wy + base + off, *Z = d.wz + base + off;
    int *PIX = d.wpix + base + off;
    int *OUT = d.plane_pts + base + off;
    const int nseg = d.seg_n[s];

    /* gather the cluster's members in ascending position order */
    {
        int w = 0;
        for (int i0 = 0; i0 < nseg; i0 += blockDim.x) {
            const int i = i0 + tid;
            const bool mine = i < nseg && d.root[base + i] == my_root;
            int tot; const int p = sp_block_scan_flag(mine, S.wsum, &tot);
            if (mine) { const float4 q = d.pts[base + i]; X[w + p] = q.x; Y[w + p] = q.y; Z[w + p] = q.z; PIX[w + p] = __float_as_int(q.w); }
            w += tot;
        }
    }
    __syncthreads();
    int n = size0, out_n = 0, call = 0;
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
                /* draw up to SP_ROUND samples (SampleConsensusModel::getSamples / drawIndexSample) */
                int nr = 0;
                for (; nr < SP_ROUND; ++nr) {
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
                for (int i = tid; i < n; i += blockDim.x) {
                    const float px = X[i], py = Y[i], pz = Z[i];
#pragma unroll
                    for (int h = 0; h < SP_ROUND; ++h) {
                        const float4 m = S.model[h];
                        const float dd = fabsf(sp_dot4(m.x, m.y, m.z, m.w, px, py, pz, 1.0f));
                        cnt[h] += (h < nr && dd < thr) ? 1 : 0;
                    }
                }
#pragma unroll
                for (int h = 0; h < SP_ROUND; ++h) {
                    const int c = __reduce_add_sync(0xffffffffu, cnt[h]);
                    if ((tid & 31) == 0 && c) atomicAdd(&S.counts[h], c);
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
             * index order (computeMeanAndCovarianceMatrix); thread q owns accumulator q */
            const float m0 = S.best_m[0], m1 = S.best_m[1], m2 = S.best_m[2], m3 = S.best_m[3];
            const bool bvalid = rs_model_valid(d.c, vertical, S.best_m);
            if (tid < 9 && bvalid) {
                float acc = 0.f; int cnt = 0;
                for (int i = 0; i < n; ++i) {
                    const float px = X[i], py = Y[i], pz = Z[i];
                    if (fabsf(sp_dot4(m0, m1, m2, m3, px, py, pz, 1.0f)) < thr) {
                        float t;
                        switch (tid) {
                            case 0: t = px * px; break; case 1: t = px * py; break; case 2: t = px * pz; break;
                            case 3: t = py * py; break; case 4: t = py * pz; break; case 5: t = pz * pz; break;
                            case 6: t = px; break; case 7: t = py; break; default: t = pz; break;
                        }
                        acc += t; ++cnt;
                    }
                }
                S.accu[tid] = acc;
                if (tid == 0) S.n_inl = cnt;
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
            /* selectWithinDistance on the refined model, then split the cluster preserving order */
            const float f0 = S.refined[0], f1 = S.refined[1], f2 = S.refined[2], f3 = S.refined[3];
            const bool rvalid = S.mvalid[0] != 0;
            int cnt = 0;
            for (int i = tid; i < n; i += blockDim.x)
                cnt += (rvalid && fabsf(sp_dot4(f0, f1, f2, f3, X[i], Y[i], Z[i], 1.0f)) < thr) ? 1 : 0;
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (tid == 0) S.counts[0] = 0;
            __syncthreads();
            if ((tid & 31) == 0 && cnt) atomicAdd(&S.counts[0], cnt);
            __syncthreads();
            n_inl = S.counts[0];
        }
        /* plane slots come from a per-segment pool; k_merge orders them by (cluster rank, call) */
        if (tid == 0) {
            int slot = -1;
            if (S.have && n_inl >= rest) {
                slot = atomicAdd(&d.n_seg[s], 1);
                if (slot >= SP_PSEG) { slot = -1; atomicOr(&d.flags[s >> 1], 4); }
            }
            S.n_round = slot;
        }
        __syncthreads();
        const int slot = S.n_round;
        const bool emit = slot >= 0;
        if (tid == 0) {
            const int row = atomicAdd(&d.log_n[s], 1);
            if (row < SP_LOGCAP) {
                int *L = d.logs + ((size_t)s * SP_LOGCAP + row) * 10;
                L[0] = ci; L[1] = call; L[2] = n; L[3] = S.scored; L[4] = S.best_tri[0]; L[5] = S.best_tri[1]; L[6] = S.best_tri[2];
                L[7] = S.best_count; L[8] = S.have ? n_inl : 0; L[9] = emit ? 1 : 0;
            }
        }
        ++call;
        if (!emit) break;
        {
            const float f0 = S.refined[0], f1 = S.refined[1], f2 = S.refined[2], f3 = S.refined[3];
            int w_in = 0, w_out = 0;
            for (int i0 = 0; i0 < n; i0 += blockDim.x) {
                const int i = i0 + tid;
                float px = 0.f, py = 0.f, pz = 0.f; int pix = 0; bool inl = false, act = i < n;
                if (act) { px = X[i]; py = Y[i]; pz = Z[i]; pix = PIX[i]; inl = fabsf(sp_dot4(f0, f1, f2, f3, px, py, pz, 1.0f)) < thr; }
                int tin; const int pin = sp_block_scan_flag(act && inl, S.wsum, &tin);
                int tout; const int pout = sp_block_scan_flag(act && !inl, S.wsum, &tout);
                /* the scans' barriers order all reads of this chunk before the writes below */
                if (act && inl) OUT[out_n + w_in + pin] = pix;
                if (act && !inl) { X[w_out + pout] = px; Y[w_out + pout] = py; Z[w_out + pout] = pz; PIX[w_out + pout] = pix; }
                w_in += tin; w_out += tout;
            }
            if (tid == 0) {
                SpSegPlane pl;
                pl.coef[0] = f0; pl.coef[1] = f1; pl.coef[2] = f2; pl.coef[3] = f3;
                pl.count = w_in; pl.off = off + out_n; pl.cluster = ci; pl.call = call - 1;
                d.seg_planes[s * SP_PSEG + slot] = pl;
            }
            out_n += w_in; n = w_out;
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
/* mergeSeperatedPlanes, StairPerception.cpp:965-1090: one thread per segment replays the sequential algorithm on
 * the (few) planes.  libc rand() is replaced by a per-frame restart of glibc's stream from seed 1 (the GUI's
 * srand(0) per redraw, modules/gui/gui.cpp:303, rewinds it the same way); horizontal planes draw first. */
__device__ __forceinline__ int mg_point(const SpDev &d, int s, const SpSegPlane *chunks, const int *nxt, int head, int k)
{
    int c = head;
    while (c >= 0 && k >= chunks[c].count) { k -= chunks[c].count; c = nxt[c]; }
    if (c < 0) return -1;
    return d.plane_pts[(size_t)s * d.N + chunks[c].off + k];
}

__global__ void k_merge(SpDev d, int nB)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nB) return;
    int rpos = 0;
    const float *X = d.x + (size_t)f * d.N, *Y = d.y + (size_t)f * d.N, *Z = d.z + (size_t)f * d.N;
    for (int hv = 0; hv < 2; ++hv) {
        const int s = 2 * f + hv;
        SpSegPlane *chunks = d.chunks + s * SP_PSEG;
        int *nxt = d.chunk_next + s * SP_PSEG;
        SpMergedPlane *mp = d.merged + s * SP_PSEG;
        /* compact the plane slots (order = cluster rank, then call order) */
        int P = min(d.n_seg[s], SP_PSEG);
        for (int i = 0; i < P; ++i) {                         /* insertion sort by (cluster rank, call) */
            const SpSegPlane pl = d.seg_planes[s * SP_PSEG + i];
            int w = i;
            while (w > 0 && (chunks[w - 1].cluster > pl.cluster || (chunks[w - 1].cluster == pl.cluster && chunks[w - 1].call > pl.call))) { chunks[w] = chunks[w - 1]; --w; }
            chunks[w] = pl;
        }
        for (int i = 0; i < P; ++i) nxt[i] = -1;
        d.n_seg[s] = P;
        int tail[SP_PSEG];
        for (int i = 0; i < P; ++i) {
            mp[i].coef[0] = chunks[i].coef[0]; mp[i].coef[1] = chunks[i].coef[1]; mp[i].coef[2] = chunks[i].coef[2]; mp[i].coef[3] = chunks[i].coef[3];
            mp[i].count = chunks[i].count; mp[i].head = i; tail[i] = i;
        }
        /* samples: 20 points per plane, index = rand() % (number of planes) */
        float sx[SP_PSEG][20], sy[SP_PSEG][20], sz[SP_PSEG][20];
        auto resample = [&](int i, int np) {
            for (int j = 0; j < 20; ++j) {
                const int r = d.rnd[(rpos++) & (SP_RAND_LEN - 1)];
                const int pi = mg_point(d, s, chunks, nxt, mp[i].head, r % np);
                sx[i][j] = X[pi]; sy[i][j] = Y[pi]; sz[i][j] = Z[pi];
            }
        };
        for (int i = 0; i < P; ++i) resample(i, P);
        for (int i = 0; i < P; ++i) {
            const float a1 = mp[i].coef[0], b1 = mp[i].coef[1], c1 = mp[i].coef[2], d1 = mp[i].coef[3];
            float s1x[20], s1y[20], s1z[20];
            for (int k = 0; k < 20; ++k) { s1x[k] = sx[i][k]; s1y[k] = sy[i][k]; s1z[k] = sz[i][k]; }
            for (int j = i + 1; j < P; ++j) {
                const float a2 = mp[j].coef[0], b2 = mp[j].coef[1], c2 = mp[j].coef[2], d2 = mp[j].coef[3];
                const float dot = a1 * a2 + (b1 * b2 + c1 * c2);
                const float angle = det_acosf(dot);
                if (((double)angle < d.c.merge_ath) || (3.14159265358979323846 - (double)angle < d.c.merge_ath)) {
                    int count = 0;
                    for (int k = 0; k < 20; ++k) { const float dd = fabsf(a2 * s1x[k] + b2 * s1y[k] + c2 * s1z[k] + d2); if (dd <= d.c.merge_thr) count++; }
                    for (int k = 0; k < 20; ++k) { const float dd = fabsf(a1 * sx[j][k] + b1 * sy[j][k] + c1 * sz[j][k] + d1); if (dd <= d.c.merge_thr) count++; }
                    if (count > 20) {
                        const int n1 = mp[i].count, n2 = mp[j].count, nn = n1 + n2;
                        float a, b, c, dd;
                        if (dot > 0) { a = (a1 * n1 + a2 * n2) / nn; b = (b1 * n1 + b2 * n2) / nn; c = (c1 * n1 + c2 * n2) / nn; dd = (d1 * n1 + d2 * n2) / nn; }
                        else { a = (a1 * n1 - a2 * n2) / nn; b = (b1 * n1 - b2 * n2) / nn; c = (c1 * n1 - c2 * n2) / nn; dd = (d1 * n1 - d2 * n2) / nn; }
                        nxt[tail[i]] = mp[j].head; tail[i] = tail[j];
                        mp[i].coef[0] = a; mp[i].coef[1] = b; mp[i].coef[2] = c; mp[i].coef[3] = dd; mp[i].count = nn;
                        for (int t = j; t + 1 < P; ++t) {
                            mp[t] = mp[t + 1]; tail[t] = tail[t + 1];
                            for (int k = 0; k < 20; ++k) { sx[t][k] = sx[t + 1][k]; sy[t][k] = sy[t + 1][k]; sz[t][k] = sz[t + 1][k]; }
                        }
                        --P;
                        resample(i, P);
                        break;
                    }
                }
            }
        }
        d.n_merged[s] = P;
    }
}

/* ==================================================================================================== */
/* computeVectorPlaneInfo, StairPerception.cpp:1196-1325: one CTA per merged plane.  The CTA stages the plane's
 * points (chunk chain order) through shared memory; the order-dependent float accumulations are owned by single
 * threads (3 for the centre, 6 for the covariance, 3 for the extents with the reference's else-if rule). */
#define PS_TILE 1024
__global__ void __launch_bounds__(256) k_plane_stats(SpDev d)
{
    __shared__ float tx[PS_TILE], ty[PS_TILE], tz[PS_TILE];
    __shared__ int tp[PS_TILE];
    __shared__ int ch_start[SP_PSEG + 1], ch_off[SP_PSEG];
    __shared__ int nch;
    __shared__ float acc[9], cen[3], evs[9], evl[3];
    __shared__ float emn[3], emx[3]; __shared__ int imn[3], imx[3];
    const int s = blockIdx.y, m = blockIdx.x;
    if (m >= d.n_merged[s]) return;
    const int f = s >> 1, tid = threadIdx.x;
    const SpMergedPlane mp = d.merged[s * SP_PSEG + m];
    const SpSegPlane *chunks = d.chunks + s * SP_PSEG;
    const int *nxt = d.chunk_next + s * SP_PSEG;
    const float *X = d.x + (size_t)f * d.N, *Y = d.y + (size_t)f * d.N, *Z = d.z + (size_t)f * d.N;
    const int *PP = d.plane_pts + (size_t)s * d.N;
    if (tid == 0) {
        int c = mp.head, k = 0, st = 0;
        while (c >= 0 && k < SP_PSEG) { ch_start[k] = st; ch_off[k] = chunks[c].off; st += chunks[c].count; c = nxt[c]; ++k; }
        ch_start[k] = st; nch = k;
    }
    __syncthreads();
    const int n = mp.count;
    auto stage = [&](int t0) {
        for (int e = tid; e < PS_TILE; e += blockDim.x) {
            const int i = t0 + e;
            if (i < n) {
                int c = 0;
                while (c + 1 < nch && i >= ch_start[c + 1]) ++c;
                const int pix = PP[ch_off[c] + (i - ch_start[c])];
                tx[e] = X[pix]; ty[e] = Y[pix]; tz[e] = Z[pix]; tp[e] = pix;
            }
        }
    };
    /* 1. centre */
    float a = 0.f;
    for (int t0 = 0; t0 < n; t0 += PS_TILE) {
        __syncthreads(); stage(t0); __syncthreads();
        const int cnt = min(PS_TILE, n - t0);
        if (tid < 3) { const float *src = tid == 0 ? tx : (tid == 1 ? ty : tz); for (int e = 0; e < cnt; ++e) a += src[e]; }
    }
    if (tid < 3) cen[tid] = a / (float)n;
    __syncthreads();
    const float cx = cen[0], cy = cen[1], cz = cen[2];
    /* 2. covariance (sequential float), eigen-decomposition */
    a = 0.f;
    for (int t0 = 0; t0 < n; t0 += PS_TILE) {
        __syncthreads(); stage(t0); __syncthreads();
        const int cnt = min(PS_TILE, n - t0);
        if (tid < 6) {
            for (int e = 0; e < cnt; ++e) {
                const float px = tx[e] - cx, py = ty[e] - cy, pz = tz[e] - cz;
                float t;
                switch (tid) { case 0: t = px * px; break; case 1: t = px * py; break; case 2: t = px * pz; break;
                               case 3: t = py * py; break; case 4: t = py * pz; break; default: t = pz * pz; break; }
                a += t;
            }
        }
    }
    if (tid < 6) acc[tid] = a / (float)n;
    __syncthreads();
    if (tid == 0) {
        const float cov[9] = { acc[0], acc[1], acc[2], acc[1], acc[3], acc[4], acc[2], acc[4], acc[5] };
        float ev[3], evc[9];
        det_jacobi3f(cov, ev, evc);
        for (int j = 0; j < 3; ++j) { evl[j] = ev[j]; for (int k = 0; k < 3; ++k) evs[3 * j + k] = evc[k * 3 + j]; }
    }
    __syncthreads();
    /* 3. extents along the eigenvectors, with the if / else-if rule of :1279-1310 */
    float mn = FLT_MAX, mx = -FLT_MAX; int in_ = -1, ix_ = -1;
    float e0 = 0.f, e1 = 0.f, e2 = 0.f;
    if (tid < 3) { e0 = evs[3 * tid]; e1 = evs[3 * tid + 1]; e2 = evs[3 * tid + 2]; }
    for (int t0 = 0; t0 < n; t0 += PS_TILE) {
        __syncthreads(); stage(t0); __syncthreads();
        const int cnt = min(PS_TILE, n - t0);
        if (tid < 3) {
            for (int e = 0; e < cnt; ++e) {
                const float px = tx[e] - cx, py = ty[e] - cy, pz = tz[e] - cz;
                const float pj = px * e0 + (py * e1 + pz * e2);
                if (pj > mx) { mx = pj; ix_ = tp[e]; }
                else if (pj < mn) { mn = pj; in_ = tp[e]; }
            }
        }
    }
    if (tid < 3) { emn[tid] = mn; emx[tid] = mx; imn[tid] = in_; imx[tid] = ix_; }
    __syncthreads();
    if (tid == 0) {
        sp_plane_record r;
        r.type = (s & 1) ? 0 : 1; r.ptype = 3; r.count = n; r.first = 0;
        const float qnan = __int_as_float(0x7fc00000);
        for (int k = 0; k < 4; ++k) r.coef[k] = mp.coef[k];
        for (int k = 0; k < 3; ++k) { r.center[k] = cen[k]; r.eval[k] = evl[k]; r.pmin[k] = emn[k]; r.pmax[k] = emx[k]; r.idx_min[k] = imn[k]; r.idx_max[k] = imx[k]; }
        for (int k = 0; k < 9; ++k) r.evec[k] = evs[k];
        for (int a2 = 0; a2 < 3; ++a2) {
            for (int k = 0; k < 3; ++k) { r.pt_min[a2][k] = qnan; r.pt_max[a2][k] = qnan; }
            if (imn[a2] >= 0) { r.pt_min[a2][0] = X[imn[a2]]; r.pt_min[a2][1] = Y[imn[a2]]; r.pt_min[a2][2] = Z[imn[a2]]; }
            if (imx[a2] >= 0) { r.pt_max[a2][0] = X[imx[a2]]; r.pt_max[a2][1] = Y[imx[a2]]; r.pt_max[a2][2] = Z[imx[a2]]; }
        }
        for (int k = 0; k < 3; ++k) r.pcenter[k] = (r.pt_max[0][k] + r.pt_min[0][k]) / 2;
        d.info[s * SP_PSEG + m] = r;
    }
}

/* sortPlanesByXValues (StairPerception.cpp:1383-1441) + the per-frame result record; one thread per frame.
 * order[f][i] = (segment-local merged plane index) | hv << 16 is kept for the plane point gather. */
__global__ void k_finalize(SpDev d, int nB, int *order, int with_planes)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nB) return;
    sp_frame_result &R = d.results[f];
    const int *c = d.counts + f * SP_NCOUNT;
    R.status = c[0] | (d.flags[f] ? (0x100 | (d.flags[f] << 9)) : 0);
    R.n_valid = c[1]; R.n_crop = c[2]; R.n_voxel = c[3]; R.n_nonan = c[4]; R.n_noleg = c[5]; R.n_h = c[6]; R.n_v = c[7];
    R.n_clusters_h = d.n_clusters[2 * f]; R.n_clusters_v = d.n_clusters[2 * f + 1];
    R.n_seg_h = d.n_seg[2 * f]; R.n_seg_v = d.n_seg[2 * f + 1];
    int nh = d.n_merged[2 * f], nv = d.n_merged[2 * f + 1];
    R.n_planes_h = nh; R.n_planes_v = nv;
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
        sp_plane_record r = d.info[s * SP_PSEG + m];
        r.first = first; first += r.count;
        R.planes[np] = r;
        order[f * SP_MAX_PLANES + np] = m | ((in_h ? 0 : 1) << 16);
        ++np;
    }
    R.n_planes = np; R.n_plane_points = first;
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
    const int *PP = d.plane_pts + (size_t)s * d.N;
    int c = d.merged[s * SP_PSEG + m].head, w = 0;
    while (c >= 0) {
        const int cnt = chunks[c].count, off = chunks[c].off;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[w + i] = PP[off + i];
        w += cnt; c = nxt[c];
    }
}
