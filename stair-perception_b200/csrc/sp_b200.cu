/*
 * sp_b200.cu -- host side of libsp_b200.so: context, workspaces, the per-chunk launch sequence, stage taps.
 * C ABI declared in include/sp_b200.h.  No CPU compute path: every processing entry point launches CUDA kernels.
 */
#include "sp_common.cuh"
#include "sp_front.cuh"
#include "sp_strip.cuh"
#include "sp_voxel.cuh"
#include "sp_seg.cuh"
#include "sp_knn.cuh"
#include "sp_stair.h"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <random>
#include <string>
#include <vector>

static_assert(offsetof(sp_frame_result, planes) == SP_COMPACT_HEADER && sizeof(sp_frame_result) == SP_COMPACT_HEADER + SP_MAX_PLANES * sizeof(sp_plane_record),
              "compact records = the 16 leading counters + n_planes plane records");

namespace {

std::string g_create_error;

enum { STG_NORMALS = 0, STG_VOXEL, STG_CLUSTER, STG_RANSAC, STG_PLANES, STG_COUNT };
const char *kStageNames[STG_COUNT] = { "ingest_normals", "voxel", "cluster", "ransac", "merge_stats_finalize" };

/* pcl::deg2rad(float) */
inline float deg2radf(float a) { return a * 0.017453293f; }

/* monotone bijection float <-> int32 */
inline int32_t f2ord(float f) { int32_t b; memcpy(&b, &f, 4); return b >= 0 ? b : (int32_t)(b ^ 0x7fffffff); }
inline float ord2f(int32_t o) { int32_t b = o >= 0 ? o : (int32_t)(o ^ 0x7fffffff); float f; memcpy(&f, &b, 4); return f; }

/* a8 (extractPerpendicularPlanePoints, StairPerception.cpp:696-707) compares angle = acos((double)nx) with two
 * thresholds.  acos is monotone, so each comparison is a half-line in nx; the end points are found once, here,
 * by bisection over the float ordering using the host libm the reference itself would call.  The kernel then
 * classifies with float compares only and gives the same answer as the reference expression for every float. */
template <class Pred> float lowest_true(Pred pred, float lo, float hi)     /* pred monotone: false...true on [lo,hi] */
{
    if (!pred(hi)) return INFINITY;
    if (pred(lo)) return lo;
    int32_t a = f2ord(lo), b = f2ord(hi);                                  /* pred(a) false, pred(b) true */
    while ((int64_t)b - a > 1) { const int32_t m = (int32_t)(((int64_t)a + b) / 2); if (pred(ord2f(m))) b = m; else a = m; }
    return ord2f(b);
}
template <class Pred> float highest_true(Pred pred, float lo, float hi)    /* pred monotone: true...false on [lo,hi] */
{
    if (!pred(lo)) return -INFINITY;
    if (pred(hi)) return hi;
    int32_t a = f2ord(lo), b = f2ord(hi);                                  /* pred(a) true, pred(b) false */
    while ((int64_t)b - a > 1) { const int32_t m = (int32_t)(((int64_t)a + b) / 2); if (pred(ord2f(m))) a = m; else b = m; }
    return ord2f(a);
}

void make_consts(const sp_params &p, SpConst &c)
{
    c.x0 = (float)p.x_min; c.x1 = (float)p.x_max; c.y0 = (float)p.y_min; c.y1 = (float)p.y_max; c.z0 = (float)p.z_min; c.z1 = (float)p.z_max;
    c.ivx = 1.0f / (float)p.voxel_x; c.ivy = 1.0f / (float)p.voxel_y; c.ivz = 1.0f / (float)p.voxel_z;
    const float th = (float)p.noleg_distance;
    c.noleg_th2 = th * th;
    const double t1 = (double)deg2radf((float)p.parallel_angle_diff), t2 = (double)deg2radf((float)p.perpendicular_angle_diff);
    c.h_ge = lowest_true([&](float v) { return acos((double)v) < t1; }, -1.0f, 1.0f);
    c.h_le = highest_true([&](float v) { return M_PI - acos((double)v) < t1; }, -1.0f, 1.0f);
    auto vp = [&](float v) { return fabs(acos((double)v) - M_PI / 2) < t2; };
    c.v_lo = lowest_true(vp, -1.0f, 0.0f);
    c.v_hi = highest_true(vp, 0.0f, 1.0f);
    if (!vp(0.0f)) { c.v_lo = INFINITY; c.v_hi = -INFINITY; }
    const double radius = (double)(float)p.cluster_tolerance;
    c.r2 = (float)(radius * radius);
    c.cell = (float)radius * 1.0001f;
    c.inv_cell = 1.0f / c.cell;
    c.min_c = p.min_cluster_size; c.max_c = p.max_cluster_size;
    c.seg_thr = (float)p.seg_threshold;
    c.sin_angle = fabs(sin((double)deg2radf((float)p.seg_plane_angle_diff)));
    c.max_iters = p.seg_max_iters;
    c.rest = std::max(3, p.seg_rest_point);       /* < 3 makes the reference's extract loop spin forever */
    c.merge_ath = (double)deg2radf((float)p.merge_angle_diff);
    c.merge_thr = (float)p.merge_threshold;
    c.log_prob = det_log(1.0 - 0.99);
}

/* glibc rand(): TYPE_3 additive feedback generator (stdlib/random_r.c), default seed 1 */
void glibc_rand_stream(int n, int *out)
{
    std::vector<int32_t> r(344 + n);
    r[0] = 1;
    for (int i = 1; i < 31; ++i) {
        const int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        int64_t w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (int32_t)w;
    }
    for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
    for (int i = 34; i < 344 + n; ++i) r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
    for (int i = 0; i < n; ++i) out[i] = (int)(((uint32_t)r[344 + i]) >> 1);
}

} // namespace

struct sp_ctx {
    int device = 0, W = 0, H = 0, N = 0, B = 0;
    sp_params params;
    SpDev d;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    std::vector<void *> allocs;
    void *cub_temp = nullptr; size_t cub_bytes = 0;
    float4 *d_in[2] = { nullptr, nullptr }; float *d_R[2] = { nullptr, nullptr };
    int *d_order = nullptr;
    sp_frame_result *h_results = nullptr;     /* pinned, two chunks: the host copies chunk i out while chunk i+1 runs */
    cudaEvent_t ev_res[2] = { nullptr, nullptr };
    cudaEvent_t ev[STG_COUNT + 1];
    cudaEvent_t ev_in[2], ev_done[2];
    float stage_ms[STG_COUNT];
    /* per-kernel trace of the last chunk (SP_TRACE=1 or sp_set_trace): one event after every launch */
    /* a3' kNN normals (unorganised clouds): allocated on first use */
    int normal_mode = 0;
    int force_vbits = 0;                      /* SP_VBITS=n: sort at least n key bits (32 = the four-pass path of wrapped voxel indices) */
    int *knn_mm = nullptr, *knn_bstart = nullptr, *knn_bend = nullptr;
    SpKnnGrid *knn_grid = nullptr;
    float4 *knn_pts = nullptr;
    /* depth-image ingest (sp_process_depth_frames): staging buffers, allocated on first use */
    /* depth-image entry point: a ring of four input buffers (depth, colour, the records k_depth_to_records makes of them, rotations), so
     * that the copies run two chunks ahead of the two lanes (with two buffers a lane waited for its input while the other computed) */
    void *d_depth[4] = { nullptr, nullptr, nullptr, nullptr }; unsigned char *d_color[4] = { nullptr, nullptr, nullptr, nullptr };
    float4 *d_din[4] = { nullptr, nullptr, nullptr, nullptr }; float *d_dR[4] = { nullptr, nullptr, nullptr, nullptr };
    cudaEvent_t evd_in[4] = { nullptr, nullptr, nullptr, nullptr }, evd_done[4] = { nullptr, nullptr, nullptr, nullptr };
    float *d_maps = nullptr;                 /* colmap[W] then rowmap[H] */
    /* K1K2 variant: 0 = tile kernel with plain loads (k_ingest_normals), 1 = the row-marching TMA kernel (sp_strip.cuh),
     * 2 = tile kernel fed by one TMA tensor copy per tile (k_ingest_normals_tma).  SP_K1 = tile | strip | tma picks one. */
    int k1_variant = 2;
    bool use_strip = true;
    int num_sms = 148;
    void *encode_tiled = nullptr;            /* cuTensorMapEncodeTiled, through the runtime's driver entry point */
    bool trace = false;
    std::vector<cudaEvent_t> tev;
    std::vector<const char *> tname;
    int tn = 0;
    /* CUDA graphs of the chunk pipeline for small single-chunk calls (the reference's one-frame-per-call mode): the ~35 launches of
     * a chunk become one graph launch.  Keyed by everything baked into the nodes; SP_GRAPH=0 switches it off. */
    struct ChunkGraph { const void *din; const void *dR; const void *res; int nB, show_mode; cudaGraphExec_t exec; uint64_t used; int64_t launches; };
    std::vector<ChunkGraph> graphs;
    uint64_t graph_clock = 0;
    int graph_max_frames = 8;
    bool capturing = false;
    bool no_taper = false;             /* SP_TAPER=0: host-input batches in uniform chunks */
    int last_nB = 0;
    sp_frame_result *last_res = nullptr;      /* device records of the last chunk: the lane's own array, or a slice of the batch array */
    sp_frame_result *d_batch_res = nullptr;   /* records of a whole batch on the device (compact entry point), grown on demand */
    size_t batch_res_cap = 0;
    long long *d_coff = nullptr; size_t coff_cap = 0;
    float *d_cloud = nullptr, *h_cloud = nullptr; size_t cloud_cap = 0;   /* packed plane clouds of a chunk (batched stair model), points */
    int64_t launches = 0;
    std::string err;
    std::mutex mu;
    /* second lane: batches of two or more chunks alternate between this context and a twin (own workspace, own stream), so that
     * the issue- and latency-bound tail of one chunk (clustering, RANSAC, plane statistics) shares the SMs with the streaming
     * head of the next.  The twin is created on first use and shares the input staging buffers of its owner. */
    sp_ctx *twin = nullptr, *owner = nullptr; /* twins chain: ctx -> twin -> twin's twin (SP_LANES, default 2, at most 4) */
    int want_lanes = 2;
    sp_ctx *last = nullptr;                   /* the lane that ran the last chunk: taps, stair model, stage times read it */
    bool no_twin = false;                     /* SP_NO_TWIN=1 */
};

namespace {

#define ERR(c) ((c)->owner ? (c)->owner->err : (c)->err)      /* a twin reports through its owner */
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ERR(ctx) = std::string(#call) + ": " + cudaGetErrorString(e_); return SP_ERR_CUDA; } } while (0)
inline sp_ctx *lane_of_last_chunk(sp_ctx *ctx) { return ctx->last ? ctx->last : ctx; }

template <class T> cudaError_t dalloc(sp_ctx *ctx, T **p, size_t n)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) { ctx->allocs.push_back(q); *p = (T *)q; }
    return e;
}

int frame_bits(int B) { int b = 0; while ((1 << b) < B) ++b; return b; }

/* trace mark: an event after the launch(es) named `name` (no-op unless tracing) */
void mark(sp_ctx *ctx, const char *name)
{
    if (!ctx->trace || ctx->capturing) return;
    if (ctx->tn >= (int)ctx->tev.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        ctx->tev.push_back(e); ctx->tname.push_back(name);
    }
    ctx->tname[ctx->tn] = name;
    cudaEventRecord(ctx->tev[ctx->tn++], ctx->stream);
}

/* the per-chunk launch sequence; only_stage: -1 = everything `show_mode` asks for, 0 / 1 = one stage group, 2 = all */
int run_chunk(sp_ctx *ctx, const float4 *d_in, const float *d_R, int nB, int show_mode, int only_stage, sp_frame_result *res = nullptr)
{
    SpDev d = ctx->d;
    d.in = d_in; d.R = d_R; d.B = nB;
    if (res) d.results = res;                                 /* records go straight into the batch array */
    ctx->last_res = d.results;
    cudaStream_t st = ctx->stream;
    const int N = ctx->N;
    ctx->last_nB = nB;
    ctx->tn = 0;
    const bool all = only_stage < 0 || only_stage == 2;
    mark(ctx, "begin");
    if (!ctx->capturing) CK(cudaEventRecord(ctx->ev[0], st));
    if ((all || only_stage == 0) && ctx->normal_mode == 0) {
        k_init_frames<<<(nB * SP_NCOUNT + 255) / 256, 256, 0, st>>>(d, nB);
        mark(ctx, "k_init_frames");
        const bool can_tma = ctx->encode_tiled != nullptr && (reinterpret_cast<uintptr_t>(d_in) & 15u) == 0;
        if (ctx->k1_variant == 2 && can_tma) {
            typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                          const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            CUtensorMap tmap;
            const cuuint64_t dims[4] = { 4, (cuuint64_t)ctx->W, (cuuint64_t)ctx->H, (cuuint64_t)nB };
            const cuuint64_t strides[3] = { 16, (cuuint64_t)ctx->W * 16, (cuuint64_t)N * 16 };
            const cuuint32_t box[4] = { 4, K1_RW, K1_RH, 1 }, estr[4] = { 1, 1, 1, 1 };
            const CUresult cr = reinterpret_cast<encode_fn>(ctx->encode_tiled)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float4 *>(d_in), dims, strides, box, estr,
                                                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
            if (cr != CUDA_SUCCESS) { ERR(ctx) = "cuTensorMapEncodeTiled failed (" + std::to_string((int)cr) + ")"; return SP_ERR_CUDA; }
            dim3 g((ctx->W + K1_TW - 1) / K1_TW, (ctx->H + K1_TH - 1) / K1_TH, nB);
            k_ingest_normals_tma<<<g, K1_THREADS, sizeof(K1Smem), st>>>(d, tmap);
            mark(ctx, "k_ingest_normals");
        } else if (ctx->k1_variant == 1 && can_tma) {
            /* one tensor map per chunk: {4 floats, W, H, frames} over the caller's records; a box is half a strip row */
            typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                          const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            CUtensorMap tmap;
            const cuuint64_t dims[4] = { 4, (cuuint64_t)ctx->W, (cuuint64_t)ctx->H, (cuuint64_t)nB };
            const cuuint64_t strides[3] = { 16, (cuuint64_t)ctx->W * 16, (cuuint64_t)N * 16 };
            const cuuint32_t box[4] = { 4, ST_BOX, 1, 1 }, estr[4] = { 1, 1, 1, 1 };
            const CUresult cr = reinterpret_cast<encode_fn>(ctx->encode_tiled)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float4 *>(d_in), dims, strides, box, estr,
                                                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
            if (cr != CUDA_SUCCESS) { ERR(ctx) = "cuTensorMapEncodeTiled failed (" + std::to_string((int)cr) + ")"; return SP_ERR_CUDA; }
            const int ns = (ctx->W + ST_SWMAX - 1) / ST_SWMAX, SW = (ctx->W + ns - 1) / ns;
            const int slots = 2 * ctx->num_sms;                                /* two resident CTAs per SM */
            int nb = (4 * slots + nB * ns - 1) / (nB * ns);                    /* about four work items per CTA slot ... */
            nb = std::max(nb, (ctx->H + 127) / 128);                          /* ... bands of at most 128 rows ... */
            nb = std::max(1, std::min(nb, ctx->H / 16));                      /* ... and of at least 16 (the 12 halo rows of a band) */
            const int BR = (ctx->H + nb - 1) / nb;
            nb = (ctx->H + BR - 1) / BR;
            const int items = nB * ns * nb;
            k_ingest_normals_strip<<<std::min(items, slots), ST_T, sizeof(StSmem), st>>>(d, tmap, ns, SW, nb, BR);
            mark(ctx, "k_ingest_normals");
        } else {
            dim3 g((ctx->W + K1_TW - 1) / K1_TW, (ctx->H + K1_TH - 1) / K1_TH, nB);
            k_ingest_normals<<<g, K1_THREADS, sizeof(K1Smem), st>>>(d);
            mark(ctx, "k_ingest_normals");
        }
        ctx->launches += 2;
    } else if (all || only_stage == 0) {
        /* a3': unorganised cloud, kNN covariance normals */
        const int k = std::max(1, std::min(ctx->params.normal_compute_points, KNN_CAP));
        const int gx = std::min(d.nblk, 148 * 8);
        k_init_frames<<<(nB * SP_NCOUNT + 255) / 256, 256, 0, st>>>(d, nB);
        k_knn_init<<<(nB * 8 + 255) / 256, 256, 0, st>>>(ctx->knn_mm, nB);
        k_ingest_plain<<<dim3(gx, nB), 256, 0, st>>>(d, ctx->knn_mm);
        mark(ctx, "k_ingest_plain");
        k_knn_grid<<<(nB + 127) / 128, 128, 0, st>>>(d, ctx->knn_mm, ctx->knn_grid, k, nB);
        for (int pass = 0; pass < 2; ++pass) {                 /* pass 0 probes the surface density, pass 1 uses the refined cell edge */
            if (pass == 1) k_knn_refine<<<(nB + 127) / 128, 128, 0, st>>>(ctx->knn_grid, k, nB);
            k_knn_key<<<dim3(gx, nB), 256, 0, st>>>(d, ctx->knn_grid);
            size_t tb = ctx->cub_bytes;
            CK(cub::DeviceRadixSort::SortPairs(ctx->cub_temp, tb, d.key_a, d.key_b, d.val_a, d.val_b, (size_t)nB * N, 0, 32 + frame_bits(nB), st));
            CK(cudaMemsetAsync(ctx->knn_bstart, 0, sizeof(int) * ((size_t)nB << KNN_BITS), st));
            CK(cudaMemsetAsync(ctx->knn_bend, 0, sizeof(int) * ((size_t)nB << KNN_BITS), st));
            k_knn_starts<<<dim3(gx, nB), 256, 0, st>>>(d, ctx->knn_bstart, ctx->knn_bend, ctx->knn_pts, ctx->knn_grid);
        }
        mark(ctx, "knn_grid_build(2x)");
        k_knn_normals<<<dim3(148 * 8, nB), KNN_WARPS * 32, 0, st>>>(d, ctx->knn_grid, ctx->knn_bstart, ctx->knn_bend, ctx->knn_pts, k);
        mark(ctx, "k_knn_normals");
        ctx->launches += 6 + 2 * (2 + 2 + (32 + frame_bits(nB) + 7) / 8);
    }
    if (!ctx->capturing) CK(cudaEventRecord(ctx->ev[1], st));
    const bool want_voxel = (all && show_mode >= SP_SHOW_VOXEL) || only_stage == 1;
    if (want_voxel) {
        /* no host round trip: the per-frame sizes and pass counts stay on the device (SpFrameGrid) */
        k_vx_grid<<<(nB + 127) / 128, 128, 0, st>>>(d, nB, ctx->force_vbits);
        k_vx_offsets<<<nB, 256, 0, st>>>(d);
        mark(ctx, "k_vx_grid+offsets");
        k_vx_key<<<dim3(d.nblk, nB), 256, 0, st>>>(d);
        mark(ctx, "k_vx_key");
        for (int p = 0; p < VX_MAXPASS; ++p) {                              /* a frame whose grid needs fewer passes skips the rest */
            k_vx_hist<<<dim3(d.ntile, nB), VX_T, 0, st>>>(d, p);
            k_vx_scan<<<nB, VX_T, 0, st>>>(d, p);
            k_vx_scatter<<<dim3(d.ntile, nB), VX_T, 0, st>>>(d, p);
        }
        mark(ctx, "radix_sort");
        k_vx_pick<<<dim3(d.nblk, nB), 256, 0, st>>>(d);
        mark(ctx, "k_vx_pick");
        k_scan_blocks<<<nB, 256, 0, st>>>(d);
        mark(ctx, "k_scan_blocks");
        k_scatter_lists<<<dim3(d.nblk, nB), 256, 0, st>>>(d);
        mark(ctx, "k_scatter_lists");
        ctx->launches += 6 + 3 * VX_MAXPASS;
    }
    if (!ctx->capturing) CK(cudaEventRecord(ctx->ev[2], st));
    const bool want_seg = all && show_mode >= SP_SHOW_SEG_HORIZONTAL_PLANE;
    const int seg_gx = std::max(SEG_GRID_X, std::min(592, std::max(N / 8192, 1184 / (2 * nB))));   /* CTAs per segment: big single clouds and small batches need more */
    if (want_seg) {
        CK(cudaMemsetAsync(d.ccnt, 0, sizeof(int) * ((size_t)2 * nB << SP_HASH_BITS), st));
        mark(ctx, "memset_ccnt");
        k_cl_count<<<dim3(seg_gx, 2 * nB), 256, 0, st>>>(d);
        mark(ctx, "k_cl_count");
        k_cl_scan<<<2 * nB, 1024, 0, st>>>(d);
        mark(ctx, "k_cl_scan");
        k_cl_scatter<<<dim3(seg_gx, 2 * nB), 256, 0, st>>>(d);
        mark(ctx, "k_cl_scatter");
        k_cl_runs<<<2 * nB, 1024, 0, st>>>(d);
        mark(ctx, "k_cl_runs");
        k_cl_union<<<dim3(seg_gx, 2 * nB), 256, 0, st>>>(d);
        mark(ctx, "k_cl_union");
        k_cl_flatten<<<dim3(seg_gx, 2 * nB), 256, 0, st>>>(d);
        mark(ctx, "k_cl_flatten");
        k_cl_rank<<<2 * nB, RK_THREADS, 0, st>>>(d);
        mark(ctx, "k_cl_rank");
        ctx->launches += 7;
    }
    if (!ctx->capturing) CK(cudaEventRecord(ctx->ev[3], st));
    if (want_seg) {
        k_clear_planes<<<(2 * nB * SP_PSEG + 255) / 256, 256, 0, st>>>(d, 2 * nB);
        mark(ctx, "k_clear_planes");
        k_ransac<<<dim3(SP_CMAX, 2 * nB), RS_THREADS, 0, st>>>(d);
        mark(ctx, "k_ransac");
        ctx->launches += 2;
    }
    if (!ctx->capturing) CK(cudaEventRecord(ctx->ev[4], st));
    if (all) {
        if (want_seg) {
            k_merge<<<(nB + MG_WARPS - 1) / MG_WARPS, MG_WARPS * 32, 0, st>>>(d, nB);
            mark(ctx, "k_merge");
            ctx->launches += 1;
            if (show_mode >= SP_SHOW_STAIR) {
                k_plane_stats<<<dim3(SP_PSEG, 2 * nB), PS_THREADS, 0, st>>>(d);
                mark(ctx, "k_plane_stats");
                ctx->launches += 1;
            }
        }
        k_finalize<<<nB, 64, 0, st>>>(d, nB, ctx->d_order, (want_seg && show_mode >= SP_SHOW_STAIR) ? 1 : 0, show_mode);
        mark(ctx, "k_finalize");
        k_gather_plane_points<<<dim3(SP_MAX_PLANES, nB), 256, 0, st>>>(d, ctx->d_order);
        mark(ctx, "k_gather_plane_points");
        ctx->launches += 2;
    }
    if (!ctx->capturing) CK(cudaEventRecord(ctx->ev[5], st));
    CK(cudaGetLastError());
    return SP_OK;
}

int collect_times(sp_ctx *ctx)
{
    for (int i = 0; i < STG_COUNT; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]) != cudaSuccess) ms = 0.f;
        ctx->stage_ms[i] = ms;
    }
    (void)cudaGetLastError();
    return 0;
}

} // namespace

extern "C" {

const char *sp_last_error(const sp_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int sp_set_params(sp_ctx *ctx, const sp_params *params)
{
    if (!ctx || !params) return SP_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!(params->voxel_x > 0) || !(params->voxel_y > 0) || !(params->voxel_z > 0) || !(params->cluster_tolerance > 0) ||
        params->seg_max_iters < 0) { ERR(ctx) = "invalid parameters"; return SP_ERR_ARG; }
    ctx->params = *params;
    make_consts(*params, ctx->d.c);
    for (auto &q : ctx->graphs) cudaGraphExecDestroy(q.exec);              /* the derived constants are baked into the recorded launches */
    ctx->graphs.clear();
    for (sp_ctx *t = ctx->twin; t; t = t->twin) { t->params = *params; t->d.c = ctx->d.c; }
    return SP_OK;
}

static int create_impl(const sp_params *params, int device, int width, int height, int max_batch, bool staging, sp_ctx **out)
{
    if (!params || !out || width < 1 || height < 1 || max_batch < 1 || max_batch > 4096) { g_create_error = "bad argument"; return SP_ERR_ARG; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        g_create_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "ordinal out of range");
        return SP_ERR_CUDA;
    }
    sp_ctx *ctx = new sp_ctx();
    ctx->device = device; ctx->W = width; ctx->H = height; ctx->N = width * height; ctx->B = max_batch;
    { const char *t = getenv("SP_TRACE"); ctx->trace = t && *t && *t != '0'; }
    { const char *t = getenv("SP_TAPER"); ctx->no_taper = t && *t == '0'; }
    { const char *t = getenv("SP_GRAPH"); if (t && *t) ctx->graph_max_frames = std::max(0, atoi(t)) == 0 ? 0 : std::max(1, atoi(t)); }
    { const char *t = getenv("SP_VBITS"); ctx->force_vbits = t && *t ? std::max(0, std::min(32, atoi(t))) : 0; }
    { const char *t = getenv("SP_NO_TWIN"); ctx->no_twin = t && *t && *t != '0'; }
    { const char *t = getenv("SP_LANES"); if (t && *t) ctx->want_lanes = std::max(1, std::min(4, atoi(t))); if (ctx->want_lanes < 2) ctx->no_twin = true; }
    memset(&ctx->d, 0, sizeof(ctx->d));
    memset(ctx->stage_ms, 0, sizeof(ctx->stage_ms));
    auto fail = [&](const char *what, cudaError_t err) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        sp_destroy(ctx);
        return SP_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    if (sp_set_params(ctx, params) != SP_OK) { g_create_error = ERR(ctx); sp_destroy(ctx); return SP_ERR_ARG; }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    for (int i = 0; i <= STG_COUNT; ++i) if ((e = cudaEventCreate(&ctx->ev[i])) != cudaSuccess) return fail("cudaEventCreate", e);
    for (int i = 0; i < 2; ++i) {
        if ((e = cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
        if ((e = cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    }
    SpDev &d = ctx->d;
    const size_t B = max_batch, N = ctx->N, BN = B * N, S = 2 * B;
    d.W = width; d.H = height; d.N = ctx->N; d.B = max_batch;
    d.nblk = (int)((N + 1023) / 1024);
#define DA(ptr, n) if ((e = dalloc(ctx, &(ptr), (n))) != cudaSuccess) return fail("cudaMalloc " #ptr, e)
    if (staging) { DA(ctx->d_in[0], BN); DA(ctx->d_in[1], BN); DA(ctx->d_R[0], B * 9); DA(ctx->d_R[1], B * 9); }
    DA(d.xyzw, BN); DA(d.nrm, BN);
    DA(d.mm, B * 8); DA(d.counts, B * SP_NCOUNT); DA(d.grid, B); DA(d.flags, B);
    d.ntile = (int)(((N + 3) / 4 * 4 + VX_TILE - 1) / VX_TILE);
    DA(d.vcnt, B * d.nblk); DA(d.voff, B * d.nblk); DA(d.vth, B * d.ntile * VX_BINS);
    DA(d.key_a, BN); DA(d.key_b, BN); DA(d.val_a, BN + 4 * B); DA(d.val_b, BN + 4 * B);      /* k_voxel_frame: slices of (N + 3) & ~3 u32 per frame */
    DA(d.rep, BN); DA(d.code, BN); DA(d.blk, B * d.nblk * 5); DA(d.voxel_idx, BN); DA(d.voxel_code, BN);
    /* per-segment point arrays: the two segments of a frame share one slice of N entries (sp_seg_base) */
    DA(d.pts, BN); DA(d.seg_n, S); DA(d.ccnt, S << SP_HASH_BITS); DA(d.cstart, S * ((1 << SP_HASH_BITS) + 1)); DA(d.spts, BN);
    DA(d.parent, BN); DA(d.root, BN); DA(d.csize, BN);
    DA(d.n_clusters, S); DA(d.n_cand, S);
    DA(d.cand_root, S * SP_CMAX); DA(d.cand_size, S * SP_CMAX); DA(d.cand_off, S * SP_CMAX); DA(d.cand_slot, S * SP_CMAX);
    DA(d.wx, BN); DA(d.wy, BN); DA(d.wz, BN); DA(d.wpix, BN); DA(d.plane_pts, BN);
    DA(d.vx, BN); DA(d.vy, BN); DA(d.vz, BN); DA(d.vpix, BN); DA(d.plane_x, BN); DA(d.plane_y, BN); DA(d.plane_z, BN);
    DA(d.seg_planes, S * SP_PSEG); DA(d.logs, S * SP_LOGCAP * 10); DA(d.log_n, S);
    DA(d.n_seg, S); DA(d.chunks, S * SP_PSEG); DA(d.chunk_next, S * SP_PSEG); DA(d.merged, S * SP_PSEG); DA(d.n_merged, S);
    DA(d.info, S * SP_PSEG); DA(d.results, B); DA(d.plane_idx, BN); DA(ctx->d_order, B * SP_MAX_PLANES);
    int *mt = nullptr, *rnd = nullptr;
    DA(mt, SP_MT_LEN); DA(rnd, SP_RAND_LEN);
#undef DA
    {
        std::vector<int> h(SP_MT_LEN);
        std::mt19937 rng(12345u);                               /* boost::mt19937 seeded 12345 (sac_model.h) */
        for (int i = 0; i < SP_MT_LEN; ++i) h[i] = (int)(rng() >> 1);   /* boost::uniform_int<>(0, INT_MAX) */
        if ((e = cudaMemcpy(mt, h.data(), sizeof(int) * SP_MT_LEN, cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy mt", e);
        std::vector<int> r(SP_RAND_LEN);
        glibc_rand_stream(SP_RAND_LEN, r.data());
        if ((e = cudaMemcpy(rnd, r.data(), sizeof(int) * SP_RAND_LEN, cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy rnd", e);
        d.mt = mt; d.rnd = rnd;
    }
    {
        size_t tb = 0;
        e = cub::DeviceRadixSort::SortPairs(nullptr, tb, d.key_a, d.key_b, d.val_a, d.val_b, BN, 0, 32 + frame_bits(max_batch), ctx->stream);
        if (e != cudaSuccess) return fail("cub temp size", e);
        ctx->cub_bytes = tb;
        void *q = nullptr;
        if ((e = cudaMalloc(&q, std::max<size_t>(tb, 16))) != cudaSuccess) return fail("cudaMalloc cub", e);
        ctx->allocs.push_back(q); ctx->cub_temp = q;
    }
    if ((e = cudaMallocHost((void **)&ctx->h_results, sizeof(sp_frame_result) * B * 2)) != cudaSuccess) return fail("cudaMallocHost", e);
    for (int i = 0; i < 2; ++i) if ((e = cudaEventCreateWithFlags(&ctx->ev_res[i], cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaFuncSetAttribute(k_ingest_normals, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1Smem))) != cudaSuccess)
        return fail("cudaFuncSetAttribute", e);
    if ((e = cudaFuncSetAttribute(k_ingest_normals_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1Smem))) != cudaSuccess)
        return fail("cudaFuncSetAttribute", e);
    if ((e = cudaFuncSetAttribute(k_ingest_normals_strip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StSmem))) != cudaSuccess)
        return fail("cudaFuncSetAttribute", e);
    (void)cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
    {
        const char *t = getenv("SP_K1");
        const std::string want = t ? t : "";
        ctx->k1_variant = want == "tile" ? 0 : want == "strip" ? 1 : 2;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ctx->encode_tiled, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess ||
            !ctx->encode_tiled) { (void)cudaGetLastError(); ctx->encode_tiled = nullptr; ctx->k1_variant = 0; }
    }
    *out = ctx;
    return SP_OK;
}

int sp_create(const sp_params *params, int device, int width, int height, int max_batch, sp_ctx **out)
{
    return create_impl(params, device, width, height, max_batch, true, out);
}

void sp_destroy(sp_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->twin) { sp_destroy(ctx->twin); ctx->twin = nullptr; }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto &q : ctx->graphs) cudaGraphExecDestroy(q.exec);
    ctx->graphs.clear();
    for (void *p : ctx->allocs) cudaFree(p);
    if (ctx->d_batch_res) cudaFree(ctx->d_batch_res);
    if (ctx->d_coff) cudaFree(ctx->d_coff);
    if (ctx->d_cloud) cudaFree(ctx->d_cloud);
    if (ctx->h_cloud) cudaFreeHost(ctx->h_cloud);
    if (ctx->h_results) cudaFreeHost(ctx->h_results);
    for (int i = 0; i <= STG_COUNT; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 2; ++i) { if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]); if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]); }
    for (int i = 0; i < 4; ++i) { if (ctx->evd_in[i]) cudaEventDestroy(ctx->evd_in[i]); if (ctx->evd_done[i]) cudaEventDestroy(ctx->evd_done[i]); }
    for (cudaEvent_t e : ctx->tev) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) if (ctx->ev_res[i]) cudaEventDestroy(ctx->ev_res[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

/* R = m3 * m2 * m1 with the two signed permutations of computeRotationMatrix (test/reader.cpp:233-248): every product
 * term is 0 or +-m2(i,j), so the result is exact whatever the summation order: R(i,j) = s * m2(p,q). */
static void sp_wrap_rotation(const float *m2, float *R9)
{
    const float m1[9] = { 0, 0, 1, -1, 0, 0, 0, -1, 0 }, m3[9] = { 0, 0, -1, 1, 0, 0, 0, -1, 0 };
    float u[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) u[3 * i + j] = (m3[3 * i] * m2[j] + m3[3 * i + 1] * m2[3 + j]) + m3[3 * i + 2] * m2[6 + j];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R9[3 * i + j] = (u[3 * i] * m1[j] + u[3 * i + 1] * m1[3 + j]) + u[3 * i + 2] * m1[6 + j];
}

/* Eigen::QuaternionBase::toRotationMatrix (Geometry/Quaternion.h), coefficient for coefficient */
static void sp_quat_to_matrix(float w, float x, float y, float z, float *m2)
{
    const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z, twx = tx * w, twy = ty * w, twz = tz * w,
                txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    m2[0] = 1.f - (tyy + tzz); m2[1] = txy - twz; m2[2] = txz + twy;
    m2[3] = txy + twz; m2[4] = 1.f - (txx + tzz); m2[5] = tyz - twx;
    m2[6] = txz - twy; m2[7] = tyz + twx; m2[8] = 1.f - (txx + tyy);
}

/* Eigen's float quaternion product as an x86-64 (SSE2) build evaluates it, Geometry/arch/Geometry_SSE.h [Eigen-recall]:
 *   res = (a * b.wwww - a.zxyx * b.yzxx) + sign(+,+,+,-) * (a.yzxz * b.zxyz + a.wwwy * b.xyzy)      (lanes x, y, z, w) */
struct SpQuat { float x, y, z, w; };
static SpQuat sp_quat_mul(const SpQuat &a, const SpQuat &b)
{
    SpQuat r;
    r.x = (a.x * b.w - a.z * b.y) + (a.y * b.z + a.w * b.x);
    r.y = (a.y * b.w - a.x * b.z) + (a.z * b.x + a.w * b.y);
    r.z = (a.z * b.w - a.y * b.x) + (a.x * b.y + a.w * b.z);
    r.w = (a.w * b.w - a.x * b.x) + -(a.z * b.z + a.y * b.y);
    return r;
}

void sp_rotation_from_euler(float pitch, float roll, float yaw, float *R9)
{
    /* m2 = AngleAxisf(roll, X) * AngleAxisf(pitch, Y) * AngleAxisf(yaw, Z) (reader.cpp:240-242).  Eigen evaluates a product of
     * AngleAxis objects as QUATERNION products ((qx * qy) * qz, AngleAxis::operator*) and converts once (Matrix3f = RotationBase
     * -> toRotationMatrix).  Angles: v / 180.0 * M_PI in double, narrowed by the AngleAxisf constructor; the quaternion of an
     * AngleAxis is (cos(angle / 2), sin(angle / 2) * axis) in float (QuaternionBase::operator=(AngleAxis)). */
    const float ar = (float)(roll / 180.0 * M_PI), ap = (float)(pitch / 180.0 * M_PI), ay = (float)(yaw / 180.0 * M_PI);
    const float hr = 0.5f * ar, hp = 0.5f * ap, hy = 0.5f * ay;
    const SpQuat qx = { sinf(hr) * 1.f, sinf(hr) * 0.f, sinf(hr) * 0.f, cosf(hr) };
    const SpQuat qy = { sinf(hp) * 0.f, sinf(hp) * 1.f, sinf(hp) * 0.f, cosf(hp) };
    const SpQuat qz = { sinf(hy) * 0.f, sinf(hy) * 0.f, sinf(hy) * 1.f, cosf(hy) };
    const SpQuat q = sp_quat_mul(sp_quat_mul(qx, qy), qz);
    float m2[9];
    sp_quat_to_matrix(q.w, q.x, q.y, q.z, m2);
    sp_wrap_rotation(m2, R9);
}

void sp_rotation_from_quat(float w, float x, float y, float z, float *R9)
{
    /* quaternionf.normalized() (reader.cpp:267): coeffs() / norm(), norm = sqrt of the packet reduction of the squares in
     * coefficient order (x, y, z, w): SSE2 predux = (x2 + z2) + (y2 + w2) [Eigen-recall, Core/arch/SSE/PacketMath.h] */
    const float n = sqrtf((x * x + z * z) + (y * y + w * w));
    w /= n; x /= n; y /= n; z /= n;
    float m2[9];
    sp_quat_to_matrix(w, x, y, z, m2);
    sp_wrap_rotation(m2, R9);
}

/* the twin lane of a context (created on first use; a batch that does not get one simply runs on a single lane) */
static sp_ctx *twin_of(sp_ctx *ctx, sp_ctx *after)
{
    if (ctx->no_twin || ctx->normal_mode != 0) return nullptr;
    if (after->twin) return after->twin;
    sp_ctx *t = nullptr;
    if (create_impl(&ctx->params, ctx->device, ctx->W, ctx->H, ctx->B, false, &t) != SP_OK) { ctx->no_twin = true; (void)cudaGetLastError(); return nullptr; }
    t->owner = ctx; t->trace = false; t->force_vbits = ctx->force_vbits;
    after->twin = t;
    return t;
}

/* the lanes a batch of nchunks chunks runs on: the context itself, then as many twins as wanted and as the batch can use */
static int lanes_for(sp_ctx *ctx, int nchunks, sp_ctx **lanes)
{
    int nl = 1;
    lanes[0] = ctx;
    if (ctx->trace) return nl;
    while (nl < ctx->want_lanes && nl < nchunks) {
        sp_ctx *t = twin_of(ctx, lanes[nl - 1]);
        if (!t) break;
        lanes[nl++] = t;
    }
    return nl;
}

/* Chunk schedule of a host-input batch: ctx->B frames per chunk, but a batch of several chunks ends with halved ones (..., 64, 32):
 * the kernels of the LAST chunk are the only ones no later H2D copy hides -- 1.4 ms of exposed work after a 32-frame chunk instead
 * of 13 ms after 512.  (Doubling chunks at the head, to shorten the first exposed copy, measured no gain for depth input and
 * -2 % for clouds.)  SP_TAPER=0: uniform chunks. */
static void chunk_schedule(const sp_ctx *ctx, int nframes, bool host_in, std::vector<int> &cstart, std::vector<int> &csize)
{
    const int B = ctx->B;
    const bool taper = host_in && nframes > B && !ctx->no_taper;
    int f = 0;
    while (f < nframes) {
        int c = std::min(B, nframes - f);
        if (taper && nframes - f <= B && c > 32) c = std::max(32, (c + 1) / 2);
        cstart.push_back(f); csize.push_back(c);
        f += c;
    }
}

/* results != NULL: fixed-size records to host memory, chunk by chunk behind the kernels.  results == NULL (compact entry
 * point): the records of the whole batch stay on the device in ctx->d_batch_res. */
static int process_impl(sp_ctx *ctx, const void *xyzrgba, const float *R9, int nframes, int show_mode, sp_frame_result *results, bool host_in, bool keep_on_device = false)
{
    if (!ctx || !xyzrgba || nframes < 0 || show_mode < 0 || show_mode > SP_SHOW_STAIR || (!results && !keep_on_device && nframes > 0)) { if (ctx) ERR(ctx) = "bad argument"; return SP_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    const size_t fb = (size_t)ctx->N * 16;
    if (keep_on_device) {
        if ((size_t)nframes > ctx->batch_res_cap) {
            if (ctx->d_batch_res) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->d_batch_res); ctx->d_batch_res = nullptr; ctx->batch_res_cap = 0; }
            CK(cudaMalloc((void **)&ctx->d_batch_res, sizeof(sp_frame_result) * (size_t)nframes));
            ctx->batch_res_cap = (size_t)nframes;
        }
        if (show_mode == SP_SHOW_ORIGINAL) { CK(cudaMemsetAsync(ctx->d_batch_res, 0, sizeof(sp_frame_result) * (size_t)nframes, ctx->stream)); return SP_OK; }
    } else if (show_mode == SP_SHOW_ORIGINAL) { memset(results, 0, sizeof(sp_frame_result) * (size_t)nframes); return SP_OK; }
    int buf = 0;
    /* chunk loop; with host input the H2D copy of chunk i+1 (copy stream) overlaps the kernels of chunk i; with two or more
     * chunks the chunks alternate between two lanes (this context and its twin), each with its own workspace and stream */
    std::vector<int> cstart, csize;
    chunk_schedule(ctx, nframes, host_in, cstart, csize);
    const int nchunks = (int)cstart.size();
    sp_ctx *lanes[4];
    const int nl = lanes_for(ctx, nchunks, lanes);
    auto issue_copy = [&](int ci, int b) -> int {
        const int f0 = cstart[ci], nB = csize[ci];
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0));
        CK(cudaMemcpyAsync(ctx->d_in[b], (const uint8_t *)xyzrgba + fb * f0, fb * nB, cudaMemcpyHostToDevice, ctx->copy_stream));
        if (R9) CK(cudaMemcpyAsync(ctx->d_R[b], R9 + 9 * (size_t)f0, sizeof(float) * 9 * nB, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->ev_in[b], ctx->copy_stream));
        return SP_OK;
    };
    /* results travel behind the kernels: a chunk's records are copied out of their pinned buffer while later chunks run */
    struct Pending { cudaEvent_t ev; const sp_frame_result *src; int f0, n; };
    std::vector<Pending> pend;
    size_t drained = 0;
    auto drain = [&](size_t keep) -> int {
        while (pend.size() - drained > keep) {
            const Pending &q = pend[drained++];
            CK(cudaEventSynchronize(q.ev));
            memcpy(results + q.f0, q.src, sizeof(sp_frame_result) * (size_t)q.n);
        }
        return SP_OK;
    };
    /* an error in the middle of a batch must not leave copies in flight on the caller's buffers: every exit of the chunk loop
     * goes through `settle`, which drains the copy stream and the lanes and forgets the half-finished chunk */
    auto settle = [&](int rc) -> int {
        if (rc != SP_OK) {
            cudaStreamSynchronize(ctx->copy_stream);
            for (int l = 0; l < nl; ++l) cudaStreamSynchronize(lanes[l]->stream);
            (void)cudaGetLastError();
            ctx->last = nullptr;
            for (int l = 0; l < nl; ++l) lanes[l]->last_nB = 0;
        }
        return rc;
    };
    auto run_batch = [&]() -> int {
    if (host_in && nchunks > 0) { const int rc = issue_copy(0, 0); if (rc) return rc; }
    for (int ci = 0; ci < nchunks; ++ci) {
        const int f0 = cstart[ci], nB = csize[ci];
        sp_ctx *L = lanes[ci % nl];
        const int hb = (ci / nl) & 1;                          /* the lane's pinned result buffer this chunk uses */
        const float4 *din; const float *dR;
        if (host_in) {
            CK(cudaStreamWaitEvent(L->stream, ctx->ev_in[buf], 0));
            din = ctx->d_in[buf]; dR = R9 ? ctx->d_R[buf] : nullptr;
            if (ci + 1 < nchunks) { const int rc = issue_copy(ci + 1, buf ^ 1); if (rc) return rc; }
        } else {
            din = (const float4 *)((const uint8_t *)xyzrgba + fb * f0); dR = R9 ? R9 + 9 * (size_t)f0 : nullptr;
        }
        sp_frame_result *dres = keep_on_device ? ctx->d_batch_res + f0 : L->d.results;
        CK(cudaMemsetAsync(dres, 0, sizeof(sp_frame_result) * nB, L->stream));
        int rc = SP_OK;
        if (nchunks == 1 && nB <= ctx->graph_max_frames && !ctx->trace && L == ctx && ctx->normal_mode == 0) {
            /* small single-chunk call: replay (or record once) the chunk's launch sequence as a CUDA graph */
            sp_ctx::ChunkGraph *cg = nullptr;
            for (auto &q : ctx->graphs) if (q.din == din && q.dR == dR && q.res == dres && q.nB == nB && q.show_mode == show_mode) cg = &q;
            if (!cg) {
                cudaGraph_t graph = nullptr;
                const int64_t l0 = ctx->launches;
                CK(cudaStreamBeginCapture(L->stream, cudaStreamCaptureModeThreadLocal));
                ctx->capturing = true;
                rc = run_chunk(L, din, dR, nB, show_mode, -1, keep_on_device ? dres : nullptr);
                ctx->capturing = false;
                const cudaError_t ce = cudaStreamEndCapture(L->stream, &graph);
                if (rc == SP_OK && ce != cudaSuccess) { ERR(ctx) = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce); rc = SP_ERR_CUDA; }
                if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                cudaGraphExec_t exec = nullptr;
                const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ie != cudaSuccess) { ERR(ctx) = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie); return SP_ERR_CUDA; }
                if (ctx->graphs.size() >= 8) {                             /* evict the least recently used */
                    size_t lru = 0;
                    for (size_t k = 1; k < ctx->graphs.size(); ++k) if (ctx->graphs[k].used < ctx->graphs[lru].used) lru = k;
                    cudaGraphExecDestroy(ctx->graphs[lru].exec);
                    ctx->graphs.erase(ctx->graphs.begin() + (long)lru);
                }
                ctx->graphs.push_back({ din, dR, dres, nB, show_mode, exec, 0, ctx->launches - l0 });
                ctx->launches = l0;
                cg = &ctx->graphs.back();
            }
            cg->used = ++ctx->graph_clock;
            CK(cudaGraphLaunch(cg->exec, L->stream));
            ctx->launches += cg->launches;
            ctx->last_nB = nB; ctx->last_res = dres; ctx->tn = 0;
        } else rc = run_chunk(L, din, dR, nB, show_mode, -1, keep_on_device ? dres : nullptr);
        if (rc) return rc;
        CK(cudaEventRecord(ctx->ev_done[buf], L->stream));
        if (!keep_on_device) {
            CK(cudaMemcpyAsync(L->h_results + (size_t)hb * ctx->B, L->d.results, sizeof(sp_frame_result) * nB, cudaMemcpyDeviceToHost, L->stream));
            CK(cudaEventRecord(L->ev_res[hb], L->stream));
            pend.push_back({ L->ev_res[hb], L->h_results + (size_t)hb * ctx->B, f0, nB });
            { const int rd = drain((size_t)nl); if (rd) return rd; }
        }
        buf ^= 1;
    }
    { const int rd = drain(0); if (rd) return rd; }
    if (nchunks > 0) {
        for (int l = 0; l < nl; ++l) CK(cudaStreamSynchronize(lanes[l]->stream));
        ctx->last = lanes[(nchunks - 1) % nl];
        collect_times(ctx->last);
    }
    return SP_OK;
    };
    return settle(run_batch());
}

int sp_process_depth_frames(sp_ctx *ctx, const void *depth, int depth_type, const uint8_t *bgrx, const sp_intrinsics *K, int mirror,
                            const float *R9, int nframes, int show_mode, sp_frame_result *results)
{
    if (!ctx || !depth || !K || (depth_type != 0 && depth_type != 1) || nframes < 0 || show_mode < 0 || show_mode > SP_SHOW_STAIR ||
        (!results && nframes > 0) || !(K->fx != 0.f) || !(K->fy != 0.f)) { if (ctx) ERR(ctx) = "bad argument"; return SP_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->last = nullptr;                                           /* this entry point runs on the context's own lane */
    CK(cudaSetDevice(ctx->device));
    if (show_mode == SP_SHOW_ORIGINAL) { memset(results, 0, sizeof(sp_frame_result) * (size_t)nframes); return SP_OK; }
    const size_t N = ctx->N, B = ctx->B, dsz = depth_type == 0 ? 4 : 2;
    if (!ctx->d_maps) {
        cudaError_t e;
        for (int i = 0; i < 4; ++i) {
            void *q = nullptr;
            if ((e = cudaMalloc(&q, B * N * 4)) != cudaSuccess) { ERR(ctx) = std::string("cudaMalloc (depth staging): ") + cudaGetErrorString(e); return SP_ERR_CUDA; }
            ctx->allocs.push_back(q); ctx->d_depth[i] = q;
            if ((e = cudaMalloc(&q, B * N * 4)) != cudaSuccess) { ERR(ctx) = std::string("cudaMalloc (colour staging): ") + cudaGetErrorString(e); return SP_ERR_CUDA; }
            ctx->allocs.push_back(q); ctx->d_color[i] = (unsigned char *)q;
            if (i < 2 && ctx->d_in[i] && ctx->d_R[i]) { ctx->d_din[i] = ctx->d_in[i]; ctx->d_dR[i] = ctx->d_R[i]; }
            else {
                if ((e = cudaMalloc(&q, B * N * 16)) != cudaSuccess) { ERR(ctx) = std::string("cudaMalloc (record staging): ") + cudaGetErrorString(e); return SP_ERR_CUDA; }
                ctx->allocs.push_back(q); ctx->d_din[i] = (float4 *)q;
                if ((e = cudaMalloc(&q, B * 9 * sizeof(float))) != cudaSuccess) { ERR(ctx) = std::string("cudaMalloc (rotation staging): ") + cudaGetErrorString(e); return SP_ERR_CUDA; }
                ctx->allocs.push_back(q); ctx->d_dR[i] = (float *)q;
            }
            if ((e = cudaEventCreateWithFlags(&ctx->evd_in[i], cudaEventDisableTiming)) != cudaSuccess || (e = cudaEventCreateWithFlags(&ctx->evd_done[i], cudaEventDisableTiming)) != cudaSuccess) {
                ERR(ctx) = std::string("cudaEventCreate: ") + cudaGetErrorString(e); return SP_ERR_CUDA;
            }
        }
        if ((e = dalloc(ctx, &ctx->d_maps, (size_t)ctx->W + ctx->H)) != cudaSuccess) { ERR(ctx) = std::string("cudaMalloc (maps): ") + cudaGetErrorString(e); return SP_ERR_CUDA; }
    }
    {
        /* K2G::prepareMake3D, k2g.cpp:506-520: (i - c + 0.5) / f evaluated in double, stored as float */
        std::vector<float> maps((size_t)ctx->W + ctx->H);
        for (int i = 0; i < ctx->W; ++i) maps[i] = (float)((i - K->cx + 0.5) / K->fx);
        for (int i = 0; i < ctx->H; ++i) maps[(size_t)ctx->W + i] = (float)((i - K->cy + 0.5) / K->fy);
        CK(cudaMemcpyAsync(ctx->d_maps, maps.data(), maps.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    std::vector<int> cstart, csize;
    chunk_schedule(ctx, nframes, true, cstart, csize);
    const int nchunks = (int)cstart.size();
    auto issue_copy = [&](int ci, int b) -> int {
        const int f0 = cstart[ci], nB = csize[ci];
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->evd_done[b], 0));
        CK(cudaMemcpyAsync(ctx->d_depth[b], (const uint8_t *)depth + N * dsz * f0, N * dsz * nB, cudaMemcpyHostToDevice, ctx->copy_stream));
        if (bgrx) CK(cudaMemcpyAsync(ctx->d_color[b], bgrx + N * 4 * f0, N * 4 * nB, cudaMemcpyHostToDevice, ctx->copy_stream));
        if (R9) CK(cudaMemcpyAsync(ctx->d_dR[b], R9 + 9 * (size_t)f0, sizeof(float) * 9 * nB, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->evd_in[b], ctx->copy_stream));
        return SP_OK;
    };
    /* two lanes like sp_process_frames: chunks alternate between this context and its twin */
    sp_ctx *lanes[4];
    const int nl = lanes_for(ctx, nchunks, lanes);
    struct Pending { cudaEvent_t ev; const sp_frame_result *src; int f0, n; };
    std::vector<Pending> pend;
    size_t drained = 0;
    auto drain = [&](size_t keep) -> int {
        while (pend.size() - drained > keep) {
            const Pending &q = pend[drained++];
            CK(cudaEventSynchronize(q.ev));
            memcpy(results + q.f0, q.src, sizeof(sp_frame_result) * (size_t)q.n);
        }
        return SP_OK;
    };
    const int NBUF = 4;
    auto settle = [&](int rc) -> int {                       /* as in process_impl: no copy stays in flight after an error */
        if (rc != SP_OK) {
            cudaStreamSynchronize(ctx->copy_stream);
            for (int l = 0; l < nl; ++l) cudaStreamSynchronize(lanes[l]->stream);
            (void)cudaGetLastError();
            ctx->last = nullptr;
            for (int l = 0; l < nl; ++l) lanes[l]->last_nB = 0;
        }
        return rc;
    };
    auto run_batch = [&]() -> int {
    for (int ci = 0; ci < std::min(nchunks, NBUF - 1); ++ci) { const int rc = issue_copy(ci, ci); if (rc) return rc; }
    for (int ci = 0; ci < nchunks; ++ci) {
        const int f0 = cstart[ci], nB = csize[ci], buf = ci % NBUF;
        sp_ctx *L = lanes[ci % nl];
        const int hb = (ci / nl) & 1;
        CK(cudaStreamWaitEvent(L->stream, ctx->evd_in[buf], 0));
        if (ci + NBUF - 1 < nchunks) { const int rc = issue_copy(ci + NBUF - 1, (ci + NBUF - 1) % NBUF); if (rc) return rc; }   /* the buffer chunk ci - 1 used */
        k_depth_to_records<<<dim3(std::min(ctx->d.nblk, 148), nB), 256, 0, L->stream>>>(ctx->d_depth[buf], depth_type, bgrx ? ctx->d_color[buf] : nullptr,
                                                                                           ctx->d_maps, ctx->d_maps + ctx->W, ctx->W, ctx->H, mirror ? 1 : 0, ctx->d_din[buf]);
        L->launches += 1;
        CK(cudaMemsetAsync(L->d.results, 0, sizeof(sp_frame_result) * nB, L->stream));
        const int rc = run_chunk(L, ctx->d_din[buf], R9 ? ctx->d_dR[buf] : nullptr, nB, show_mode, -1);
        if (rc) return rc;
        CK(cudaEventRecord(ctx->evd_done[buf], L->stream));
        CK(cudaMemcpyAsync(L->h_results + (size_t)hb * ctx->B, L->d.results, sizeof(sp_frame_result) * nB, cudaMemcpyDeviceToHost, L->stream));
        CK(cudaEventRecord(L->ev_res[hb], L->stream));
        pend.push_back({ L->ev_res[hb], L->h_results + (size_t)hb * ctx->B, f0, nB });
        { const int rd = drain((size_t)nl); if (rd) return rd; }
    }
    { const int rd = drain(0); if (rd) return rd; }
    if (nchunks > 0) {
        for (int l = 0; l < nl; ++l) CK(cudaStreamSynchronize(lanes[l]->stream));
        ctx->last = lanes[(nchunks - 1) % nl];
        collect_times(ctx->last);
    }
    return SP_OK;
    };
    return settle(run_batch());
}

int sp_process_frames(sp_ctx *ctx, const void *xyzrgba, const float *R9, int nframes, int show_mode, sp_frame_result *results)
{
    if (!ctx) return SP_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return process_impl(ctx, xyzrgba, R9, nframes, show_mode, results, true);
}

int sp_process_frames_device(sp_ctx *ctx, const void *d_xyzrgba, const float *d_R9, int nframes, int show_mode, sp_frame_result *results)
{
    if (!ctx) return SP_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return process_impl(ctx, d_xyzrgba, d_R9, nframes, show_mode, results, false);
}

int sp_process_frames_compact(sp_ctx *ctx, const void *xyzrgba, const float *R9, int input_on_device, int nframes, int show_mode,
                              void *d_out, int64_t cap_bytes, int64_t *d_offsets, int64_t *total_bytes)
{
    if (!ctx || !xyzrgba || nframes < 0 || (!d_out && cap_bytes > 0) || cap_bytes < 0) { if (ctx) ERR(ctx) = "bad argument"; return SP_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int rc = process_impl(ctx, xyzrgba, R9, nframes, show_mode, nullptr, input_on_device == 0, true);
    if (rc != SP_OK) return rc;
    if ((size_t)nframes + 1 > ctx->coff_cap) {
        if (ctx->d_coff) cudaFree(ctx->d_coff);
        ctx->d_coff = nullptr; ctx->coff_cap = 0;
        CK(cudaMalloc((void **)&ctx->d_coff, sizeof(long long) * ((size_t)nframes + 1)));
        ctx->coff_cap = (size_t)nframes + 1;
    }
    /* every lane has been synchronised by process_impl: the records are complete */
    k_compact_offsets<<<1, 1024, 0, ctx->stream>>>(ctx->d_batch_res, nframes, ctx->d_coff);
    if (nframes > 0) k_compact_copy<<<nframes, 128, 0, ctx->stream>>>(ctx->d_batch_res, nframes, ctx->d_coff, (unsigned char *)d_out, (long long)cap_bytes);
    ctx->launches += 2;
    if (d_offsets) CK(cudaMemcpyAsync(d_offsets, ctx->d_coff, sizeof(long long) * ((size_t)nframes + 1), cudaMemcpyDeviceToDevice, ctx->stream));
    long long total = 0;
    CK(cudaMemcpyAsync(&total, ctx->d_coff + nframes, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (total_bytes) *total_bytes = (int64_t)total;
    if (total > (long long)cap_bytes) { ERR(ctx) = "compact buffer too small"; return SP_ERR_CAPACITY; }
    return SP_OK;
}

int64_t sp_compact_bytes_max(int nframes) { return (int64_t)nframes * (int64_t)sizeof(sp_frame_result); }

int sp_expand_compact(const void *compact, int64_t bytes, int nframes, sp_frame_result *out)
{
    if ((!compact && bytes > 0) || !out || nframes < 0 || bytes < 0) return SP_ERR_ARG;
    const uint8_t *p = (const uint8_t *)compact, *end = p + bytes;
    for (int f = 0; f < nframes; ++f) {
        if (end - p < SP_COMPACT_HEADER) return SP_ERR_ARG;
        memset(&out[f], 0, sizeof(sp_frame_result));
        memcpy(&out[f], p, SP_COMPACT_HEADER);
        const int np = out[f].n_planes;
        if (np < 0 || np > SP_MAX_PLANES || end - (p + SP_COMPACT_HEADER) < (int64_t)np * (int64_t)sizeof(sp_plane_record)) return SP_ERR_ARG;
        memcpy(out[f].planes, p + SP_COMPACT_HEADER, (size_t)np * sizeof(sp_plane_record));
        p += SP_COMPACT_HEADER + (size_t)np * sizeof(sp_plane_record);
    }
    return p == end ? SP_OK : SP_ERR_ARG;
}

int sp_run_stage_device(sp_ctx *ctx, const void *d_xyzrgba, const float *d_R9, int nframes, int stage)
{
    if (!ctx || !d_xyzrgba || nframes < 1 || nframes > ctx->B || stage < 0 || stage > 2) { if (ctx) ERR(ctx) = "bad argument"; return SP_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    ctx->last = nullptr;
    return run_chunk(ctx, (const float4 *)d_xyzrgba, d_R9, nframes, SP_SHOW_STAIR, stage);
}

void *sp_stream(sp_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int sp_stage_ms(sp_ctx *ctx, float *out, int cap)
{
    if (!ctx || !out) return 0;
    ctx = lane_of_last_chunk(ctx);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    collect_times(ctx);
    const int n = std::min(cap, (int)STG_COUNT);
    for (int i = 0; i < n; ++i) out[i] = ctx->stage_ms[i];
    return n;
}
const char *sp_stage_name(int i) { return (i >= 0 && i < STG_COUNT) ? kStageNames[i] : ""; }

int sp_set_normal_mode(sp_ctx *ctx, int mode)
{
    if (!ctx || (mode != 0 && mode != 1)) { if (ctx) ERR(ctx) = "bad argument"; return SP_ERR_ARG; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    if (mode == 1 && !ctx->knn_pts) {
        const size_t B = ctx->B, N = ctx->N;
        cudaError_t e;
        if ((e = dalloc(ctx, &ctx->knn_mm, B * 8)) != cudaSuccess || (e = dalloc(ctx, &ctx->knn_grid, B)) != cudaSuccess ||
            (e = dalloc(ctx, &ctx->knn_bstart, B << KNN_BITS)) != cudaSuccess || (e = dalloc(ctx, &ctx->knn_bend, B << KNN_BITS)) != cudaSuccess ||
            (e = dalloc(ctx, &ctx->knn_pts, B * N)) != cudaSuccess) {
            ERR(ctx) = std::string("cudaMalloc (kNN workspaces): ") + cudaGetErrorString(e);
            ctx->knn_pts = nullptr;
            return SP_ERR_CUDA;
        }
    }
    ctx->normal_mode = mode;
    for (auto &q : ctx->graphs) cudaGraphExecDestroy(q.exec);
    ctx->graphs.clear();
    return SP_OK;
}

void sp_set_trace(sp_ctx *ctx, int on) { if (ctx) { std::lock_guard<std::mutex> lk(ctx->mu); ctx->trace = on != 0; ctx->tn = 0; } }

int sp_trace_ms(sp_ctx *ctx, int i, const char **name, float *ms)
{
    if (!ctx) return SP_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (i < 0) return std::max(0, ctx->tn - 1);
    if (i + 1 >= ctx->tn || !name || !ms) return SP_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    *name = ctx->tname[i + 1];
    if (cudaEventElapsedTime(ms, ctx->tev[i], ctx->tev[i + 1]) != cudaSuccess) { *ms = 0.f; (void)cudaGetLastError(); }
    return SP_OK;
}
int64_t sp_launch_count(const sp_ctx *ctx) { int64_t n = 0; for (const sp_ctx *t = ctx; t; t = t->twin) n += t->launches; return n; }

/* ------------------------------------------------------------------------------------------------ taps */
static int tap_build(sp_ctx *ctx, int frame, const std::string &name, std::vector<uint8_t> &out)
{
    if (frame < 0 || frame >= ctx->last_nB) { ERR(ctx) = "frame out of range"; return SP_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const SpDev &d = ctx->d;
    const size_t N = ctx->N, f = frame;
    auto d2h = [&](void *dst, const void *src, size_t bytes) { return bytes ? cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) : cudaSuccess; };
    std::vector<int> counts(SP_NCOUNT);
    CK(d2h(counts.data(), d.counts + f * SP_NCOUNT, sizeof(int) * SP_NCOUNT));
    auto put_i = [&](const std::vector<int> &v) { out.resize(v.size() * 4); if (!v.empty()) memcpy(out.data(), v.data(), out.size()); };
    if (name == "xyz" || name == "normals") {
        out.resize(N * 12);
        float *o = (float *)out.data();
        std::vector<float4> p(N);
        CK(d2h(p.data(), (name == "xyz" ? d.xyzw : d.nrm) + f * N, N * 16));
        for (size_t i = 0; i < N; ++i) { o[3 * i] = p[i].x; o[3 * i + 1] = p[i].y; o[3 * i + 2] = p[i].z; }
        return SP_OK;
    }
    if (name == "counts") { put_i(counts); return SP_OK; }
    if (name == "minmax") {
        int mm[8]; CK(d2h(mm, d.mm + f * 8, sizeof(mm)));
        out.resize(24); float *o = (float *)out.data();
        for (int k = 0; k < 6; ++k) o[k] = ord2f(mm[k]);
        return SP_OK;
    }
    if (name == "voxel_key") {
        /* voxel index per pixel, from the frame's sorted (voxel index, pixel) pairs */
        SpFrameGrid g; CK(d2h(&g, d.grid + f, sizeof(g)));
        const size_t M = (size_t)std::max(0, g.M);
        std::vector<unsigned> k(M), v(M);
        const size_t Ns = (N + 3) & ~(size_t)3;
        CK(d2h(k.data(), reinterpret_cast<unsigned *>(g.sorted_in_b ? d.key_b : d.key_a) + f * Ns, M * 4));
        CK(d2h(v.data(), (g.sorted_in_b ? d.val_b : d.val_a) + f * Ns, M * 4));
        out.resize(N * 4); uint32_t *o = (uint32_t *)out.data();
        for (size_t i = 0; i < N; ++i) o[i] = 0xFFFFFFFFu;
        for (size_t i = 0; i < M; ++i) if (v[i] < N) o[v[i]] = k[i];
        return SP_OK;
    }
    if (name == "voxel_idx" || name == "nonan_idx" || name == "noleg_idx" || name == "voxel_code") {
        const int nv = counts[3];
        std::vector<int> idx(nv); std::vector<unsigned char> code(nv);
        CK(d2h(idx.data(), d.voxel_idx + f * N, (size_t)nv * 4)); CK(d2h(code.data(), d.voxel_code + f * N, (size_t)nv));
        if (name == "voxel_code") { out.assign(code.begin(), code.end()); return SP_OK; }
        const int need = name == "voxel_idx" ? SPC_HEAD : (name == "nonan_idx" ? SPC_NONAN : SPC_NOLEG);
        std::vector<int> v;
        for (int i = 0; i < nv; ++i) if (code[i] & need) v.push_back(idx[i]);
        put_i(v); return SP_OK;
    }
    const bool isv = name.size() > 2 && name[0] == 'v' && name[1] == '_';
    const bool ish = name.size() > 2 && name[0] == 'h' && name[1] == '_';
    if (isv || ish) {
        const size_t s = 2 * f + (isv ? 1 : 0);
        const std::string sub = name.substr(2);
        int n = 0; CK(d2h(&n, d.seg_n + s, 4));
        int n_h = 0; CK(d2h(&n_h, d.seg_n + 2 * f, 4));
        const size_t sbase = f * N + (isv ? (size_t)n_h : 0);              /* sp_seg_base: the vertical set follows the horizontal one */
        if (sub == "idx") {
            std::vector<float4> p(n); CK(d2h(p.data(), d.pts + sbase, (size_t)n * 16));
            std::vector<int> v(n);
            for (int i = 0; i < n; ++i) memcpy(&v[i], &p[i].w, 4);
            put_i(v); return SP_OK;
        }
        if (sub == "root") {
            std::vector<int> r(n), cs(n); CK(d2h(r.data(), d.root + sbase, (size_t)n * 4)); CK(d2h(cs.data(), d.csize + sbase, (size_t)n * 4));
            for (int i = 0; i < n; ++i) { const int sz = cs[r[i]]; if (sz < ctx->d.c.min_c || sz > ctx->d.c.max_c) r[i] = -1; }
            put_i(r); return SP_OK;
        }
        if (sub == "log") {
            int ln = 0; CK(d2h(&ln, d.log_n + s, 4)); ln = std::min(ln, SP_LOGCAP);
            std::vector<int> L((size_t)ln * 10); CK(d2h(L.data(), d.logs + s * SP_LOGCAP * 10, L.size() * 4));
            std::vector<int> ord(ln); for (int i = 0; i < ln; ++i) ord[i] = i;
            std::sort(ord.begin(), ord.end(), [&](int a, int b) { return L[a * 10] != L[b * 10] ? L[a * 10] < L[b * 10] : L[a * 10 + 1] < L[b * 10 + 1]; });
            std::vector<int> o; for (int i : ord) o.insert(o.end(), L.begin() + i * 10, L.begin() + i * 10 + 10);
            put_i(o); return SP_OK;
        }
        const bool seg = sub.rfind("seg_", 0) == 0, mrg = sub.rfind("merge_", 0) == 0;
        if (seg || mrg) {
            const std::string what = sub.substr(seg ? 4 : 6);
            int nseg = 0, nm = 0; CK(d2h(&nseg, d.n_seg + s, 4)); CK(d2h(&nm, d.n_merged + s, 4));
            std::vector<SpSegPlane> ch(SP_PSEG); std::vector<int> nx(SP_PSEG); std::vector<SpMergedPlane> mp(SP_PSEG);
            CK(d2h(ch.data(), d.chunks + s * SP_PSEG, sizeof(SpSegPlane) * SP_PSEG)); CK(d2h(nx.data(), d.chunk_next + s * SP_PSEG, 4 * SP_PSEG));
            CK(d2h(mp.data(), d.merged + s * SP_PSEG, sizeof(SpMergedPlane) * SP_PSEG));
            std::vector<int> pp(n); CK(d2h(pp.data(), d.plane_pts + sbase, (size_t)n * 4));
            std::vector<float> coef; std::vector<int> cnt, idx;
            if (seg) {
                for (int i = 0; i < nseg; ++i) { coef.insert(coef.end(), ch[i].coef, ch[i].coef + 4); cnt.push_back(ch[i].count); idx.insert(idx.end(), pp.begin() + ch[i].off, pp.begin() + ch[i].off + ch[i].count); }
            } else {
                for (int i = 0; i < nm; ++i) {
                    coef.insert(coef.end(), mp[i].coef, mp[i].coef + 4); cnt.push_back(mp[i].count);
                    for (int c = mp[i].head; c >= 0; c = nx[c]) idx.insert(idx.end(), pp.begin() + ch[c].off, pp.begin() + ch[c].off + ch[c].count);
                }
            }
            if (what == "coef") { out.resize(coef.size() * 4); if (!coef.empty()) memcpy(out.data(), coef.data(), out.size()); return SP_OK; }
            if (what == "count") { put_i(cnt); return SP_OK; }
            if (what == "idx") { put_i(idx); return SP_OK; }
        }
    }
    if (name == "plane_idx") {
        sp_frame_result r; CK(d2h(&r, ctx->last_res + f, sizeof(r)));
        std::vector<int> v(r.n_plane_points); CK(d2h(v.data(), d.plane_idx + f * N, (size_t)r.n_plane_points * 4));
        put_i(v); return SP_OK;
    }
    ERR(ctx) = "unknown tap: " + name;
    return SP_ERR_ARG;
}

int64_t sp_tap_bytes(sp_ctx *ctx, int frame, const char *name)
{
    if (!ctx || !name) return SP_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    std::vector<uint8_t> v;
    const int rc = tap_build(lane_of_last_chunk(ctx), frame, name, v);
    return rc < 0 ? rc : (int64_t)v.size();
}

int64_t sp_tap_copy(sp_ctx *ctx, int frame, const char *name, void *dst, int64_t cap)
{
    if (!ctx || !name || (!dst && cap > 0)) return SP_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    std::vector<uint8_t> v;
    const int rc = tap_build(lane_of_last_chunk(ctx), frame, name, v);
    if (rc < 0) return rc;
    const int64_t n = std::min<int64_t>(cap, (int64_t)v.size());
    if (n > 0) memcpy(dst, v.data(), (size_t)n);
    return n;
}

int sp_copy_plane_points(sp_ctx *ctx, int frame, int plane, float *xyz, int32_t *pixel_idx, int cap)
{
    if (!ctx) return SP_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx = lane_of_last_chunk(ctx);
    if (frame < 0 || frame >= ctx->last_nB || plane < 0 || cap < 0) { ERR(ctx) = "bad argument"; return SP_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const SpDev &d = ctx->d;
    const size_t N = ctx->N, f = frame;
    sp_frame_result r; CK(cudaMemcpy(&r, ctx->last_res + f, sizeof(r), cudaMemcpyDeviceToHost));
    if (plane >= r.n_planes) { ERR(ctx) = "plane out of range"; return SP_ERR_ARG; }
    const int n = std::min(cap, r.planes[plane].count);
    std::vector<int> idx(n);
    if (n) CK(cudaMemcpy(idx.data(), d.plane_idx + f * N + r.planes[plane].first, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (pixel_idx && n) memcpy(pixel_idx, idx.data(), (size_t)n * 4);
    if (xyz && n) {
        std::vector<float4> p(N);
        CK(cudaMemcpy(p.data(), d.xyzw + f * N, N * 16, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; ++i) { xyz[3 * i] = p[idx[i]].x; xyz[3 * i + 1] = p[idx[i]].y; xyz[3 * i + 2] = p[idx[i]].z; }
    }
    return n;
}

namespace {
/* point access for the stair model: plane clouds stay on the device */
struct DevicePoints : sp_stairmodel::StairPoints {
    sp_ctx *ctx; int frame; const sp_frame_result *fr;
    bool gather(int plane, std::vector<float> &xyz) override
    {
        if (plane < 0 || plane >= fr->n_planes) return false;
        const SpDev &d = ctx->d; const size_t N = ctx->N, f = frame;
        const int n = fr->planes[plane].count;
        std::vector<int> idx((size_t)n);
        if (n && cudaMemcpy(idx.data(), d.plane_idx + f * N + fr->planes[plane].first, (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
        if (hp.empty()) {
            hp.resize(N);
            if (cudaMemcpy(hp.data(), d.xyzw + f * N, N * 16, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
        }
        for (int i = 0; i < n; ++i) { const float4 &p = hp[(size_t)idx[(size_t)i]]; xyz.push_back(p.x); xyz.push_back(p.y); xyz.push_back(p.z); }
        return true;
    }
    bool minmax(int plane, const float *c, const float *dir, float *max_xyz, float *min_xyz) override
    {
        if (plane < 0 || plane >= fr->n_planes || fr->planes[plane].count <= 0) return false;
        const SpDev &d = ctx->d; const size_t N = ctx->N, f = frame;
        const sp_plane_record &pr = fr->planes[plane];
        k_minmax_proj<<<1, 256, 0, ctx->stream>>>(d, frame, pr.first, pr.count, c[0], c[1], c[2], dir[0], dir[1], dir[2], ctx->d_order);
        ctx->launches += 1;
        int ord[2] = { -1, -1 };
        if (cudaMemcpyAsync(ord, ctx->d_order, sizeof(ord), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) return false;
        for (int w = 0; w < 2; ++w) {
            float *xyz = w == 0 ? max_xyz : min_xyz;
            if (ord[w] < 0) return false;
            int pix = -1;
            if (cudaMemcpy(&pix, d.plane_idx + f * N + pr.first + ord[w], 4, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
            if (cudaMemcpy(xyz, d.xyzw + f * N + pix, 12, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
        }
        return true;
    }
    std::vector<float4> hp;
};
} // namespace

int sp_stair_model(sp_ctx *ctx, int frame, sp_stair *out)
{
    if (!ctx) return SP_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx = lane_of_last_chunk(ctx);
    if (!out || frame < 0 || frame >= ctx->last_nB) { ERR(ctx) = "bad argument"; return SP_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<sp_frame_result> fr(1);
    CK(cudaMemcpy(fr.data(), ctx->last_res + frame, sizeof(sp_frame_result), cudaMemcpyDeviceToHost));
    DevicePoints pts; pts.ctx = ctx; pts.frame = frame; pts.fr = fr.data();
    sp_stairmodel::stair_model(ctx->params, fr[0], pts, out);
    return SP_OK;
}

namespace {
/* point access for the batched stair model: the chunk's plane clouds were packed on the device and copied once */
struct HostPoints : sp_stairmodel::StairPoints {
    const sp_frame_result *fr; const float *xyz;               /* this frame's record and its packed plane clouds */
    bool gather(int plane, std::vector<float> &out) override
    {
        if (plane < 0 || plane >= fr->n_planes) return false;
        const sp_plane_record &pr = fr->planes[plane];
        out.insert(out.end(), xyz + 3 * (size_t)pr.first, xyz + 3 * ((size_t)pr.first + (size_t)pr.count));
        return true;
    }
    bool minmax(int plane, const float *c, const float *dir, float *max_xyz, float *min_xyz) override
    {
        if (plane < 0 || plane >= fr->n_planes || fr->planes[plane].count <= 0) return false;
        const sp_plane_record &pr = fr->planes[plane];
        const float *p = xyz + 3 * (size_t)pr.first;
        /* findMinMaxProjPoint (StairPerception.cpp:2780-2808), the arithmetic of k_minmax_proj: strict compares, first occurrence */
        float bmx = -FLT_MAX, bmn = FLT_MAX; int amx = -1, amn = -1;
        for (int i = 0; i < pr.count; ++i) {
            const float px = p[3 * i] - c[0], py = p[3 * i + 1] - c[1], pz = p[3 * i + 2] - c[2];
            const float proj = dir[0] * px + (dir[1] * py + dir[2] * pz);
            if (proj > bmx) { bmx = proj; amx = i; }
            if (proj < bmn) { bmn = proj; amn = i; }
        }
        if (amx < 0 || amn < 0) return false;
        for (int k = 0; k < 3; ++k) { max_xyz[k] = p[3 * amx + k]; min_xyz[k] = p[3 * amn + k]; }
        return true;
    }
};
} // namespace

int sp_stair_model_batch(sp_ctx *ctx, int first_frame, int nframes, sp_stair *out, int nthreads)
{
    if (!ctx) return SP_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    sp_ctx *own = ctx;
    ctx = lane_of_last_chunk(ctx);
    if (!out || first_frame < 0 || nframes < 0 || first_frame + nframes > ctx->last_nB) { ERR(ctx) = "bad argument"; return SP_ERR_ARG; }
    if (nframes == 0) return SP_OK;
    CK(cudaSetDevice(ctx->device));
    const sp_frame_result *dres = ctx->last_res + first_frame;
    SpDev d = ctx->d;
    d.xyzw += (size_t)first_frame * ctx->N; d.plane_idx += (size_t)first_frame * ctx->N;
    /* one scan, one pack kernel, three copies for the whole chunk: offsets, records, packed clouds */
    if ((size_t)nframes + 1 > ctx->coff_cap) {
        if (ctx->d_coff) cudaFree(ctx->d_coff);
        ctx->d_coff = nullptr; ctx->coff_cap = 0;
        CK(cudaMalloc((void **)&ctx->d_coff, sizeof(long long) * ((size_t)nframes + 1)));
        ctx->coff_cap = (size_t)nframes + 1;
    }
    k_plane_cloud_offsets<<<1, 1024, 0, ctx->stream>>>(dres, nframes, ctx->d_coff);
    std::vector<long long> off((size_t)nframes + 1);
    std::vector<sp_frame_result> fr((size_t)nframes);
    CK(cudaMemcpyAsync(off.data(), ctx->d_coff, sizeof(long long) * off.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(fr.data(), dres, sizeof(sp_frame_result) * (size_t)nframes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const long long total = off[(size_t)nframes];
    if ((size_t)total > ctx->cloud_cap) {
        if (ctx->d_cloud) cudaFree(ctx->d_cloud);
        if (ctx->h_cloud) cudaFreeHost(ctx->h_cloud);
        ctx->d_cloud = nullptr; ctx->h_cloud = nullptr; ctx->cloud_cap = 0;
        const size_t cap = (size_t)total + (size_t)total / 4 + 1024;
        CK(cudaMalloc((void **)&ctx->d_cloud, cap * 12));
        CK(cudaMallocHost((void **)&ctx->h_cloud, cap * 12));
        ctx->cloud_cap = cap;
    }
    if (total > 0) {
        k_plane_cloud_pack<<<dim3(8, nframes), 256, 0, ctx->stream>>>(d, dres, ctx->d_coff, ctx->d_cloud, (long long)ctx->cloud_cap);
        CK(cudaMemcpyAsync(ctx->h_cloud, ctx->d_cloud, (size_t)total * 12, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->launches += 2;
    /* the host model, like the reference's (a few tens of planes per frame), one frame per task */
    const int nt = std::max(1, std::min(nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency(), std::min(nframes, 64)));
    const sp_params params = own->params;
    auto work = [&](int t) {
        for (int f = t; f < nframes; f += nt) {
            HostPoints pts; pts.fr = &fr[(size_t)f]; pts.xyz = ctx->h_cloud + 3 * off[(size_t)f];
            sp_stairmodel::stair_model(params, fr[(size_t)f], pts, out + f);
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto &t : th) t.join();
    }
    return SP_OK;
}

/* DProcessor::resolve_result, /root/reference/test/processor.cpp:46-82: range-gated hold-last-good filter on the first step */
void sp_resolve_result(sp_result_def *result, int has_stair, float step_height_m, float step_depth_m, float line_h_m, float line_d_m)
{
    if (!result) return;
    result->has_stair = has_stair ? 1 : 0;
    if (!has_stair) return;
    const float height = step_height_m * 1000, width = step_depth_m * 1000, v_height = line_h_m * 1000, v_depth = line_d_m * 1000;
    if (height > 80 && height < 150) result->height = height;
    if (width > 250 && width < 330) result->width = width;
    if (v_height > 700 && v_height < 1150) result->v_height = v_height;
    if (v_depth > 0 && v_depth < 600) result->v_depth = v_depth;
}

/* DProcessor::pack_and_send_robot, test/processor.cpp:85-100 with the packed union of test/processor.h:75-95:
 * { 0xFF, flag, v_depth, v_height, depth (= width), height, 0xFE }, little-endian floats, 19 bytes */
int sp_pack_robot(const sp_result_def *result, uint8_t *out19)
{
    if (!result || !out19) return SP_ERR_ARG;
    out19[0] = 0xFF;
    out19[1] = (uint8_t)(result->has_stair ? 1 : 0);
    memcpy(out19 + 2, &result->v_depth, 4);
    memcpy(out19 + 6, &result->v_height, 4);
    memcpy(out19 + 10, &result->width, 4);
    memcpy(out19 + 14, &result->height, 4);
    out19[18] = 0xFE;
    return 19;
}

int sp_minmax_proj(sp_ctx *ctx, int frame, int plane, const float *center3, const float *dir3, float *max_xyz, float *min_xyz,
                   int32_t *max_pixel, int32_t *min_pixel)
{
    if (!ctx) return SP_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx = lane_of_last_chunk(ctx);
    if (!dir3 || frame < 0 || frame >= ctx->last_nB || plane < 0) { ERR(ctx) = "bad argument"; return SP_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const SpDev &d = ctx->d;
    const size_t N = ctx->N, f = frame;
    sp_frame_result *r = new sp_frame_result;
    cudaError_t e = cudaMemcpy(r, ctx->last_res + f, sizeof(*r), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess || plane >= r->n_planes) { const bool bad = e == cudaSuccess; delete r; ERR(ctx) = bad ? "plane out of range" : cudaGetErrorString(e); return bad ? SP_ERR_ARG : SP_ERR_CUDA; }
    const sp_plane_record pr = r->planes[plane];
    delete r;
    const float *c = center3 ? center3 : pr.center;
    k_minmax_proj<<<1, 256, 0, ctx->stream>>>(d, frame, pr.first, pr.count, c[0], c[1], c[2], dir3[0], dir3[1], dir3[2], ctx->d_order);
    ctx->launches += 1;
    int ord[2] = { -1, -1 };
    CK(cudaMemcpyAsync(ord, ctx->d_order, sizeof(ord), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const float qnan = std::nanf("");
    for (int w = 0; w < 2; ++w) {
        float *xyz = w == 0 ? max_xyz : min_xyz; int32_t *pixel = w == 0 ? max_pixel : min_pixel;
        int pix = -1;
        if (ord[w] >= 0) CK(cudaMemcpy(&pix, d.plane_idx + f * N + pr.first + ord[w], 4, cudaMemcpyDeviceToHost));
        if (pixel) *pixel = pix;
        if (xyz) {
            xyz[0] = xyz[1] = xyz[2] = qnan;
            if (pix >= 0) CK(cudaMemcpy(xyz, d.xyzw + f * N + pix, 12, cudaMemcpyDeviceToHost));
        }
    }
    return SP_OK;
}

} /* extern "C" */
