/*
 * sp_common.cuh -- device-side data layout of the B200 segmentation front-end.
 *
 * One context processes chunks of up to `B` independent frames of N = W*H points.  Everything a frame needs
 * lives in HBM in structure-of-arrays form, frame-major:
 *   xyzw, nrm             [B][N] float4    levelled + masked points {x, y, z, 0} and normals {nx, ny, nz, 0} (pixel order)
 *   key/val (x2)          [B*N]  u64/u32   (frame << 32 | voxel index) and pixel index, before/after the radix sort
 *   rep, code             [B][N]           per sorted position: voxel representative pixel and its class bits
 *   "segments"            s = 2*frame + {0: horizontal set, 1: vertical set}; the two segments of a frame share a slice of N entries
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>
#include "../../include/sp_b200.h"
#include "../../include/sp_detmath.h"

#define SP_CMAX 128          /* clusters per segment that can produce planes (size >= seg_rest_point) */
#define SP_PSEG 64           /* plane slots per segment out of RANSAC */
#define SP_LOGCAP 256        /* RANSAC call-log rows per segment */
#define SP_MT_LEN 16384      /* pre-computed mt19937(12345) draws */
#define SP_RAND_LEN 16384    /* pre-computed glibc rand() draws */
#define SP_ROUND 8           /* hypotheses scored per sweep over a cluster */
#define SP_HASH_BITS 16      /* radius-graph hash grid: buckets per segment */
#define SP_NCOUNT 16

/* code bits per sorted position (voxel stage) */
#define SPC_HEAD 1
#define SPC_NONAN 2
#define SPC_NOLEG 4
#define SPC_H 8
#define SPC_V 16

struct SpFrameGrid {           /* voxel grid of one frame, voxel_grid_indices.hpp:292-318 */
    int minb[3];
    int mul[3];
    int status;                /* 0 ok, 1 = <100 points after crop, 2 = overflow guard */
    int cells;                 /* dx * dy * dz: every voxel index of the frame is below it */
    int sorted_in_b;           /* which of the two (key, pixel) buffers holds the frame's sorted pairs (odd number of sort passes: b) */
    int npass;                 /* 9-bit radix passes the frame's voxel indices need (0 when process() stops before the voxel filter) */
    int M;                     /* cropped points of the frame = pairs to sort (0 when process() stops before the voxel filter) */
};

struct SpSegPlane {            /* one plane out of RANSAC (slot order = cluster rank, then call order) */
    float coef[4];
    int count;                 /* 0 = unused slot */
    int off;                   /* offset of its pixel list in plane_pts[seg] */
    int cluster, call;
};

struct SpMergedPlane {         /* after mergeSeperatedPlanes: a chain of SpSegPlane chunks */
    float coef[4];
    int count;
    int head;                  /* first chunk (slot index), chunks linked through chunk_next */
    int pad0, pad1;
};

struct SpConst {               /* derived parameters, narrowed exactly where the reference narrows them */
    float x0, x1, y0, y1, z0, z1;        /* crop limits (float members, StairPerception.hpp:32-37) */
    float ivx, ivy, ivz;                 /* 1.0f / leaf (voxel_grid_indices.h:131) */
    float noleg_th2;                     /* th*th, th = (float)noleg_distance */
    float h_ge, h_le, v_lo, v_hi;        /* a8 as float cut-points on nx (see sp_b200.cu: make_cutpoints) */
    float r2, cell, inv_cell;            /* clustering: cell edge = tolerance x 1.0001, points are binned by x * (1 / edge) */
    int min_c, max_c;
    float seg_thr;
    double sin_angle;
    int max_iters, rest;
    double merge_ath;
    float merge_thr;
    double log_prob;                     /* det_log(1 - 0.99) */
};

struct SpDev {
    int W, H, N, B;
    SpConst c;
    /* input */
    const float4 *in;          /* [B][N] {x, y, z, rgba bits} */
    const float *R;            /* [B][9] or NULL */
    /* stage 1 */
    float4 *xyzw;              /* [B][N] levelled + masked points {x, y, z, 0} (pixel order): one 32-byte sector per gathered point */
    float4 *nrm;               /* [B][N] normals {nx, ny, nz, 0} */
    int *mm;                   /* [B][8] ordered-int min xyz, max xyz, + 2 pad */
    int *counts;               /* [B][SP_NCOUNT] */
    SpFrameGrid *grid;
    int *vcnt;                 /* [B][nblk] cropped pixels per 1024-pixel block (K1K2 counts them, the voxel stage compacts by them) */
    int *voff;                 /* [B][nblk] exclusive scan of vcnt inside the frame */
    unsigned *vth;             /* [B][ntile][512] sort: per (tile, digit) counts, then first destinations */
    int ntile;                 /* sort tiles per frame slice */
    int *rep;                  /* [B][N] per sorted position: the voxel's representative pixel (run heads) */
    unsigned char *code;       /* [B][N] per sorted position: class bits (0 = not a run head) */
    int *blk;                  /* [B][nblk][5] per-1024-block counts: head, nonan, noleg, H, V */
    /* voxel */
    unsigned long long *key_a, *key_b;   /* voxel stage: as u32, a slice of (N + 3) & ~3 per frame: the frame's cropped points (voxel index, pixel) before / after a sort pass; the kNN grid build uses them as u64 */
    unsigned *val_a, *val_b;
    int nblk;
    int *voxel_idx;            /* [B][N] */
    unsigned char *voxel_code; /* [B][N] */
    /* segments */
    float4 *pts;               /* [B][N] {x, y, z, pixel}: the horizontal set of a frame, then its vertical set (sp_seg_base) */
    int *seg_n;                /* [2B] */
    int *ccnt;                 /* [2B][1 << SP_HASH_BITS] points per cell bucket (zero between chunks) */
    int *cstart;               /* [2B][(1 << SP_HASH_BITS) + 1] exclusive scan of ccnt */
    float4 *spts;              /* [2B][N] points in bucket-major order {x, y, z, list position} */
    int *parent, *root, *csize;   /* [2B][N] */
    int *n_clusters;           /* [2B] clusters within [min,max] */
    int *n_cand;               /* [2B] */
    int *cand_root, *cand_size, *cand_off, *cand_slot;   /* [2B][SP_CMAX] */
    float *wx, *wy, *wz; int *wpix;      /* [2B][N] cluster workspaces: points cluster by cluster (k_cl_rank) */
    float *vx, *vy, *vz; int *vpix;      /* [2B][N] twin workspaces (k_ransac ping-pongs the shrinking remainder) */
    int *plane_pts;            /* [2B][N] pixel lists of the planes */
    float *plane_x, *plane_y, *plane_z;  /* [2B][N] their coordinates, same order */
    SpSegPlane *seg_planes;    /* [2B][SP_PSEG] */
    int *logs, *log_n;         /* [2B][SP_LOGCAP][10], [2B] */
    /* merge + stats */
    int *n_seg;                /* [2B] planes out of RANSAC (compacted) */
    SpSegPlane *chunks;        /* [2B][SP_PSEG] compacted */
    int *chunk_next;           /* [2B][SP_PSEG] */
    SpMergedPlane *merged;     /* [2B][SP_PSEG] */
    int *n_merged;             /* [2B] */
    sp_plane_record *info;     /* [2B][SP_PSEG] */
    /* output */
    sp_frame_result *results;  /* [B] */
    int *plane_idx;            /* [B][N] final plane point lists */
    /* tables */
    const int *mt;             /* SP_MT_LEN draws of mt19937(12345) >> 1 */
    const int *rnd;            /* SP_RAND_LEN draws of glibc rand() from seed 1 */
    int *flags;                /* [B] capacity flags */
};

/* A frame's horizontal (s = 2f) and vertical (s = 2f + 1) point sets share ONE slice of N entries of every per-segment array:
 * both are subsets of the frame's voxel representatives, so together they never exceed N; the vertical set starts where the
 * horizontal one ends (seg_n is final once k_scan_blocks has run). */
__device__ __forceinline__ size_t sp_seg_base(const SpDev &d, int s) { return (size_t)(s >> 1) * d.N + ((s & 1) ? (size_t)d.seg_n[s - 1] : 0); }

__device__ __forceinline__ int sp_f2ord(float f) { int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }
__device__ __forceinline__ float sp_ord2f(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }
__device__ __forceinline__ bool sp_finite(float f) { return (__float_as_uint(f) & 0x7f800000u) != 0x7f800000u; }
__device__ __forceinline__ bool sp_finite3(float a, float b, float c) { return sp_finite(a) && sp_finite(b) && sp_finite(c); }

/* Eigen 4-float packet dot as used by pcl::SampleConsensusModelPlane: (a0 + a1) + (a2 + a3) */
__device__ __forceinline__ float sp_dot4(float a, float b, float c, float d, float x, float y, float z, float w)
{
    return (a * x + b * y) + (c * z + d * w);
}

/* block-wide exclusive scan of a 0/1 flag; blockDim.x must be a multiple of 32 and <= 1024.
 * `warp_sums` is 33 ints of shared memory.  Returns the exclusive prefix; *total = block total. */
__device__ __forceinline__ int sp_block_scan_flag(bool flag, int *warp_sums, int *total)
{
    const unsigned ballot = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int within = __popc(ballot & ((1u << lane) - 1u));
    __syncthreads();
    if (lane == 0) warp_sums[wid] = __popc(ballot);
    __syncthreads();
    if (wid == 0) {
        int v = lane < nw ? warp_sums[lane] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane < nw) warp_sums[lane] = inc - v;
        if (lane == 31) warp_sums[32] = inc;
    }
    __syncthreads();
    *total = warp_sums[32];
    return warp_sums[wid] + within;
}

/* the same for three flags at once, packed in 10-bit fields (a | b << 10 | c << 20): blockDim.x <= 1023 keeps every count in its field */
__device__ __forceinline__ int sp_block_scan_flags3(bool fa, bool fb, bool fc, int *warp_sums, int *total)
{
    const unsigned ba = __ballot_sync(0xffffffffu, fa), bb = __ballot_sync(0xffffffffu, fb), bc = __ballot_sync(0xffffffffu, fc);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int within = __popc(ba & lt) | (__popc(bb & lt) << 10) | (__popc(bc & lt) << 20);
    __syncthreads();
    if (lane == 0) warp_sums[wid] = __popc(ba) | (__popc(bb) << 10) | (__popc(bc) << 20);
    __syncthreads();
    if (wid == 0) {
        int v = lane < nw ? warp_sums[lane] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane < nw) warp_sums[lane] = inc - v;
        if (lane == 31) warp_sums[32] = inc;
    }
    __syncthreads();
    *total = warp_sums[32];
    return warp_sums[wid] + within;
}
