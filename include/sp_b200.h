/*
 * sp_b200.h -- C ABI of the B200-native stair-perception segmentation front-end (libsp_b200.so).
 *
 * The reference has no FFI layer: its seam is the C++ class StairDetection in libstairperception.so
 *   ctor     StairDetection(ParameterDef)            /root/reference/modules/stairperception/StairPerception.hpp:144-184
 *   process  bool process(cloud_in, show_mode, stair, vector_plane_sorted, cloud_show, normal_show)
 *                                                    /root/reference/modules/stairperception/StairPerception.hpp:127-132
 *                                                    /root/reference/modules/stairperception/StairPerception.cpp:5-214
 * and, upstream of it, the levelling free functions  /root/reference/test/reader.cpp:228-324.
 * Each entry point below names the reference interface it replaces.  Plain pointers and sizes only; no
 * torch / PCL / Eigen types; never throws; status < 0 = argument or CUDA error (text via sp_last_error).
 * "No stair" / "too few points" are results, not errors (reference: silent `return false`,
 * StairPerception.cpp:24, :56).
 *
 * There is NO CPU path in this library: every entry point that computes runs CUDA kernels on the
 * context's device and fails with SP_ERR_CUDA when no device is available.
 */
#ifndef SP_B200_H_
#define SP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SP_OK 0
#define SP_ERR_ARG (-1)
#define SP_ERR_CUDA (-2)
#define SP_ERR_CAPACITY (-3)

#define SP_MAX_PLANES 64        /* per frame, after merge (H + V) */

/* ShowModeDef, /root/reference/include/common.h:79-92 */
enum sp_show_mode {
    SP_SHOW_ORIGINAL = 0, SP_SHOW_RANGE_LIMITED, SP_SHOW_VOXEL, SP_SHOW_NOLEG, SP_SHOW_HORIZONTAL_CLOUD,
    SP_SHOW_VERTICAL_CLOUD, SP_SHOW_SEG_HORIZONTAL_PLANE, SP_SHOW_SEG_VERTICAL_PLANE,
    SP_SHOW_MERGE_HORIZONTAL_PLANE, SP_SHOW_MERGE_VERTICAL_PLANE, SP_SHOW_STAIR
};

/* ParameterDef, /root/reference/include/common.h:94-131 -- same 35 fields, same order, same types.
 * The library narrows them to float exactly where the StairDetection ctor does (StairPerception.hpp:146-180). */
typedef struct sp_params {
    double x_min, x_max, y_min, y_max, z_min, z_max;
    int normal_compute_points;
    double voxel_x, voxel_y, voxel_z;
    double parallel_angle_diff, perpendicular_angle_diff;
    double cluster_tolerance;
    int min_cluster_size, max_cluster_size;
    double seg_threshold;
    int seg_max_iters, seg_rest_point;
    double seg_plane_angle_diff;
    double merge_threshold, merge_angle_diff;
    int min_num_points;
    double min_length, max_length, min_width, max_width;
    double max_g_height, counter_max_distance;
    double min_height, max_height;
    double vertical_angle_diff, mcv_angle_diff, center_vector_angle_diff, vertical_plane_angle_diff;
    double noleg_distance;
} sp_params;

/* Hand-off record per plane = every field of `Plane` (/root/reference/include/common.h:150-185) except
 * the bulk cloud, which stays on the device as an index list (sp_copy_plane_points). */
typedef struct sp_plane_record {
    int32_t type;            /* Plane::Type: 0 vertical, 1 horizontal */
    int32_t ptype;           /* Plane::Ptype; sortPlanesByXValues sets 3 = others (StairPerception.cpp:1448) */
    int32_t count;           /* points in the plane cloud */
    int32_t first;           /* offset of this plane's points inside the frame's plane point list */
    float coef[4];           /* a, b, c, d */
    float center[3];
    float pcenter[3];
    float evec[9];           /* evec[3*j+k] = component k of eigen_vectors[j] */
    float eval[3];
    float pmin[3], pmax[3];  /* min[a], max[a] */
    float pt_min[3][3], pt_max[3][3];   /* xyz of points_min[a] / points_max[a]; NaN when never set */
    int32_t idx_min[3], idx_max[3];     /* pixel index of those points, -1 when never set */
} sp_plane_record;

/* Fixed-size per-frame result (what a multi-GPU run gathers). */
typedef struct sp_frame_result {
    int32_t status;          /* 0 = ran to the end; 1 = stopped at the "<100 points after crop" guard
                                (StairPerception.cpp:56); 2 = voxel-grid overflow guard
                                (voxel_grid_indices.hpp:295-303); bit 8 set = a device-side capacity was hit */
    int32_t n_valid;         /* rgb != 0 (remove_nan_points, StairPerception.cpp:338) and finite */
    int32_t n_crop;          /* after passThroughCloudANDNormal */
    int32_t n_voxel;         /* occupied voxels */
    int32_t n_nonan;         /* after removeNANNormalPoints */
    int32_t n_noleg;         /* after removeClosetPoints */
    int32_t n_h, n_v;        /* extractPerpendicularPlanePoints */
    int32_t n_clusters_h, n_clusters_v;   /* clusters within [min_cluster_size, max_cluster_size] */
    int32_t n_seg_h, n_seg_v;             /* planes out of segmentHorizontalPlanes / segmentVerticalPlanes */
    int32_t n_planes_h, n_planes_v;       /* after mergeSeperatedPlanes */
    int32_t n_planes;                     /* records filled below (sorted by sortPlanesByXValues) */
    int32_t n_plane_points;               /* total length of the frame's plane point list */
    sp_plane_record planes[SP_MAX_PLANES];
} sp_frame_result;

typedef struct sp_ctx sp_ctx;

/* Replaces StairDetection::StairDetection(ParameterDef) (StairPerception.hpp:144).  Because the reference
 * constructs the object once per frame (test/processor.cpp:156), all device state lives in this context
 * instead: streams, workspaces for `max_batch` frames of width x height points.  device = CUDA ordinal. */
int sp_create(const sp_params *params, int device, int width, int height, int max_batch, sp_ctx **out);
void sp_destroy(sp_ctx *ctx);
const char *sp_last_error(const sp_ctx *ctx);   /* ctx may be NULL: returns the last create error */

/* Replaces DProcessor::update_parameters -> new StairDetection(parameters) (test/processor.cpp:102-109,156). */
int sp_set_params(sp_ctx *ctx, const sp_params *params);

/* Replaces computeRotationMatrix (test/reader.cpp:228-249 Euler, :251-271 quaternion).  Host-side, 9 floats,
 * row-major R = m3 * R_imu * m1. */
void sp_rotation_from_euler(float pitch_deg, float roll_deg, float yaw_deg, float *R9);
void sp_rotation_from_quat(float w, float x, float y, float z, float *R9);

/* Replaces rotationCloudPoints (test/reader.cpp:273-324) + StairDetection::process up to the hand-off to
 * stairModel (StairPerception.cpp:5-188) for `nframes` independent frames.
 *   xyzrgba : nframes x (width*height) x 16 B records {float x, y, z; uint32 rgba} (the PCD record layout)
 *   R9      : nframes x 9 row-major levelling rotations, or NULL = already levelled (files mode, reader.cpp:493)
 *   results : nframes records, caller-allocated
 * sp_process_frames takes HOST pointers (copies inside); sp_process_frames_device takes DEVICE pointers to
 * inputs already resident in HBM and still writes `results` to host memory.
 * ORDERING PRECONDITION of the *_device entry points (also sp_run_stage_device): the library launches on its own
 * non-blocking streams and does not know the stream that produced d_xyzrgba / d_R9.  The caller must make those
 * buffers complete before the call (cudaStreamSynchronize / cudaEventSynchronize on the producer) and must not
 * overwrite them until the call returns.  The Python binding synchronises torch's current stream for CUDA tensors.  nframes may exceed max_batch:
 * the call loops over chunks, and from the second chunk on alternates between two lanes: the context's own workspace and
 * a twin (same size, created on first use on the same device, own stream), so that one chunk's latency-bound tail overlaps
 * the next chunk's streaming head.  The environment variable SP_NO_TWIN=1 (read at sp_create) keeps a single lane and a
 * single workspace; SP_LANES=n (1..4, default 2) sets the lane count (three lanes measured the same as two, four slower).  Results, taps and the stair model do not depend on the lane count.  show_mode: SP_SHOW_*
 * (SP_SHOW_ORIGINAL does no work, StairPerception.cpp:24; the other taps stop the pipeline where process() returns early). */
int sp_process_frames(sp_ctx *ctx, const void *xyzrgba, const float *R9, int nframes, int show_mode,
                      sp_frame_result *results);
int sp_process_frames_device(sp_ctx *ctx, const void *d_xyzrgba, const float *d_R9, int nframes, int show_mode,
                             sp_frame_result *results);

/* Compact records -- what a multi-GPU run exchanges and what a consumer of many frames reads: per frame the 16 leading
 * int32 counters of sp_frame_result (64 bytes) followed by its n_planes sp_plane_record, frames back to back (about 3 KB per
 * frame instead of 14 400 B).  Same processing as sp_process_frames(_device); the records never pass through host memory:
 *   input_on_device : 0 = xyzrgba / R9 are host pointers, 1 = device pointers (same ordering precondition as above)
 *   d_out           : DEVICE buffer of cap_bytes (sp_compact_bytes_max(nframes) always suffices)
 *   d_offsets       : DEVICE array of nframes + 1 int64 byte offsets of the frames inside d_out, or NULL
 *   total_bytes     : host, bytes written (also set when the call fails with SP_ERR_CAPACITY)
 * sp_expand_compact (host only, no device work) turns a compact buffer back into fixed-size records.
 * Replaces, for a batch, what DProcessor keeps of a frame: `vector_plane_sorted` without the clouds
 * (/root/reference/test/processor.cpp:173-196). */
int sp_process_frames_compact(sp_ctx *ctx, const void *xyzrgba, const float *R9, int input_on_device, int nframes, int show_mode,
                              void *d_out, int64_t cap_bytes, int64_t *d_offsets, int64_t *total_bytes);
int64_t sp_compact_bytes_max(int nframes);
int sp_expand_compact(const void *compact, int64_t bytes, int nframes, sp_frame_result *out);

/* Depth-image ingest -- the step BEFORE the path, so that 2-8 bytes per pixel cross PCIe instead of 16:
 * replaces K2G::updateCloud (/root/reference/modules/k2g/k2g.cpp:231-308, pinhole maps of K2G::prepareMake3D :506-520)
 * followed by everything sp_process_frames does.
 *   depth      : nframes x width*height undistorted depth in millimetres; depth_type 0 = float32 (libfreenect2's undistorted
 *                frame), 1 = uint16 (the depth PNG of Reader::save_images, test/reader.cpp:128)
 *   bgrx       : registered colour, 4 bytes per pixel (b, g, r, unused), or NULL = every valid pixel white.  The front-end's
 *                validity mask is "r, g and b all non-zero" (StairPerception.cpp:338-356), so colour matters.
 *   K          : IR camera intrinsics (libfreenect2 IrCameraParams fx, fy, cx, cy)
 *   mirror     : the horizontal flip of K2G(true) (k2g.cpp:247-253)
 * The frame loaded after the call (taps, sp_copy_plane_points) is indexed like the cloud: pixel = row * width + column. */
typedef struct sp_intrinsics { float fx, fy, cx, cy; } sp_intrinsics;
int sp_process_depth_frames(sp_ctx *ctx, const void *depth, int depth_type, const uint8_t *bgrx, const sp_intrinsics *K, int mirror,
                            const float *R9, int nframes, int show_mode, sp_frame_result *results);

/* Stage taps of the LAST chunk processed (the ShowModeDef taps of process(), StairPerception.cpp:49-176, plus
 * the intermediate arrays the parity tests compare).  `frame` indexes the last chunk.  Names:
 *   "xyz" (N x 3 float, levelled+masked)  "normals" (N x 3 float)  "minmax" (6 float)  "voxel_key" (N u32)
 *   "voxel_idx" "h_idx" "v_idx" (int32 pixel indices)   "h_root" "v_root" (int32 per H/V point: smallest
 *   member position of its cluster)   "h_seg_coef" "v_seg_coef" (P x 4 float)  "h_seg_count" "v_seg_count"
 *   "h_seg_idx" "v_seg_idx" (int32 pixel indices, planes concatenated)   "h_log" "v_log" (int32 x 10 per RANSAC
 *   call: cluster, call, n, scored hypotheses, best sample i0 i1 i2, best count, inliers, emitted)
 *   "h_merge_coef/count/idx" "v_merge_coef/count/idx" (after mergeSeperatedPlanes)
 *   "plane_idx" (int32 pixel indices of the final planes, concatenated)
 * Returns bytes (sp_tap_bytes), bytes copied (sp_tap_copy) or <0 on error / unknown name. */
int64_t sp_tap_bytes(sp_ctx *ctx, int frame, const char *name);
int64_t sp_tap_copy(sp_ctx *ctx, int frame, const char *name, void *dst, int64_t cap);

/* Replaces reading Plane::cloud (common.h:152): xyz (3 floats / point) and pixel indices of plane `plane`
 * of frame `frame` of the last chunk, in the reference's point order.  Either output may be NULL. */
int sp_copy_plane_points(sp_ctx *ctx, int frame, int plane, float *xyz, int32_t *pixel_idx, int cap);

/* Replaces StairDetection::findMinMaxProjPoint (StairPerception.cpp:2780-2808), which the stair model calls on hand-off
 * planes (:2680, :2695, :2698, :2888, :2951, :2986, :3053): the points of plane `plane` (frame `frame` of the last chunk)
 * whose projection of (p - centre) on `dir3` is largest / smallest, reduced on the device so the plane's cloud never has
 * to travel.  center3 = NULL uses the plane's own centre.  Outputs may be NULL; xyz = NaN and pixel = -1 for an empty plane. */
int sp_minmax_proj(sp_ctx *ctx, int frame, int plane, const float *center3, const float *dir3, float *max_xyz, float *min_xyz,
                   int32_t *max_pixel, int32_t *min_pixel);

/* The consumer of the hand-off records: StairDetection::stairModel (StairPerception.cpp:1621-2773) restated without PCL,
 * run on the host like the reference's (a few tens of planes), fed from the per-plane records of frame `frame` of the last
 * chunk; the plane clouds stay on the device (findMinMaxProjPoint runs there as a reduction; only a merged ground plane
 * pulls its points).  Returns SP_OK and fills *out: has_stair (what process() returns, StairPerception.cpp:190), the list
 * nodes in order -- kind 0 = concave line, 1 = step with {height, depth, line.h, line.d} = the four columns of
 * printStairEstParam (:265-290) -- and the final Plane::Ptype of every hand-off plane.  plane_h / plane_v index the frame's
 * hand-off planes (-1 = merged ground plane, -2 - k = virtual riser k of a tread-to-tread pair).
 * When a stair is found the model goes on like process() does (:192-193): the two side ("slide") planes of stairModel's step 8
 * (:2562-2612, a sequential if / else-if extreme scan over the stair planes' points along the section normal) and
 * computePlaneCounter (:2861-3075): the four contour points of every plane of the stair (what the GUI draws with addPolygon,
 * modules/gui/gui.cpp:326-330) as intersections of the step lines with the side planes.  contour_width = the reference's
 * counter.width (0 = untouched, 2 = front edge only, 4 = closed); points the reference never writes stay (0, 0, 0) like PCL's
 * default-constructed points.  A virtual riser has no cloud: its lower edge passes through (0, 0, 0) in the reference as well
 * (findMinMaxProjPoint on an empty cloud). */
#define SP_MAX_STAIR_NODES 64
#define SP_MAX_VIRTUAL_PLANES 32
typedef struct sp_stair_node {
    int32_t kind, count, plane_h, plane_v;
    float line_h, line_d, line_coeff[6];
    double height, depth;
} sp_stair_node;
typedef struct sp_stair {
    int32_t has_stair, n_steps, n_nodes;
    int32_t ptype[SP_MAX_PLANES];
    float vnormal[3], snormal[3];
    int32_t n_virtual;
    sp_stair_node nodes[SP_MAX_STAIR_NODES];
    float slide_left[4], slide_right[4];                     /* plane coefficients A, B, C, D */
    float contour[SP_MAX_PLANES][4][3];                      /* per hand-off plane */
    int32_t contour_width[SP_MAX_PLANES];
    float ground_contour[4][3];                              /* the merged ground plane (plane index -1) */
    int32_t ground_contour_width;
    int32_t virtual_contour_width[SP_MAX_VIRTUAL_PLANES];
    float virtual_contour[SP_MAX_VIRTUAL_PLANES][4][3];      /* virtual riser k (plane index -2 - k) */
} sp_stair;
int sp_stair_model(sp_ctx *ctx, int frame, sp_stair *out);
/* The same consumer for many frames at once -- frames [first_frame, first_frame + nframes) of the last chunk: the plane clouds of
 * all of them are packed by one kernel and cross PCIe in one copy (with the records and the per-frame offsets: three copies per
 * chunk instead of several blocking ones per findMinMaxProjPoint call), then the host model runs one frame per task on
 * `nthreads` host threads (0 = all).  out[i] is bit-identical to sp_stair_model(ctx, first_frame + i, ...). */
int sp_stair_model_batch(sp_ctx *ctx, int first_frame, int nframes, sp_stair *out, int nthreads);

/* The step AFTER the path (host-side, no device work): the driver's hold-last-good filter and the 19-byte robot packet.
 * sp_result_def = ResultDef (/root/reference/include/common.h:133-141).  sp_resolve_result replaces DProcessor::resolve_result
 * (/root/reference/test/processor.cpp:46-82): inputs are the first step's height, depth, line.h and line.d in metres as the
 * stair model reports them; each field of *result is overwritten only when the new value passes its range gate.
 * sp_pack_robot replaces DProcessor::pack_and_send_robot (test/processor.cpp:85-100, test/processor.h:75-95). */
typedef struct sp_result_def { int32_t has_stair; float height, width, v_height, v_depth; double time; } sp_result_def;
void sp_resolve_result(sp_result_def *result, int has_stair, float step_height_m, float step_depth_m, float line_h_m, float line_d_m);
int sp_pack_robot(const sp_result_def *result, uint8_t *out19);

/* Timing of the last chunk: CUDA-event milliseconds per stage; returns number of stages written. */
int sp_stage_ms(sp_ctx *ctx, float *out, int cap);
const char *sp_stage_name(int i);
/* number of kernels this library launched since sp_create (its own + CUB) */
int64_t sp_launch_count(const sp_ctx *ctx);

/* Normal estimation mode.  0 (default) = organised-image normals, the live path (normalEstimationIntegral,
 * StairPerception.cpp:579-592).  1 = k-nearest-neighbour covariance normals for unorganised clouds (normalEstimationOMP,
 * StairPerception.cpp:524-541) with k = ParameterDef::normal_compute_points (<= 128); create the context with
 * width = number of points, height = 1.  Query = search surface = every finite point of the frame. */
int sp_set_normal_mode(sp_ctx *ctx, int mode);

/* Per-kernel trace of the last chunk (the reference's TIME_RECORD per-stage timers, StairPerception.hpp:13,181-190,
 * at kernel granularity).  Off by default; sp_set_trace(ctx, 1) or the environment variable SP_TRACE=1 makes the
 * library record a CUDA event after every launch.  sp_trace_ms(ctx, -1, 0, 0) returns the number of entries;
 * sp_trace_ms(ctx, i, &name, &ms) the i-th entry's kernel name and duration. */
void sp_set_trace(sp_ctx *ctx, int on);
int sp_trace_ms(sp_ctx *ctx, int i, const char **name, float *ms);

/* Run only one stage on device-resident inputs (bench / profiling of a single kernel family).
 * stage: 0 = ingest+normals, 1 = voxel (key+sort+pick), 2 = whole pipeline w/o D2H. */
int sp_run_stage_device(sp_ctx *ctx, const void *d_xyzrgba, const float *d_R9, int nframes, int stage);

/* CUDA stream handle (cudaStream_t) the context launches on, for callers that time with events. */
void *sp_stream(sp_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SP_B200_H_ */
