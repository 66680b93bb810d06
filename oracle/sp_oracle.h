/*
 * sp_oracle.h -- C API of the CPU ORACLE for the stair-perception segmentation front-end.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * The oracle is a plain C++ restatement (no PCL / Eigen / Boost -- none exist in this image) of the
 * reference's per-frame path StairDetection::process()
 *   (/root/reference/modules/stairperception/StairPerception.cpp:5-214) and of the PCL algorithms it
 * calls.  PARITY UNPINNED at stage level: the reference ships no test, golden vector or fixture for
 * any intermediate stage, and PCL (un-vendored, version un-pinned) cannot be built here.  The only
 * external anchor is the published step table for samples/0001 (pic/screenshot2.png), which
 * tests/test_oracle_anchor.py checks at mm level.
 */
#ifndef SP_ORACLE_H_
#define SP_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field order and types as ParameterDef, /root/reference/include/common.h:94-131. */
typedef struct orc_params {
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
} orc_params;

/* Per-plane hand-off record = everything in `Plane` (common.h:150-185) except the bulk cloud.
 * Same layout as sp_plane_record in include/sp_b200.h (kept textually separate on purpose). */
typedef struct orc_plane_record {
    int32_t type;            /* Plane::Type: 0 vertical, 1 horizontal */
    int32_t ptype;           /* Plane::Ptype as set by sortPlanesByXValues: 3 = others */
    int32_t count;           /* points in the plane cloud */
    int32_t first;           /* offset of this plane's point list inside the "plane_idx" tap */
    float coef[4];
    float center[3];
    float pcenter[3];
    float evec[9];           /* evec[3*j+k] = component k of eigen_vectors[j] */
    float eval[3];
    float pmin[3], pmax[3];  /* min[a], max[a] */
    float pt_min[3][3], pt_max[3][3];   /* xyz of points_min[a], points_max[a] (NaN if never set) */
    int32_t idx_min[3], idx_max[3];     /* original pixel index of those points, -1 if never set */
} orc_plane_record;

typedef struct orc_ctx orc_ctx;

orc_ctx *orc_create(const orc_params *p);
void orc_destroy(orc_ctx *c);

/* Arithmetic modes.  name/value:
 *   "normal_arith": 0 = PCL integral images with double accumulators (faithful restatement),
 *                   1 = "box32": direct k x k box sums in float32 in the fixed order documented in DESIGN.md
 *                       (the portable definition the CUDA kernel reproduces bit for bit).  Default 1.
 *   "voxel_order":  0 = stable (ascending input index inside a voxel; the documented rule), default;
 *                   1 = libstdc++ std::sort on the key only (what the reference binary would do with g++).
 *   "eig_math":     0 = portable det_* math from include/sp_detmath.h, the header the product compiles too (default);
 *                   1 = the same restatements on libm trig; 2 = the oracle's own numerics (oracle/indep_math.h: double
 *                   Jacobi eigen-solver, libm acos / log) -- shares nothing with the product.
 * returns 0, or -1 for an unknown name. */
int orc_set_mode(orc_ctx *c, const char *name, int value);

/* normal_mode: 0 = organised integral-image normals (live path, StairPerception.cpp:40),
 *              1 = kNN covariance normals (normalEstimationOMP, :524-541; unorganised clouds).
 * R9: row-major 3x3 levelling rotation or NULL (cloud already levelled, files mode test/reader.cpp:493).
 * keep_taps: 0 = only the final result (timed baseline), 1 = store every stage tap.
 * returns has_planes (number of planes handed to the stair model, >= 0), or <0 on argument error. */
int orc_process(orc_ctx *c, const void *xyzrgba, int width, int height, const float *R9,
                int show_mode, int normal_mode, int keep_taps);

/* Run nframes independent frames on nthreads host threads (one private context each); returns
 * total number of planes found, fills per-frame plane counts if plane_counts != NULL. */
int64_t orc_process_batch(const orc_params *p, const void *xyzrgba, int width, int height,
                          const float *R9 /* nframes x 9 or NULL */, int nframes, int nthreads,
                          int normal_arith, int32_t *plane_counts);

/* Stage taps: raw little-endian arrays by name (see oracle/README.md for the list). */
int64_t orc_tap_bytes(orc_ctx *c, const char *name);          /* -1 if absent */
int64_t orc_tap_copy(orc_ctx *c, const char *name, void *dst, int64_t cap);
/* names of all taps, '\n'-separated, into dst; returns bytes needed */
int64_t orc_tap_names(orc_ctx *c, char *dst, int64_t cap);

/* Audit of the last orc_process(keep_taps = 1) call in eig_math 0 / 1: every eigen-solve, acos and log re-done with the
 * oracle's own numerics (oracle/indep_math.h).  out = { n_eigen33, n_jacobi, n_acos, n_log, eigen33 max angle [rad],
 * the same over well-separated spectra, eigen33 max |lambda0 - ref| / lambda_max, jacobi max |eval - ref| / lambda_max,
 * jacobi max eigenvector angle, acos max abs error, log max relative error }.  Returns the number written. */
int orc_audit(orc_ctx *c, double *out, int cap);

/* per-stage wall time of the last orc_process call, milliseconds; returns number written */
int orc_stage_ms(orc_ctx *c, double *out, int cap);
const char *orc_stage_name(int i);

/* Levelling matrices, test/reader.cpp:228-271. */
void orc_rotation_from_euler(float pitch_deg, float roll_deg, float yaw_deg, float *R9);
void orc_rotation_from_quat(float w, float x, float y, float z, float *R9);

/* PRNG restatements (Appendix E known answers). */
void orc_mt19937_draws(uint32_t seed, int n, int32_t *out);   /* boost uniform_int(0,INT_MAX) = mt()>>1 */
void orc_glibc_rand(uint32_t seed, int n, int32_t *out);      /* glibc TYPE_3 rand() */

/* stand-alone stage entry points used by unit tests */
/* eigen33 smallest eigenpair of a symmetric 3x3 float matrix (row-major 9); libm = 1 uses libm trig */
void orc_eigen33(const float *cov9, int libm, float *eigenvalue, float *eigenvector3);
/* symmetric 3x3 Jacobi eigen decomposition (descending), vectors as columns, row-major 9 */
void orc_jacobi3(const float *cov9, float *evals3, float *evecs9);

/* K2G::updateCloud (k2g.cpp:231-308): depth image (mm; type 0 float32, 1 uint16) + registered colour (b,g,r,x or NULL) -> records */
void orc_depth_to_cloud(const void *depth, int depth_type, const uint8_t *bgrx, int W, int H, float fx, float fy, float cx, float cy,
                        int mirror, void *out_records);

/* findMinMaxProjPoint (StairPerception.cpp:2780-2808) on plane `plane` of the last processed frame; center3 NULL = plane centre */
int orc_minmax_proj(orc_ctx *c, int plane, const float *center3, const float *dir3, float *max_xyz, float *min_xyz,
                    int32_t *max_pixel, int32_t *min_pixel);

#ifdef __cplusplus
}
#endif
#endif
