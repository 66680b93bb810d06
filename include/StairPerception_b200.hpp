/*
 * StairPerception_b200.hpp -- header-only C++ host facade over the C ABI of libsp_b200.so (include/sp_b200.h).
 *
 * It reproduces the public surface of the reference's class
 *     StairDetection(ParameterDef)                       /root/reference/modules/stairperception/StairPerception.hpp:144-184
 *     bool process(cloud_in, show_mode, stair, vector_plane_sorted, cloud_show, normal_show)      ...:127-132, .cpp:5-214
 * with the same type and field names (ParameterDef, ShowModeDef, Plane -- /root/reference/include/common.h:79-185) over
 * POD stand-ins for the PCL containers, because PCL does not exist in this image.  The stand-ins have PCL's member
 * names (`points`, `width`, `height`, `is_dense`, `values`, `x/y/z/rgba`, `normal_x/..`), so the body compiles
 * unchanged against the reference's own types: define SP_B200_REFERENCE_TYPES and include the reference's
 * include/common.h first (PCL is absent from this image, so that variant is not compiled here).
 *
 * The stair model: `stairModel()` / `computePlaneCounter()` (StairPerception.cpp:1621-2773, :2861-3075) are the unchanged
 * consumer of `vector_plane_sorted`; in a drop-in the reference keeps calling its own (INTEGRATION.md shows the patch) and
 * process() returns "planes were handed over".  Stand-alone callers pass `stair_out`: process() then also runs the
 * library's restatement of stairModel (sp_stair_model) and returns "stair found" like the reference.
 *
 * The reference constructs a StairDetection per frame (test/processor.cpp:156); all device state therefore lives in
 * a process-level context per (device, width, height), created on first use and shared by every object.
 * No CPU path: without a CUDA device the constructor's first process() throws std::runtime_error.
 */
#ifndef STAIRPERCEPTION_B200_HPP_
#define STAIRPERCEPTION_B200_HPP_

#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "sp_b200.h"

#ifdef SP_B200_REFERENCE_TYPES
/* drop-in build inside the reference: /root/reference/include/common.h was included first and provides PCL plus
 * PointType, CloudType, ShowModeDef, ParameterDef and Plane themselves */
#else
namespace pcl {
struct alignas(16) PointXYZ { float x, y, z, pad_; };
struct alignas(16) PointXYZRGBA { float x, y, z, pad_; uint32_t rgba; uint32_t pad2_[3]; };              /* 32 B, as PCL */
struct alignas(16) Normal { float normal_x, normal_y, normal_z, pad_; float curvature; float pad2_[3]; };   /* 32 B, as PCL */
struct ModelCoefficients { std::vector<float> values; };
template <class PointT> struct PointCloud {
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    size_t size() const { return points.size(); }
};
}  // namespace pcl

typedef pcl::PointXYZRGBA PointType;
typedef pcl::PointCloud<PointType> CloudType;

/* /root/reference/include/common.h:79-92 */
enum ShowModeDef { original = 0, range_limited, voxel, noleg, horizontal_cloud, vertical_cloud, seg_horizontal_plane,
                   seg_vertical_plane, merge_horizontal_plane, merge_vertical_plane, stair };

/* /root/reference/include/common.h:94-131 -- same names, order and types */
typedef struct {
    double x_min_, x_max_, y_min_, y_max_, z_min_, z_max_;
    int normal_compute_points_;
    double voxel_x_, voxel_y_, voxel_z_;
    double parallel_angle_diff_, perpendicular_angle_diff_;
    double cluster_tolerance_;
    int min_cluster_size_, max_cluster_size_;
    double seg_threshold_;
    int seg_max_iters_, seg_rest_point_;
    double seg_plane_angle_diff_;
    double merge_threshold_, merge_angle_diff_;
    int min_num_points_;
    double min_length_, max_length_, min_width_, max_width_;
    double max_g_height_, counter_max_distance_;
    double min_height_, max_height_;
    double vertical_angle_diff_, mcv_angle_diff_, center_vector_angle_diff_, vertical_plane_angle_diff_;
    double noleg_distance_;
} ParameterDef;

/* /root/reference/include/common.h:150-185 */
typedef struct Plane_ {
    pcl::PointCloud<PointType> cloud;
    pcl::ModelCoefficients coefficients;
    pcl::PointXYZ center;
    pcl::PointXYZ pcenter;
    pcl::PointXYZ eigen_vectors[3];
    float eigen_values[3];
    float min[3], max[3];
    pcl::PointXYZRGBA points_min[3], points_max[3];
    pcl::PointCloud<pcl::PointXYZ> counter;
    typedef enum { vertical = 0, horizontal } Type;
    typedef enum { stair_component = 0, pstair_component, ground, others } Ptype;
    Ptype ptype;
    Type type;
} Plane;
#endif /* SP_B200_REFERENCE_TYPES */

namespace sp_b200 {

inline sp_params to_sp_params(const ParameterDef &p)
{
    static_assert(sizeof(ParameterDef) == sizeof(sp_params), "ParameterDef and sp_params must have the same layout");
    sp_params q;
    std::memcpy(&q, &p, sizeof(q));
    return q;
}

/* the values of /root/reference/param.yaml */
inline ParameterDef default_parameters()
{
    ParameterDef p = { -5., 5., -5., 5., -5., 5., 69, 1.0e-02, 1.0e-02, 1.0e-02, 1.9577777777777779e+01, 2.4466666666666665e+01,
                       2.9600000000000001e-02, 9, 100000, 1.0e-02, 100, 236, 7., 2.0e-02, 7., 150, 1.4999999999999999e-01, 1.,
                       5.0000000000000003e-02, 3.8000000000000000e-01, 5.0000000000000012e-02, 2.5e-01, 5.0000000000000003e-02,
                       2.0000000000000001e-01, 20., 80., 70., 20., 8.1100000000000005e-01 };
    return p;
}

/* process-level device state, one context per (device, width, height) */
class ContextPool {
public:
    static ContextPool &instance() { static ContextPool p; return p; }
    sp_ctx *get(int device, int width, int height, const sp_params &prm)
    {
        std::lock_guard<std::mutex> lk(mu_);
        const auto key = std::make_tuple(device, width, height);
        auto it = ctx_.find(key);
        if (it == ctx_.end()) {
            sp_ctx *c = nullptr;
            const int rc = sp_create(&prm, device, width, height, 1, &c);
            if (rc != SP_OK) throw std::runtime_error(std::string("sp_create failed: ") + sp_last_error(nullptr));
            it = ctx_.emplace(key, c).first;
        } else if (sp_set_params(it->second, &prm) != SP_OK) {
            throw std::runtime_error(std::string("sp_set_params failed: ") + sp_last_error(it->second));
        }
        return it->second;
    }
    ~ContextPool() { for (auto &kv : ctx_) sp_destroy(kv.second); }
private:
    std::mutex mu_;
    std::map<std::tuple<int, int, int>, sp_ctx *> ctx_;
};

/* The front-end behind StairDetection::process(): constructed per frame like the reference's object. */
class FrontEnd {
public:
    explicit FrontEnd(ParameterDef parameters, int device = 0) : parameters_(parameters), device_(device) {}

    /* Same argument meaning as the reference (StairPerception.cpp:5-10) minus the Stair& out-parameter, which only the
     * stair model writes.  `rotation` (optional, 9 floats row-major) applies rotationCloudPoints (test/reader.cpp:273-324)
     * on the device before everything else; NULL = cloud already levelled. */
    bool process(const pcl::PointCloud<PointType> &cloud_in, const ShowModeDef show_mode, std::vector<Plane> &vector_plane_sorted,
                 pcl::PointCloud<PointType> &cloud_show, pcl::PointCloud<pcl::Normal> &normal_show, const float *rotation = nullptr,
                 sp_stair *stair_out = nullptr)
    {
        if (stair_out) std::memset(stair_out, 0, sizeof(*stair_out));
        if (show_mode == original) return false;                       /* StairPerception.cpp:24 */
        if (cloud_in.points.empty()) return false;                     /* :27 */
        const int W = cloud_in.width ? (int)cloud_in.width : (int)cloud_in.points.size();
        const int H = cloud_in.height ? (int)cloud_in.height : 1;
        if ((size_t)W * H != cloud_in.points.size()) throw std::runtime_error("cloud_in is not organised (width * height != size)");
        sp_ctx *ctx = ContextPool::instance().get(device_, W, H, to_sp_params(parameters_));
        /* 32-byte in-memory points -> the 16-byte record the library takes (the PCD record layout) */
        struct Rec { float x, y, z; uint32_t rgba; };
        std::vector<Rec> rec(cloud_in.points.size());
        for (size_t i = 0; i < rec.size(); ++i) { const PointType &p = cloud_in.points[i]; rec[i] = Rec{ p.x, p.y, p.z, p.rgba }; }
        last_ = sp_frame_result();
        if (sp_process_frames(ctx, rec.data(), rotation, 1, (int)show_mode, &last_) != SP_OK)
            throw std::runtime_error(std::string("sp_process_frames failed: ") + sp_last_error(ctx));
        vector_plane_sorted.clear();
        const size_t N = rec.size();
        auto tap_i = [&](const char *name) {
            std::vector<int32_t> v;
            const int64_t nb = sp_tap_bytes(ctx, 0, name);
            if (nb > 0) { v.resize((size_t)nb / 4); sp_tap_copy(ctx, 0, name, v.data(), nb); }
            return v;
        };
        auto tap_f = [&](const char *name) {
            std::vector<float> v;
            const int64_t nb = sp_tap_bytes(ctx, 0, name);
            if (nb > 0) { v.resize((size_t)nb / 4); sp_tap_copy(ctx, 0, name, v.data(), nb); }
            return v;
        };
        std::vector<float> xyz, nrm;
        auto fill_show = [&](const std::vector<int32_t> &idx) {
            if (xyz.empty()) { xyz = tap_f("xyz"); nrm = tap_f("normals"); }
            cloud_show.points.clear(); normal_show.points.clear();
            for (int32_t i : idx) {
                PointType p = cloud_in.points[(size_t)i];
                p.x = xyz[3 * (size_t)i]; p.y = xyz[3 * (size_t)i + 1]; p.z = xyz[3 * (size_t)i + 2];
                cloud_show.points.push_back(p);
                pcl::Normal q = pcl::Normal();
                q.normal_x = nrm[3 * (size_t)i]; q.normal_y = nrm[3 * (size_t)i + 1]; q.normal_z = nrm[3 * (size_t)i + 2];
                normal_show.points.push_back(q);
            }
            cloud_show.width = normal_show.width = (uint32_t)idx.size(); cloud_show.height = normal_show.height = 1;
            cloud_show.is_dense = normal_show.is_dense = false;
        };
        auto make_cloud = [&](const int32_t *idx, int n, pcl::PointCloud<PointType> &out) {
            if (xyz.empty()) { xyz = tap_f("xyz"); nrm = tap_f("normals"); }
            out.points.resize((size_t)n);
            for (int k = 0; k < n; ++k) {
                PointType p = cloud_in.points[(size_t)idx[k]];
                p.x = xyz[3 * (size_t)idx[k]]; p.y = xyz[3 * (size_t)idx[k] + 1]; p.z = xyz[3 * (size_t)idx[k] + 2];
                out.points[(size_t)k] = p;
            }
            out.width = (uint32_t)n; out.height = 1; out.is_dense = false;
        };
        switch (show_mode) {
        case range_limited: {                                           /* :49-54 */
            std::vector<int32_t> idx;
            /* the crop survivors in input order: finite and inside the limits */
            if (xyz.empty()) { xyz = tap_f("xyz"); nrm = tap_f("normals"); }
            const float lo[3] = { (float)parameters_.x_min_, (float)parameters_.y_min_, (float)parameters_.z_min_ };
            const float hi[3] = { (float)parameters_.x_max_, (float)parameters_.y_max_, (float)parameters_.z_max_ };
            for (size_t i = 0; i < N; ++i) {
                bool ok = true;
                for (int k = 0; k < 3; ++k) { const float v = xyz[3 * i + k]; ok = ok && (v == v) && (v - v == 0.0f) && !(v < lo[k] || v > hi[k]); }
                if (ok) idx.push_back((int32_t)i);
            }
            fill_show(idx);
            return false;
        }
        case voxel: fill_show(tap_i("nonan_idx")); return false;        /* :63-68 shows the cloud after the NaN-normal filter */
        case noleg: fill_show(tap_i("noleg_idx")); return false;        /* :71-76 */
        case horizontal_cloud:                                          /* :90-101 */
        case vertical_cloud: {                                          /* :104-114 */
            fill_show(tap_i(show_mode == horizontal_cloud ? "h_idx" : "v_idx"));
            Plane plane = Plane();
            plane.cloud = cloud_show;
            plane.ptype = Plane::stair_component;
            vector_plane_sorted.push_back(plane);
            return false;
        }
        case seg_horizontal_plane: case seg_vertical_plane: case merge_horizontal_plane: case merge_vertical_plane: {
            const bool hz = show_mode == seg_horizontal_plane || show_mode == merge_horizontal_plane;
            const bool seg = show_mode == seg_horizontal_plane || show_mode == seg_vertical_plane;
            const std::string pre = std::string(hz ? "h_" : "v_") + (seg ? "seg_" : "merge_");
            const std::vector<float> coef = tap_f((pre + "coef").c_str());
            const std::vector<int32_t> cnt = tap_i((pre + "count").c_str()), idx = tap_i((pre + "idx").c_str());
            size_t off = 0;
            for (size_t k = 0; k < cnt.size(); ++k) {
                Plane plane = Plane();
                make_cloud(idx.data() + off, cnt[k], plane.cloud);
                plane.coefficients.values.assign(coef.begin() + 4 * k, coef.begin() + 4 * k + 4);
                plane.type = hz ? Plane::horizontal : Plane::vertical;
                plane.ptype = Plane::stair_component;
                vector_plane_sorted.push_back(plane);
                off += (size_t)cnt[k];
            }
            return false;
        }
        default: break;
        }
        /* stair: the sorted hand-off records (:185-188) */
        std::vector<int32_t> pix;
        for (int i = 0; i < last_.n_planes; ++i) {
            const sp_plane_record &r = last_.planes[i];
            Plane plane = Plane();
            pix.resize((size_t)r.count);
            std::vector<float> pxyz((size_t)r.count * 3);
            sp_copy_plane_points(ctx, 0, i, pxyz.data(), pix.data(), r.count);
            plane.cloud.points.resize((size_t)r.count);
            for (int k = 0; k < r.count; ++k) {
                PointType p = cloud_in.points[(size_t)pix[(size_t)k]];
                p.x = pxyz[3 * (size_t)k]; p.y = pxyz[3 * (size_t)k + 1]; p.z = pxyz[3 * (size_t)k + 2];
                plane.cloud.points[(size_t)k] = p;
            }
            plane.cloud.width = (uint32_t)r.count; plane.cloud.height = 1; plane.cloud.is_dense = false;
            plane.coefficients.values.assign(r.coef, r.coef + 4);
            plane.center.x = r.center[0]; plane.center.y = r.center[1]; plane.center.z = r.center[2];
            plane.pcenter.x = r.pcenter[0]; plane.pcenter.y = r.pcenter[1]; plane.pcenter.z = r.pcenter[2];
            for (int a = 0; a < 3; ++a) {
                plane.eigen_vectors[a].x = r.evec[3 * a]; plane.eigen_vectors[a].y = r.evec[3 * a + 1]; plane.eigen_vectors[a].z = r.evec[3 * a + 2];
                plane.eigen_values[a] = r.eval[a]; plane.min[a] = r.pmin[a]; plane.max[a] = r.pmax[a];
                plane.points_min[a] = PointType(); plane.points_max[a] = PointType();
                if (r.idx_min[a] >= 0) plane.points_min[a] = cloud_in.points[(size_t)r.idx_min[a]];
                if (r.idx_max[a] >= 0) plane.points_max[a] = cloud_in.points[(size_t)r.idx_max[a]];
                plane.points_min[a].x = r.pt_min[a][0]; plane.points_min[a].y = r.pt_min[a][1]; plane.points_min[a].z = r.pt_min[a][2];
                plane.points_max[a].x = r.pt_max[a][0]; plane.points_max[a].y = r.pt_max[a][1]; plane.points_max[a].z = r.pt_max[a][2];
            }
            plane.counter.points.resize(4);                             /* sortPlanesByXValues, :1446 */
            plane.type = r.type == 1 ? Plane::horizontal : Plane::vertical;
            plane.ptype = Plane::others;                                /* :1448 */
            vector_plane_sorted.push_back(plane);
        }
        if (stair_out) {
            /* the library's own restatement of stairModel (StairPerception.cpp:1621-2773): process() then answers
             * "stair found" exactly like the reference (:190) */
            if (sp_stair_model(ctx, 0, stair_out) != SP_OK) throw std::runtime_error(std::string("sp_stair_model failed: ") + sp_last_error(ctx));
            for (int i = 0; i < last_.n_planes && i < (int)vector_plane_sorted.size(); ++i) vector_plane_sorted[(size_t)i].ptype = (Plane::Ptype)stair_out->ptype[i];
            return stair_out->has_stair != 0;
        }
        return last_.n_planes > 0;
    }

    /* per-stage counts and status of the last frame (no counterpart in the reference: it prints them under DEBUG_CONSOLE) */
    const sp_frame_result &last_result() const { return last_; }

private:
    ParameterDef parameters_;
    int device_;
    sp_frame_result last_;
};

}  // namespace sp_b200

#ifndef SP_B200_REFERENCE_TYPES
/* stand-alone use: the reference's class name */
typedef sp_b200::FrontEnd StairDetection;
#endif

#endif /* STAIRPERCEPTION_B200_HPP_ */
