/*
 * sp_stair.h -- the consumer of the hand-off records: a PCL-free restatement of StairDetection::stairModel
 * (/root/reference/modules/stairperception/StairPerception.cpp:1621-2773) with computeLineFrom2Planes (:1096-1142) and
 * minDistaceOfTwoCloud (:1503-1613).  Host code, as in the reference (it works on a few tens of planes); the only things
 * it needs from the planes' point clouds come through StairPoints: findMinMaxProjPoint (:2780-2808) as an on-device
 * reduction, and the concatenated cloud of a merged ground plane (:1748-1794), fetched only when two ground planes exist.
 * When a stair is found: the two side ("slide") planes of step 8 (:2562-2612) and computePlaneCounter (:2861-3075), the contour
 * of every plane of the stair that the GUI draws.
 *
 * The reference's quirks are kept on purpose (SURVEY.md section 8a "other behaviours", Appendix B):
 *   - ground detection reads the extents of vector_plane_sorted[i] with the HEIGHT-LIST index i (:1720-1721) and stores that
 *     same i among the plane indices it later erases (:1727, :1782);
 *   - `if (a.ptype == b.ptype == pstair_component)` (:2443) compares the boolean with the enum value 1;
 *   - distances between planes are SQUARED metres compared with counter_max_distance in metres (:1503-1613, :2150).
 */
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/sp_b200.h"
#include "../../include/sp_detmath.h"

namespace sp_stairmodel {

enum { PT_STAIR = 0, PT_PSTAIR = 1, PT_GROUND = 2, PT_OTHERS = 3 };      /* Plane::Ptype, common.h:172-178 */
enum { TY_VERTICAL = 0, TY_HORIZONTAL = 1 };                              /* Plane::Type, common.h:166-170 */

struct MPlane {                     /* the fields of `Plane` the model reads */
    int type = TY_VERTICAL, ptype = PT_OTHERS, count = 0;
    float coef[4] = { 0, 0, 0, 0 }, center[3] = { 0, 0, 0 }, pcenter[3] = { 0, 0, 0 };
    float mn[3] = { 0, 0, 0 }, mx[3] = { 0, 0, 0 };
    float pt_min[3][3], pt_max[3][3];
    std::vector<int> src;           /* hand-off planes whose clouds this plane concatenates (empty = virtual plane) */
};

struct StairPoints {                /* access to the planes' point clouds */
    virtual ~StairPoints() {}
    /* xyz of hand-off plane `plane`, appended to `xyz` (3 floats per point, cloud order) */
    virtual bool gather(int plane, std::vector<float> &xyz) = 0;
    /* findMinMaxProjPoint on hand-off plane `plane` about `center`; ok = false when the plane is empty */
    virtual bool minmax(int plane, const float *center, const float *dir, float *max_xyz, float *min_xyz) = 0;
};

struct V3 { float x, y, z; };
inline V3 v3(float a, float b, float c) { V3 r = { a, b, c }; return r; }
inline float dot3(const V3 &a, const V3 &b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }       /* Eigen's 3-term reduction */
inline V3 normalized(const V3 &a) { const float n = sqrtf(dot3(a, a)); return v3(a.x / n, a.y / n, a.z / n); }
inline float deg2radf(float a) { return a * 0.017453293f; }                                      /* pcl::deg2rad(float) */
inline double rad2degd(double a) { return a * 57.29577951308232; }                               /* pcl::rad2deg(double) */

struct Line { float coeff[6]; float h, d; };

/* computeLineFrom2Planes, :1096-1142 */
inline Line line_from_two_planes(const float *p1, const float *p2)
{
    const float a1 = p1[0], a2 = p2[0], b1 = p1[1], b2 = p2[1], c1 = p1[2], c2 = p2[2], d1 = p1[3], d2 = p2[3];
    const float n0 = b1 * c2 - c1 * b2, n1 = c1 * a2 - a1 * c2, n2 = a1 * b2 - b1 * a2;       /* normal1.cross(normal2) */
    const float x0 = (b2 * d1 - b1 * d2) / (b1 * a2 - b2 * a1);
    const float y0 = (a2 * d1 - a1 * d2) / (a1 * b2 - a2 * b1);
    const float z0 = 0;
    const float t = -1 * (n0 * x0 + n1 * y0 + n2 * z0) / (n0 * n0 + n1 * n1 + n2 * n2);
    const float x = n0 * t + x0, y = n1 * t + y0, z = n2 * t + z0;
    Line l;
    l.coeff[0] = x0; l.coeff[1] = y0; l.coeff[2] = z0; l.coeff[3] = n0; l.coeff[4] = n1; l.coeff[5] = n2;
    l.h = x;
    l.d = (float)sqrt((double)(y * y + z * z));
    return l;
}

/* minDistaceOfTwoCloud, :1503-1613: smallest SQUARED distance between the six extreme points of each plane */
inline float min_distance(const MPlane &a, const MPlane &b)
{
    float best = FLT_MAX;
    for (int sa = 0; sa < 2; ++sa) for (int sb = 0; sb < 2; ++sb)
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            const float *p = sa ? a.pt_max[i] : a.pt_min[i], *q = sb ? b.pt_max[j] : b.pt_min[j];
            const float d = (p[0] - q[0]) * (p[0] - q[0]) + (p[1] - q[1]) * (p[1] - q[1]) + (p[2] - q[2]) * (p[2] - q[2]);
            if (d < best) best = d;
        }
    return best;
}

/* computeVectorPlaneInfo (:1196-1325) on a host copy of a cloud: used for the merged ground plane only */
inline void plane_info(const std::vector<float> &xyz, MPlane &pl)
{
    const size_t n = xyz.size() / 3;
    float xc = 0, yc = 0, zc = 0;
    for (size_t i = 0; i < n; ++i) { xc += xyz[3 * i]; yc += xyz[3 * i + 1]; zc += xyz[3 * i + 2]; }
    pl.center[0] = xc / n; pl.center[1] = yc / n; pl.center[2] = zc / n;
    float acc[6] = { 0, 0, 0, 0, 0, 0 };
    for (size_t i = 0; i < n; ++i) {
        const float px = xyz[3 * i] - pl.center[0], py = xyz[3 * i + 1] - pl.center[1], pz = xyz[3 * i + 2] - pl.center[2];
        acc[0] += px * px; acc[1] += px * py; acc[2] += px * pz; acc[3] += py * py; acc[4] += py * pz; acc[5] += pz * pz;
    }
    for (int k = 0; k < 6; ++k) acc[k] = acc[k] / (float)n;
    const float cov[9] = { acc[0], acc[1], acc[2], acc[1], acc[3], acc[4], acc[2], acc[4], acc[5] };
    float ev[3], evc[9];
    det_jacobi3f(cov, ev, evc);
    const float qn = std::nanf("");
    int imn[3] = { -1, -1, -1 }, imx[3] = { -1, -1, -1 };
    for (int a = 0; a < 3; ++a) {
        const float e0 = evc[0 * 3 + a], e1 = evc[1 * 3 + a], e2 = evc[2 * 3 + a];
        float mn = FLT_MAX, mx = -FLT_MAX;
        for (size_t i = 0; i < n; ++i) {
            const float px = xyz[3 * i] - pl.center[0], py = xyz[3 * i + 1] - pl.center[1], pz = xyz[3 * i + 2] - pl.center[2];
            const float pj = px * e0 + (py * e1 + pz * e2);
            if (pj > mx) { mx = pj; imx[a] = (int)i; }
            else if (pj < mn) { mn = pj; imn[a] = (int)i; }
        }
        pl.mn[a] = mn; pl.mx[a] = mx;
        for (int k = 0; k < 3; ++k) {
            pl.pt_min[a][k] = imn[a] >= 0 ? xyz[3 * (size_t)imn[a] + k] : qn;
            pl.pt_max[a][k] = imx[a] >= 0 ? xyz[3 * (size_t)imx[a] + k] : qn;
        }
    }
    for (int k = 0; k < 3; ++k) pl.pcenter[k] = (pl.pt_max[0][k] + pl.pt_min[0][k]) / 2;
    pl.count = (int)n;
}

/* findMinMaxProjPoint over a model plane = its source planes in concatenation order, strict compares (first wins) */
inline bool minmax_model_plane(StairPoints &pts, const MPlane &pl, const float *dir, float *max_xyz, float *min_xyz)
{
    float bmx = -FLT_MAX, bmn = FLT_MAX; bool any = false;
    for (int s : pl.src) {
        float a[3], b[3];
        if (!pts.minmax(s, pl.center, dir, a, b)) continue;
        const float pa = dir[0] * (a[0] - pl.center[0]) + (dir[1] * (a[1] - pl.center[1]) + dir[2] * (a[2] - pl.center[2]));
        const float pb = dir[0] * (b[0] - pl.center[0]) + (dir[1] * (b[1] - pl.center[1]) + dir[2] * (b[2] - pl.center[2]));
        if (pa > bmx) { bmx = pa; memcpy(max_xyz, a, 12); }
        if (pb < bmn) { bmn = pb; memcpy(min_xyz, b, 12); }
        any = true;
    }
    return any;
}

/* crossPointOfLineAndPlane, :2828-2844: line {x0, y0, z0, a, b, c}, plane {A, B, C, D} */
inline void cross_point(const float *l, const float *pl, float *pc)
{
    const float A = pl[0], B = pl[1], C = pl[2], D = pl[3];
    float t = -1 * (A * l[0] + B * l[1] + C * l[2] + D);
    t = t / (A * l[3] + B * l[4] + C * l[5]);
    pc[0] = l[3] * t + l[0]; pc[1] = l[4] * t + l[1]; pc[2] = l[5] * t + l[2];
}

struct DirCount { V3 dir; int count; };

/* returns has_stair; fills `out` (nodes in list order) and, per hand-off plane, the final Plane::Ptype */
inline bool stair_model(const sp_params &prm, const sp_frame_result &fr, StairPoints &pts, sp_stair *out)
{
    /* the float members of StairDetection (StairPerception.hpp:146-180) */
    const float max_length = (float)prm.max_length, max_width = (float)prm.max_width, min_width = (float)prm.min_width;
    const float counter_max_distance = (float)prm.counter_max_distance, min_height = (float)prm.min_height, max_height = (float)prm.max_height;
    const float center_vector_angle_diff = (float)prm.center_vector_angle_diff, mcv_angle_diff = (float)prm.mcv_angle_diff;
    const float max_g_height = (float)prm.max_g_height, vertical_angle_diff = (float)prm.vertical_angle_diff;

    memset(out, 0, sizeof(*out));
    for (int i = 0; i < SP_MAX_PLANES; ++i) out->ptype[i] = PT_OTHERS;
    std::vector<MPlane> P;
    P.reserve((size_t)fr.n_planes + 8);
    for (int i = 0; i < fr.n_planes; ++i) {
        const sp_plane_record &r = fr.planes[i];
        MPlane m;
        m.type = r.type; m.ptype = PT_OTHERS; m.count = r.count;
        memcpy(m.coef, r.coef, 16); memcpy(m.center, r.center, 12); memcpy(m.pcenter, r.pcenter, 12);
        memcpy(m.mn, r.pmin, 12); memcpy(m.mx, r.pmax, 12); memcpy(m.pt_min, r.pt_min, 36); memcpy(m.pt_max, r.pt_max, 36);
        m.src.push_back(i);
        P.push_back(m);
    }
    auto publish_ptypes = [&]() { for (const MPlane &m : P) for (int s : m.src) if (s >= 0 && s < SP_MAX_PLANES) out->ptype[s] = m.ptype; };

    /* ---- 1. mark ground (:1671-1808) ---- */
    struct PH { float height; int index; int point_num; };
    std::vector<PH> ph;
    bool first = true; float lowest_x = 0.f;
    for (size_t i = 0; i < P.size(); ++i)
        if (P[i].type == TY_HORIZONTAL) {
            PH h;
            if (first) { first = false; lowest_x = P[i].center[0]; h.height = 0; }
            else h.height = lowest_x - P[i].center[0];
            h.index = (int)i; h.point_num = P[i].count;
            ph.push_back(h);
        }
    std::vector<int> vec_g_index;
    int g_index = -1, p_ground_count = 0;
    for (size_t i = 0; i < ph.size(); ++i)
        if (ph[i].height < 2 * max_g_height && ph[i].point_num > 3000)
            if ((P[i].mx[0] - P[i].mn[0] > max_length) || (P[i].mx[1] - P[i].mn[1] > max_width)) {       /* quirk: P[i], not P[ph[i].index] */
                P[(size_t)ph[i].index].ptype = PT_GROUND;
                g_index = (int)i; vec_g_index.push_back((int)i); ++p_ground_count;                       /* quirk: stores i */
                break;
            }
    if (p_ground_count > 0) {
        for (size_t i = 0; i < ph.size(); ++i)
            if ((int)i != g_index && fabsf(ph[i].height - ph[(size_t)g_index].height) < max_g_height) {
                vec_g_index.push_back(ph[i].index);
                P[(size_t)ph[i].index].ptype = PT_GROUND;
                ++p_ground_count;
            }
        if (p_ground_count > 1) {
            MPlane g;
            std::vector<float> cloud;
            V3 gn = v3(0, 0, 0); float gd = 0; int total = 0;
            for (size_t i = 0; i < P.size(); ++i)
                if (P[i].ptype == PT_GROUND) {
                    for (int s : P[i].src) { if (!pts.gather(s, cloud)) return false; g.src.push_back(s); }
                    const float w = (float)P[i].count;
                    gn.x += P[i].coef[0] * w; gn.y += P[i].coef[1] * w; gn.z += P[i].coef[2] * w;
                    gd += P[i].coef[3] * w;
                    total += P[i].count;
                }
            gn = normalized(gn);
            g.coef[0] = gn.x; g.coef[1] = gn.y; g.coef[2] = gn.z; g.coef[3] = gd / total;
            g.type = TY_HORIZONTAL;
            plane_info(cloud, g);
            std::sort(vec_g_index.begin(), vec_g_index.end());
            for (int i = (int)vec_g_index.size() - 1; i >= 0; --i)
                if (vec_g_index[(size_t)i] >= 0 && vec_g_index[(size_t)i] < (int)P.size()) P.erase(P.begin() + vec_g_index[(size_t)i]);
            g.ptype = PT_GROUND;
            P.insert(P.begin(), g);
        }
    }
    for (size_t i = 0; i < P.size(); ++i)
        if (P[i].ptype != PT_GROUND) {
            if ((P[i].mx[0] - P[i].mn[0] > max_length) || (P[i].mx[1] - P[i].mn[1] > max_width)) P[i].ptype = PT_OTHERS;
            else P[i].ptype = PT_PSTAIR;
        }

    /* ---- 2. relations between planes (:1847-1872) ---- */
    const size_t d = P.size();
    struct Rel { V3 cd; float distance, delta_h, delta_d; };
    std::vector<Rel> rel(d * d);
    for (Rel &r : rel) { r.cd = v3(0, 0, 0); r.distance = r.delta_h = r.delta_d = 0; }
    for (size_t i = 0; i < d; ++i)
        for (size_t j = i + 1; j < d; ++j) {
            Rel &r = rel[i * d + j];
            r.distance = min_distance(P[i], P[j]);
            r.delta_h = fabsf(P[i].center[0] - P[j].center[0]);
            r.cd = v3(P[j].center[0] - P[i].center[0], P[j].center[1] - P[i].center[1], P[j].center[2] - P[i].center[2]);
        }

    /* ---- 3. main vertical plane direction (:1916-2004) ---- */
    std::vector<DirCount> mvn;
    bool two_main = false;
    {
        const float thr = deg2radf(vertical_angle_diff);
        for (size_t i = 0; i < d; ++i) {
            if (!(P[i].type == TY_VERTICAL && P[i].ptype == PT_PSTAIR)) continue;
            const V3 nrm = v3(P[i].coef[0], P[i].coef[1], P[i].coef[2]);
            const V3 dir1 = normalized(nrm);
            bool exist = false;
            for (DirCount &c : mvn) {
                const V3 dir2 = normalized(c.dir);
                const float dp = dot3(dir1, dir2);
                const float angle = (float)acos((double)dp - 0.0001);
                if (angle < thr) { c.dir.x += dir1.x; c.dir.y += dir1.y; c.dir.z += dir1.z; c.count += 1; exist = true; break; }
                else if (M_PI - (double)angle < (double)thr) { c.dir.x -= dir1.x; c.dir.y -= dir1.y; c.dir.z -= dir1.z; c.count += 1; exist = true; break; }
            }
            if (!exist) { DirCount c; c.dir = nrm; c.count = 1; mvn.push_back(c); }
        }
        for (DirCount &c : mvn) c.dir = normalized(c.dir);
        std::stable_sort(mvn.begin(), mvn.end(), [](const DirCount &a, const DirCount &b) { return a.count > b.count; });
        if (mvn.size() >= 2 && mvn[0].count == mvn[1].count) two_main = true;
    }
    if (mvn.empty()) { publish_ptypes(); return false; }

    /* ---- 4. main centre-vector direction (:2012-2349) ---- */
    for (size_t i = 0; i < d; ++i)
        for (size_t j = i + 1; j < d; ++j) rel[i * d + j].delta_d = fabsf(dot3(mvn[0].dir, rel[i * d + j].cd));
    auto off_main = [&](const MPlane &pl, const V3 &main) {
        const V3 n = v3(pl.coef[0], pl.coef[1], pl.coef[2]);
        const float angle = (float)rad2degd(acos((double)dot3(main, n) - 0.0001));
        return !((angle < vertical_angle_diff) || (180 - angle < vertical_angle_diff));
    };
    for (size_t i = 0; i < d; ++i)
        if (P[i].type == TY_VERTICAL && off_main(P[i], mvn[0].dir)) P[i].ptype = PT_OTHERS;
    std::vector<V3> cvd(2 * d, v3(0, 0, 0));
    for (size_t i = 0; i < d; ++i) {
        if (P[i].ptype != PT_PSTAIR) continue;
        const bool ih = P[i].type == TY_HORIZONTAL;
        float min_dis_same = FLT_MAX, min_dh_same = FLT_MAX, min_dis_cross = FLT_MAX, min_dh_cross = FLT_MAX;
        for (size_t j = i + 1; j < d; ++j) {
            if (P[j].ptype != PT_PSTAIR) continue;
            const Rel &r = rel[i * d + j];
            const bool jh = P[j].type == TY_HORIZONTAL;
            const bool same = ih == jh;
            /* the first branch of each pair of loops is "j horizontal", the second "j vertical"; cv_directions[i] takes the
             * first branch's hit, cv_directions[i + d] the second's */
            float &min_dis = same ? min_dis_same : min_dis_cross, &min_dh = same ? min_dh_same : min_dh_cross;
            if (r.distance < counter_max_distance && r.distance < min_dis) {
                min_dis = r.distance;
                bool ok;
                if (same) ok = r.delta_h > min_height && r.delta_h < max_height && r.delta_d > min_width && r.delta_d < max_width;
                else ok = r.delta_h < max_height / 2 && r.delta_d > min_width && r.delta_d < max_width / 2;
                if (ok && r.delta_h < min_dh) { min_dh = r.delta_h; cvd[jh ? i : i + d] = r.cd; }
            }
        }
    }
    std::vector<DirCount> mcv;
    for (size_t i = 0; i < cvd.size(); ++i) {
        if (!(cvd[i].x != 0 && cvd[i].y != 0 && cvd[i].z != 0)) continue;
        const V3 dir1 = normalized(cvd[i]);
        bool exist = false;
        for (DirCount &c : mcv) {
            const V3 dir2 = normalized(c.dir);
            const float angle = (float)rad2degd(acos((double)dot3(dir1, dir2) - 0.0001));
            if (angle < mcv_angle_diff) { c.dir.x += dir1.x; c.dir.y += dir1.y; c.dir.z += dir1.z; c.count += 1; exist = true; break; }
        }
        if (!exist) { DirCount c; c.dir = cvd[i]; c.count = 1; mcv.push_back(c); }
    }
    for (DirCount &c : mcv) c.dir = normalized(c.dir);
    std::stable_sort(mcv.begin(), mcv.end(), [](const DirCount &a, const DirCount &b) { return a.count > b.count; });
    if (mcv.empty()) { publish_ptypes(); return false; }

    /* ---- 5. fragments (:2361-2455) ---- */
    struct FT { int from, to; };
    std::vector<FT> frag;
    auto demote_if_off = [&](MPlane &pl) {
        if (pl.type != TY_VERTICAL) return;
        if (!off_main(pl, mvn[0].dir)) return;
        if (two_main) { if (off_main(pl, mvn[1].dir)) pl.ptype = PT_OTHERS; }
        else pl.ptype = PT_OTHERS;
    };
    for (size_t i = 0; i < d; ++i)
        for (size_t j = i + 1; j < d; ++j) {
            const V3 dir = normalized(rel[i * d + j].cd);
            const float angle = (float)rad2degd(acos((double)dot3(mcv[0].dir, dir) - 0.0001));
            if (angle < center_vector_angle_diff && rel[i * d + j].distance < counter_max_distance && P[i].ptype == PT_PSTAIR &&
                P[j].ptype == PT_PSTAIR) {
                demote_if_off(P[i]);
                demote_if_off(P[j]);
                if ((int)(P[i].ptype == P[j].ptype) == PT_PSTAIR) {                          /* quirk, :2443 */
                    FT ft = { (int)i, (int)j };
                    frag.push_back(ft);
                    break;
                }
            }
        }

    /* ---- 6. longest chain (:2471-2515) ---- */
    std::vector<int> chain;
    if (frag.size() > 1) {
        std::vector<std::vector<int>> lists; std::vector<int> cur;
        for (size_t i = 0; i + 1 < frag.size(); ++i) {
            if (cur.empty()) { cur.push_back(frag[i].from); cur.push_back(frag[i].to); }
            if (frag[i].to == frag[i + 1].from) cur.push_back(frag[i + 1].to);
            else { lists.push_back(cur); cur.clear(); }
        }
        lists.push_back(cur);
        size_t best = 0, at = 0;
        for (size_t i = 0; i < lists.size(); ++i) if (lists[i].size() > best) { best = lists[i].size(); at = i; }
        chain = lists[at];
    } else if (frag.size() == 1) { chain.push_back(frag[0].from); chain.push_back(frag[0].to); }
    else { publish_ptypes(); return false; }

    /* ---- 7. mark planes, main vertical normal (:2529-2559) ---- */
    V3 vn = v3(0, 0, 0);
    for (size_t i = 0; i < d; ++i)
        if (P[i].ptype != PT_GROUND) {
            P[i].ptype = PT_OTHERS;
            for (int c : chain)
                if ((int)i == c) {
                    if (P[i].type == TY_VERTICAL) {
                        const V3 n = v3(P[i].coef[0], P[i].coef[1], P[i].coef[2]);
                        const float w = (float)P[i].count;                        /* cloud.width == number of points */
                        if (dot3(n, vn) >= 0) { vn.x += n.x * w; vn.y += n.y * w; vn.z += n.z * w; }
                        else { vn.x -= n.x * w; vn.y -= n.y * w; vn.z -= n.z * w; }
                    }
                    P[i].ptype = PT_STAIR;
                    break;
                }
        }
    vn = normalized(vn);
    if (dot3(mcv[0].dir, vn) < 0) vn = v3(-vn.x, -vn.y, -vn.z);
    out->vnormal[0] = vn.x; out->vnormal[1] = vn.y; out->vnormal[2] = vn.z;
    /* ---- 8. section normal (:2567-2570); the slide planes follow below, once a stair is known to exist ---- */
    const V3 sn = normalized(v3(vn.y * 0 - vn.z * 0, vn.z * -1 - vn.x * 0, vn.x * 0 - vn.y * -1));         /* vnormal.cross((-1,0,0)) */
    out->snormal[0] = sn.x; out->snormal[1] = sn.y; out->snormal[2] = sn.z;

    /* ---- 9. modelling (:2622-2767) ---- */
    int count = 0, prev_concave = -1, prev_step = -1;
    std::vector<int> node_mh, node_mv;                         /* model plane (index into P) behind plane_h / plane_v of every node */
    auto push = [&](const sp_stair_node &n, int mh, int mv) {
        if (out->n_nodes < SP_MAX_STAIR_NODES) { out->nodes[out->n_nodes++] = n; node_mh.push_back(mh); node_mv.push_back(mv); }
        return out->n_nodes - 1;
    };
    auto src_of = [&](int model_plane) { return (P[(size_t)model_plane].src.size() == 1) ? P[(size_t)model_plane].src[0] : -1; };
    for (size_t ci = 0; ci + 1 < chain.size(); ++ci) {
        const int index = chain[ci], index_next = chain[ci + 1];
        const MPlane a = P[(size_t)index], b = P[(size_t)index_next];
        sp_stair_node n; memset(&n, 0, sizeof(n));
        if (a.type == TY_HORIZONTAL && b.type == TY_VERTICAL) {
            const Line l = line_from_two_planes(a.coef, b.coef);
            n.kind = 0; n.plane_h = src_of(index); n.plane_v = src_of(index_next); n.line_h = l.h; n.line_d = l.d; memcpy(n.line_coeff, l.coeff, 24);
            if (prev_step >= 0) out->nodes[prev_step].depth = fabs((double)(l.d - out->nodes[prev_step].line_d));
            prev_concave = push(n, index, index_next);
        } else if (a.type == TY_VERTICAL && b.type == TY_HORIZONTAL) {
            const Line l = line_from_two_planes(a.coef, b.coef);
            n.kind = 1; n.plane_v = src_of(index); n.plane_h = src_of(index_next); n.line_h = l.h; n.line_d = l.d; memcpy(n.line_coeff, l.coeff, 24);
            n.count = ++count;
            if (prev_concave >= 0) n.height = fabs((double)(l.h - out->nodes[prev_concave].line_h));
            else {
                const float dir[3] = { 1, 0, 0 }; float pmax[3] = { 0, 0, 0 }, pmin[3];
                minmax_model_plane(pts, a, dir, pmax, pmin);
                n.height = fabs((double)(pmax[0] - l.h));
            }
            prev_step = push(n, index_next, index);
        } else if (a.type == TY_HORIZONTAL && b.type == TY_HORIZONTAL) {
            const float dir[3] = { vn.x, vn.y, vn.z };
            float h0[3] = { 0, 0, 0 }, h1[3] = { 0, 0, 0 }, tmp[3];
            minmax_model_plane(pts, a, dir, h0, tmp);          /* the back point of the lower plane */
            minmax_model_plane(pts, b, dir, tmp, h1);          /* the front point of the upper plane */
            MPlane v;
            v.type = TY_VERTICAL;                              /* (the reference leaves type unset; it is only used as a vertical plane) */
            v.center[0] = (h0[0] + h1[0]) / 2; v.center[1] = (h0[1] + h1[1]) / 2; v.center[2] = (h0[2] + h1[2]) / 2;
            v.coef[0] = vn.x; v.coef[1] = vn.y; v.coef[2] = vn.z;
            v.coef[3] = -v.center[0] * vn.x - v.center[1] * vn.y - v.center[2] * vn.z;
            memcpy(v.pt_min[1], h0, 12); memcpy(v.pt_max[1], h1, 12);
            P.push_back(v);
            const int vidx = (int)P.size() - 1, vk = out->n_virtual < SP_MAX_VIRTUAL_PLANES ? out->n_virtual++ : SP_MAX_VIRTUAL_PLANES - 1;
            const Line l1 = line_from_two_planes(a.coef, v.coef);
            n.kind = 0; n.plane_h = src_of(index); n.plane_v = -2 - vk; n.line_h = l1.h; n.line_d = l1.d; memcpy(n.line_coeff, l1.coeff, 24);
            if (prev_step >= 0) out->nodes[prev_step].depth = fabs((double)(l1.d - out->nodes[prev_step].line_d));
            push(n, index, vidx);                              /* (prev_concave_line is not updated here in the reference, :2740) */
            const Line l2 = line_from_two_planes(b.coef, v.coef);
            sp_stair_node s; memset(&s, 0, sizeof(s));
            s.kind = 1; s.plane_v = -2 - vk; s.plane_h = src_of(index_next); s.line_h = l2.h; s.line_d = l2.d; memcpy(s.line_coeff, l2.coeff, 24);
            s.count = ++count;
            if (prev_concave >= 0) s.height = fabs((double)(l2.h - out->nodes[prev_concave].line_h));
            else s.height = fabs((double)(a.center[0] - b.center[0]));
            prev_step = push(s, index_next, vidx);
        }
    }
    publish_ptypes();
    out->n_steps = count;
    out->has_stair = count > 0 ? 1 : 0;
    if (count == 0) return false;

    /* ---- 8 (continued). the two side planes (:2571-2612): one sequential scan over the points of every non-ground plane of the
     * chain, projections about each plane's own centre, `if (max) ... else if (min)` as in the reference.  (The reference runs
     * it before step 9 whether or not a stair comes out; its only reader is computePlaneCounter, which needs a stair.) ---- */
    float right_max = -FLT_MAX, left_min = FLT_MAX, rmax[3] = { 0, 0, 0 }, lmin[3] = { 0, 0, 0 };
    std::vector<float> xyz;
    for (int idx : chain) {
        const MPlane &pl = P[(size_t)idx];
        if (pl.ptype == PT_GROUND) continue;                       /* (step 9 does not touch ptype) */
        xyz.clear();
        for (int sidx : pl.src) pts.gather(sidx, xyz);
        for (size_t i = 0; i + 2 < xyz.size(); i += 3) {
            const V3 vp = v3(xyz[i] - pl.center[0], xyz[i + 1] - pl.center[1], xyz[i + 2] - pl.center[2]);
            const float proj = dot3(sn, vp);
            if (proj > right_max) { right_max = proj; memcpy(rmax, &xyz[i], 12); }
            else if (proj < left_min) { left_min = proj; memcpy(lmin, &xyz[i], 12); }
        }
    }
    float sl[4] = { sn.x, sn.y, sn.z, -1 * (sn.x * lmin[0] + sn.y * lmin[1] + sn.z * lmin[2]) };
    float sr[4] = { sn.x, sn.y, sn.z, -1 * (sn.x * rmax[0] + sn.y * rmax[1] + sn.z * rmax[2]) };
    memcpy(out->slide_left, sl, 16); memcpy(out->slide_right, sr, 16);

    /* ---- computePlaneCounter (:2861-3075): walk the list; every plane collects its four contour points ---- */
    struct Contour { float p[4][3]; int width; };
    std::vector<Contour> C(P.size());
    for (Contour &c : C) { memset(c.p, 0, sizeof(c.p)); c.width = 0; }
    const float xdir[3] = { 1, 0, 0 }, vdir[3] = { vn.x, vn.y, vn.z };
    for (int ni = 0; ni < out->n_nodes; ++ni) {
        const sp_stair_node &nd = out->nodes[ni];
        const bool last = ni == out->n_nodes - 1;
        const float *ln = nd.line_coeff;
        Contour &cv = C[(size_t)node_mv[(size_t)ni]], &ch = C[(size_t)node_mh[(size_t)ni]];
        const MPlane &pv = P[(size_t)node_mv[(size_t)ni]], &ph = P[(size_t)node_mh[(size_t)ni]];
        if (nd.kind == 1) {                                                   /* step: riser below, tread behind the line */
            if (cv.width == 0) {
                float mx[3] = { 0, 0, 0 }, mn[3] = { 0, 0, 0 };              /* lower edge: through the lowest point of the riser */
                minmax_model_plane(pts, pv, xdir, mx, mn);
                const float ld[6] = { mx[0], mx[1], mx[2], ln[3], ln[4], ln[5] };
                cross_point(ld, sl, cv.p[0]); cross_point(ld, sr, cv.p[1]);
                cross_point(ln, sr, cv.p[2]); cross_point(ln, sl, cv.p[3]);
                cv.width = 4;
            } else if (cv.width == 2) {
                cross_point(ln, sr, cv.p[2]); cross_point(ln, sl, cv.p[3]);
                cv.width = 4;
            }
            if (ch.width == 0) {
                cross_point(ln, sl, ch.p[0]); cross_point(ln, sr, ch.p[1]);
                ch.width = 2;
                if (last) {                                                   /* back edge: through the farthest point of the tread */
                    float mx[3] = { 0, 0, 0 }, mn[3] = { 0, 0, 0 };
                    minmax_model_plane(pts, ph, vdir, mx, mn);
                    const float lb[6] = { mx[0], mx[1], mx[2], ln[3], ln[4], ln[5] };
                    cross_point(lb, sr, ch.p[2]); cross_point(lb, sl, ch.p[3]);
                    ch.width = 4;
                }
            }
        } else {                                                              /* concave line: tread in front, riser above */
            if (ch.width == 0) {
                float mx[3] = { 0, 0, 0 }, mn[3] = { 0, 0, 0 };              /* front edge: through the nearest point of the tread */
                minmax_model_plane(pts, ph, vdir, mx, mn);
                const float lf[6] = { mn[0], mn[1], mn[2], ln[3], ln[4], ln[5] };
                cross_point(lf, sl, ch.p[0]); cross_point(lf, sr, ch.p[1]);
                cross_point(ln, sr, ch.p[2]); cross_point(ln, sl, ch.p[3]);
                ch.width = 4;
            } else if (ch.width == 2) {
                cross_point(ln, sr, ch.p[2]); cross_point(ln, sl, ch.p[3]);
                ch.width = 4;
            }
            if (cv.width == 0) {
                cross_point(ln, sl, cv.p[0]); cross_point(ln, sr, cv.p[1]);
                cv.width = 2;
                if (last) {                                                   /* upper edge: through the highest point of the riser */
                    float mx[3] = { 0, 0, 0 }, mn[3] = { 0, 0, 0 };
                    minmax_model_plane(pts, pv, xdir, mx, mn);
                    const float lu[6] = { mn[0], mn[1], mn[2], ln[3], ln[4], ln[5] };
                    cross_point(lu, sr, cv.p[2]); cross_point(lu, sl, cv.p[3]);
                    cv.width = 4;
                }
            }
        }
    }
    {
        int vk = 0;
        for (size_t i = 0; i < P.size(); ++i) {
            if (P[i].src.size() == 1 && P[i].src[0] >= 0 && P[i].src[0] < SP_MAX_PLANES) {
                memcpy(out->contour[P[i].src[0]], C[i].p, sizeof(C[i].p)); out->contour_width[P[i].src[0]] = C[i].width;
            } else if (P[i].src.size() > 1) {
                memcpy(out->ground_contour, C[i].p, sizeof(C[i].p)); out->ground_contour_width = C[i].width;
            } else if (P[i].src.empty() && vk < SP_MAX_VIRTUAL_PLANES) {
                memcpy(out->virtual_contour[vk], C[i].p, sizeof(C[i].p)); out->virtual_contour_width[vk] = C[i].width; ++vk;
            }
        }
    }
    return true;
}

}  // namespace sp_stairmodel
