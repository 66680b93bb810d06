"""CPU ORACLE of the stair model -- TEST INFRASTRUCTURE ONLY (never imported by the product).

An independent numpy/float32 restatement of StairDetection::stairModel
(/root/reference/modules/stairperception/StairPerception.cpp:1621-2773), computeLineFrom2Planes (:1096-1142),
minDistaceOfTwoCloud (:1503-1613), findMinMaxProjPoint (:2780-2808) and, for the merged ground plane, computeVectorPlaneInfo
(:1196-1325).  It works on the oracle's own hand-off planes and their point clouds, all on the host.  Parity unpinned like the
rest of the oracle; its external anchor is the step table the reference publishes for samples/0001 (pic/screenshot2.png).
The reference's quirks are kept (height-list index used as plane index at :1720-1727, `a == b == c` at :2443, squared
distances against counter_max_distance).
"""
import math

import numpy as np

F = np.float32
PT_STAIR, PT_PSTAIR, PT_GROUND, PT_OTHERS = 0, 1, 2, 3
VERTICAL, HORIZONTAL = 0, 1
FLT_MAX = float(np.finfo(np.float32).max)


def _dot(a, b):
    return F(F(a[0] * b[0]) + F(F(a[1] * b[1]) + F(a[2] * b[2])))


def _normalized(a):
    a = np.asarray(a, dtype=F)
    n = F(np.sqrt(_dot(a, a)))
    return (a / n).astype(F)


def _deg2rad(a):
    return F(F(a) * F(0.017453293))


def _rad2deg(a):
    return float(a) * 57.29577951308232


def _acos(v):
    return math.acos(v) if -1.0 <= v <= 1.0 else float("nan")          # C's acos: NaN outside the domain, no exception


def _angle_rad(dp):
    return F(_acos(float(F(dp)) - 0.0001))


def _angle_deg(dp):
    return F(_rad2deg(_acos(float(F(dp)) - 0.0001)))


def line_from_two_planes(p1, p2):
    a1, b1, c1, d1 = [F(v) for v in p1]
    a2, b2, c2, d2 = [F(v) for v in p2]
    n = np.array([F(b1 * c2) - F(c1 * b2), F(c1 * a2) - F(a1 * c2), F(a1 * b2) - F(b1 * a2)], dtype=F)
    x0 = F(F(F(b2 * d1) - F(b1 * d2)) / F(F(b1 * a2) - F(b2 * a1)))
    y0 = F(F(F(a2 * d1) - F(a1 * d2)) / F(F(a1 * b2) - F(a2 * b1)))
    z0 = F(0)
    num = F(F(F(n[0] * x0) + F(n[1] * y0)) + F(n[2] * z0))
    den = F(F(F(n[0] * n[0]) + F(n[1] * n[1])) + F(n[2] * n[2]))
    t = F(F(F(-1) * num) / den)
    x, y, z = F(F(n[0] * t) + x0), F(F(n[1] * t) + y0), F(F(n[2] * t) + z0)
    return {"coeff": [x0, y0, z0, n[0], n[1], n[2]], "h": x, "d": F(math.sqrt(float(F(F(y * y) + F(z * z)))))}


class MPlane:
    def __init__(self):
        self.type, self.ptype, self.count = VERTICAL, PT_OTHERS, 0
        self.coef = np.zeros(4, F)
        self.center = np.zeros(3, F)
        self.pcenter = np.zeros(3, F)
        self.mn = np.zeros(3, F)
        self.mx = np.zeros(3, F)
        self.pt_min = np.full((3, 3), np.nan, F)
        self.pt_max = np.full((3, 3), np.nan, F)
        self.xyz = np.zeros((0, 3), F)          # the plane's cloud
        self.src = []


def min_distance(a, b):
    pa = np.concatenate([a.pt_min, a.pt_max]).astype(F)
    pb = np.concatenate([b.pt_min, b.pt_max]).astype(F)
    d = pa[:, None, :] - pb[None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    d2 = d2[~np.isnan(d2)]
    return F(d2.min()) if d2.size else F(FLT_MAX)


def _jacobi3(cov):
    """det_jacobi3f of include/sp_detmath.h restated in float32 (cyclic Jacobi, 12 sweeps, descending order)"""
    a = np.array([[cov[0], cov[1], cov[2]], [cov[1], cov[4], cov[5]], [cov[2], cov[5], cov[8]]], dtype=F)
    v = np.eye(3, dtype=F)
    for _ in range(12):
        for pi in range(3):
            p = 1 if pi == 2 else 0
            q = 1 if pi == 0 else 2
            apq = a[p, q]
            if apq == 0:
                continue
            theta = F(F(a[q, q] - a[p, p]) / F(F(2.0) * apq))
            t = F(F(1.0) / F(F(abs(theta)) + F(np.sqrt(F(F(theta * theta) + F(1.0))))))
            if theta < 0:
                t = F(-t)
            c = F(F(1.0) / F(np.sqrt(F(F(t * t) + F(1.0)))))
            s = F(t * c)
            r = 3 - p - q
            app, aqq = a[p, p], a[q, q]
            a[p, p] = F(app - F(t * apq))
            a[q, q] = F(aqq + F(t * apq))
            a[p, q] = a[q, p] = F(0)
            arp, arq = a[r, p], a[r, q]
            a[r, p] = a[p, r] = F(F(c * arp) - F(s * arq))
            a[r, q] = a[q, r] = F(F(s * arp) + F(c * arq))
            for k in range(3):
                vkp, vkq = v[k, p], v[k, q]
                v[k, p] = F(F(c * vkp) - F(s * vkq))
                v[k, q] = F(F(s * vkp) + F(c * vkq))
    dd = [a[0, 0], a[1, 1], a[2, 2]]
    o = [0, 1, 2]
    if dd[o[0]] < dd[o[1]]:
        o[0], o[1] = o[1], o[0]
    if dd[o[1]] < dd[o[2]]:
        o[1], o[2] = o[2], o[1]
        if dd[o[0]] < dd[o[1]]:
            o[0], o[1] = o[1], o[0]
    return np.array([dd[k] for k in o], F), v[:, o]


def plane_info(pl):
    """computeVectorPlaneInfo (:1196-1325): sequential float sums, `if .. else if` extent scan"""
    xyz = pl.xyz.astype(F)
    n = len(xyz)
    pl.center = (np.cumsum(xyz, axis=0, dtype=F)[-1] / F(n)).astype(F)
    c = (xyz - pl.center).astype(F)
    prods = np.stack([c[:, 0] * c[:, 0], c[:, 0] * c[:, 1], c[:, 0] * c[:, 2], c[:, 1] * c[:, 1], c[:, 1] * c[:, 2], c[:, 2] * c[:, 2]], axis=1).astype(F)
    acc = (np.cumsum(prods, axis=0, dtype=F)[-1] / F(n)).astype(F)
    cov = [acc[0], acc[1], acc[2], acc[1], acc[3], acc[4], acc[2], acc[4], acc[5]]
    _, evc = _jacobi3(cov)
    for a in range(3):
        e = evc[:, a]
        pj = (c[:, 0] * e[0] + (c[:, 1] * e[1] + c[:, 2] * e[2])).astype(F)
        mx, mn, imx, imn = F(-FLT_MAX), F(FLT_MAX), -1, -1
        for i in range(n):
            if pj[i] > mx:
                mx, imx = pj[i], i
            elif pj[i] < mn:
                mn, imn = pj[i], i
        pl.mn[a], pl.mx[a] = mn, mx
        pl.pt_min[a] = xyz[imn] if imn >= 0 else np.nan
        pl.pt_max[a] = xyz[imx] if imx >= 0 else np.nan
    pl.pcenter = ((pl.pt_max[0] + pl.pt_min[0]) / F(2)).astype(F)
    pl.count = n


def minmax_proj(pl, v):
    """findMinMaxProjPoint (:2780-2808): first point with the largest / smallest projection of (p - centre) on v"""
    v = np.asarray(v, F)
    c = (pl.xyz - pl.center).astype(F)
    proj = (v[0] * c[:, 0] + (v[1] * c[:, 1] + v[2] * c[:, 2])).astype(F)
    return pl.xyz[int(np.argmax(proj))], pl.xyz[int(np.argmin(proj))]


def cross_point(line, plane):
    """crossPointOfLineAndPlane (:2828-2844): line = (x0, y0, z0, a, b, c), plane = (A, B, C, D)"""
    l, q = np.asarray(line, F), np.asarray(plane, F)
    t = F(F(-1) * F(F(F(F(q[0] * l[0]) + F(q[1] * l[1])) + F(q[2] * l[2])) + q[3]))
    t = F(t / F(F(F(q[0] * l[3]) + F(q[1] * l[4])) + F(q[2] * l[5])))
    return np.array([F(F(l[3] * t) + l[0]), F(F(l[4] * t) + l[1]), F(F(l[5] * t) + l[2])], F)


def _extreme(pl, v):
    """findMinMaxProjPoint on a plane that may have no cloud (a virtual riser): PCL's default points are (0, 0, 0)"""
    if pl.xyz is None or len(pl.xyz) == 0:
        return np.zeros(3, F), np.zeros(3, F)
    return minmax_proj(pl, v)


def stair_model(params, planes, xyz, plane_idx):
    with np.errstate(all="ignore"):
        return _stair_model(params, planes, xyz, plane_idx)


def _stair_model(params, planes, xyz, plane_idx):
    """planes: the oracle's "planes" tap; xyz: (N, 3) levelled cloud; plane_idx: concatenated pixel lists.
    Returns {"has_stair", "steps": [(height, depth, line.h, line.d)], "nodes": [...], "ptype": [...]}"""
    g = lambda k: F(params[k])
    max_length, max_width, min_width = g("max_length"), g("max_width"), g("min_width")
    cmd, min_height, max_height = g("counter_max_distance"), g("min_height"), g("max_height")
    cvad, mcvad, max_g, vad = g("center_vector_angle_diff"), g("mcv_angle_diff"), g("max_g_height"), g("vertical_angle_diff")
    P = []
    for i, r in enumerate(planes):
        m = MPlane()
        m.type, m.count = int(r["type"]), int(r["count"])
        m.coef, m.center, m.pcenter = r["coef"].astype(F), r["center"].astype(F), r["pcenter"].astype(F)
        m.mn, m.mx, m.pt_min, m.pt_max = r["pmin"].astype(F), r["pmax"].astype(F), r["pt_min"].astype(F), r["pt_max"].astype(F)
        m.xyz = xyz[plane_idx[int(r["first"]):int(r["first"]) + int(r["count"])]].astype(F)
        m.src = [i]
        P.append(m)
    out = {"has_stair": False, "steps": [], "nodes": [], "ptype": [PT_OTHERS] * len(planes)}

    def publish():
        for m in P:
            for s in m.src:
                out["ptype"][s] = m.ptype
        return out

    # 1. ground
    ph, first, lowest = [], True, F(0)
    for i, m in enumerate(P):
        if m.type == HORIZONTAL:
            if first:
                first, lowest, h = False, m.center[0], F(0)
            else:
                h = F(lowest - m.center[0])
            ph.append((h, i, m.count))
    vec_g, g_index, cnt = [], -1, 0
    for i, (h, idx, npts) in enumerate(ph):
        if h < F(2) * max_g and npts > 3000:
            if (F(P[i].mx[0] - P[i].mn[0]) > max_length) or (F(P[i].mx[1] - P[i].mn[1]) > max_width):
                P[idx].ptype = PT_GROUND
                g_index = i
                vec_g.append(i)
                cnt += 1
                break
    if cnt > 0:
        for i, (h, idx, npts) in enumerate(ph):
            if i != g_index and F(abs(F(h - ph[g_index][0]))) < max_g:
                vec_g.append(idx)
                P[idx].ptype = PT_GROUND
                cnt += 1
        if cnt > 1:
            gp = MPlane()
            gn, gd, total, clouds = np.zeros(3, F), F(0), 0, []
            for m in P:
                if m.ptype == PT_GROUND:
                    clouds.append(m.xyz)
                    gp.src += m.src
                    w = F(m.count)
                    gn = (gn + m.coef[:3] * w).astype(F)
                    gd = F(gd + F(m.coef[3] * w))
                    total += m.count
            gn = _normalized(gn)
            gp.coef = np.array([gn[0], gn[1], gn[2], F(gd / F(total))], F)
            gp.type = HORIZONTAL
            gp.xyz = np.concatenate(clouds).astype(F)
            plane_info(gp)
            for i in sorted(vec_g, reverse=True):
                if 0 <= i < len(P):
                    del P[i]
            gp.ptype = PT_GROUND
            P.insert(0, gp)
    for m in P:
        if m.ptype != PT_GROUND:
            big = (F(m.mx[0] - m.mn[0]) > max_length) or (F(m.mx[1] - m.mn[1]) > max_width)
            m.ptype = PT_OTHERS if big else PT_PSTAIR

    # 2. relations
    d = len(P)
    dist = np.zeros((d, d), F)
    dh = np.zeros((d, d), F)
    dd = np.zeros((d, d), F)
    cd = np.zeros((d, d, 3), F)
    for i in range(d):
        for j in range(i + 1, d):
            dist[i, j] = min_distance(P[i], P[j])
            dh[i, j] = F(abs(F(P[i].center[0] - P[j].center[0])))
            cd[i, j] = (P[j].center - P[i].center).astype(F)

    # 3. main vertical normal
    def cluster_dirs(items, both_ways, thr_rad=None, thr_deg=None):
        cl = []
        for raw in items:
            d1 = _normalized(raw)
            hit = False
            for c in cl:
                d2 = _normalized(c[0])
                dp = _dot(d1, d2)
                if thr_rad is not None:
                    ang = _angle_rad(dp)
                    if ang < thr_rad:
                        c[0] = (c[0] + d1).astype(F); c[1] += 1; hit = True; break
                    if both_ways and (math.pi - float(ang)) < float(thr_rad):
                        c[0] = (c[0] - d1).astype(F); c[1] += 1; hit = True; break
                else:
                    if _angle_deg(dp) < thr_deg:
                        c[0] = (c[0] + d1).astype(F); c[1] += 1; hit = True; break
            if not hit:
                cl.append([np.asarray(raw, F).copy(), 1])
        cl = [[_normalized(c[0]), c[1]] for c in cl]
        cl.sort(key=lambda c: -c[1])                                   # stable, like insertion sort on a short vector
        return cl

    mvn = cluster_dirs([m.coef[:3] for m in P if m.type == VERTICAL and m.ptype == PT_PSTAIR], True, thr_rad=_deg2rad(vad))
    two_main = len(mvn) >= 2 and mvn[0][1] == mvn[1][1]
    if not mvn:
        return publish()

    # 4. main centre-vector direction
    for i in range(d):
        for j in range(i + 1, d):
            dd[i, j] = F(abs(_dot(mvn[0][0], cd[i, j])))

    def off_main(m, main):
        ang = _angle_deg(_dot(main, m.coef[:3]))
        return not ((ang < vad) or (F(F(180) - ang) < vad))

    for m in P:
        if m.type == VERTICAL and off_main(m, mvn[0][0]):
            m.ptype = PT_OTHERS
    cvd = np.zeros((2 * d, 3), F)
    for i in range(d):
        if P[i].ptype != PT_PSTAIR:
            continue
        ih = P[i].type == HORIZONTAL
        best = {True: [F(FLT_MAX), F(FLT_MAX)], False: [F(FLT_MAX), F(FLT_MAX)]}       # same / cross: [min distance, min delta_h]
        for j in range(i + 1, d):
            if P[j].ptype != PT_PSTAIR:
                continue
            jh = P[j].type == HORIZONTAL
            same = ih == jh
            b = best[same]
            if dist[i, j] < cmd and dist[i, j] < b[0]:
                b[0] = dist[i, j]
                if same:
                    ok = dh[i, j] > min_height and dh[i, j] < max_height and dd[i, j] > min_width and dd[i, j] < max_width
                else:
                    ok = dh[i, j] < F(max_height / F(2)) and dd[i, j] > min_width and dd[i, j] < F(max_width / F(2))
                if ok and dh[i, j] < b[1]:
                    b[1] = dh[i, j]
                    cvd[i if jh else i + d] = cd[i, j]
    mcv = cluster_dirs([v for v in cvd if v[0] != 0 and v[1] != 0 and v[2] != 0], False, thr_deg=mcvad)
    if not mcv:
        return publish()

    # 5. fragments
    def demote(m):
        if m.type != VERTICAL or not off_main(m, mvn[0][0]):
            return
        if two_main:
            if off_main(m, mvn[1][0]):
                m.ptype = PT_OTHERS
        else:
            m.ptype = PT_OTHERS

    frag = []
    for i in range(d):
        for j in range(i + 1, d):
            ang = _angle_deg(_dot(mcv[0][0], _normalized(cd[i, j])))
            if ang < cvad and dist[i, j] < cmd and P[i].ptype == PT_PSTAIR and P[j].ptype == PT_PSTAIR:
                demote(P[i])
                demote(P[j])
                if int(P[i].ptype == P[j].ptype) == PT_PSTAIR:
                    frag.append((i, j))
                    break

    # 6. longest chain
    if len(frag) > 1:
        lists, cur = [], []
        for i in range(len(frag) - 1):
            if not cur:
                cur = [frag[i][0], frag[i][1]]
            if frag[i][1] == frag[i + 1][0]:
                cur.append(frag[i + 1][1])
            else:
                lists.append(cur)
                cur = []
        lists.append(cur)
        chain = max(lists, key=len) if lists else []
        chain = lists[[len(x) for x in lists].index(max(len(x) for x in lists))]
    elif len(frag) == 1:
        chain = [frag[0][0], frag[0][1]]
    else:
        return publish()

    # 7. marks, vnormal
    vn = np.zeros(3, F)
    for i, m in enumerate(P):
        if m.ptype != PT_GROUND:
            m.ptype = PT_OTHERS
            if i in chain:
                if m.type == VERTICAL:
                    n3, w = m.coef[:3], F(m.count)
                    vn = (vn + n3 * w).astype(F) if _dot(n3, vn) >= 0 else (vn - n3 * w).astype(F)
                m.ptype = PT_STAIR
    vn = _normalized(vn)
    if _dot(mcv[0][0], vn) < 0:
        vn = (-vn).astype(F)
    out["vnormal"] = vn

    # 9. modelling
    count, prev_concave, prev_step = 0, None, None
    for ci in range(len(chain) - 1):
        a, b = P[chain[ci]], P[chain[ci + 1]]
        if a.type == HORIZONTAL and b.type == VERTICAL:
            l = line_from_two_planes(a.coef, b.coef)
            node = {"kind": 0, "line": l, "mh": a, "mv": b}
            if prev_step is not None:
                prev_step["depth"] = abs(float(F(l["d"] - prev_step["line"]["d"])))
            out["nodes"].append(node)
            prev_concave = node
        elif a.type == VERTICAL and b.type == HORIZONTAL:
            l = line_from_two_planes(a.coef, b.coef)
            count += 1
            node = {"kind": 1, "line": l, "count": count, "height": 0.0, "depth": 0.0, "mh": b, "mv": a}
            if prev_concave is not None:
                node["height"] = abs(float(F(l["h"] - prev_concave["line"]["h"])))
            else:
                pmax, _ = minmax_proj(a, [1, 0, 0])
                node["height"] = abs(float(F(pmax[0] - l["h"])))
            out["nodes"].append(node)
            prev_step = node
        elif a.type == HORIZONTAL and b.type == HORIZONTAL:
            h0, _ = minmax_proj(a, vn)
            _, h1 = minmax_proj(b, vn)
            v = MPlane()
            v.center = ((h0 + h1) / F(2)).astype(F)
            v.coef = np.array([vn[0], vn[1], vn[2], F(F(F(-v.center[0] * vn[0]) - F(v.center[1] * vn[1])) - F(v.center[2] * vn[2]))], F)
            P.append(v)
            l1 = line_from_two_planes(a.coef, v.coef)
            node = {"kind": 0, "line": l1, "mh": a, "mv": v}
            if prev_step is not None:
                prev_step["depth"] = abs(float(F(l1["d"] - prev_step["line"]["d"])))
            out["nodes"].append(node)
            l2 = line_from_two_planes(b.coef, v.coef)
            count += 1
            st = {"kind": 1, "line": l2, "count": count, "height": 0.0, "depth": 0.0, "mh": b, "mv": v}
            st["height"] = abs(float(F(l2["h"] - prev_concave["line"]["h"]))) if prev_concave is not None else abs(float(F(a.center[0] - b.center[0])))
            out["nodes"].append(st)
            prev_step = st
    publish()
    out["has_stair"] = count > 0
    out["steps"] = [(n["height"], n["depth"], float(n["line"]["h"]), float(n["line"]["d"])) for n in out["nodes"] if n["kind"] == 1]
    if count == 0:
        return out

    # 8. the side planes (:2562-2612): sequential `if .. else if` extreme scan along the section normal over the stair planes
    sn = _normalized(np.cross(vn.astype(np.float64), np.array([-1.0, 0.0, 0.0])).astype(F))
    right_max, left_min, rmax, lmin = F(-FLT_MAX), F(FLT_MAX), np.zeros(3, F), np.zeros(3, F)
    for idx in chain:
        m = P[idx]
        if m.ptype == PT_GROUND:
            continue
        c = (m.xyz - m.center).astype(F)
        proj = (sn[0] * c[:, 0] + (sn[1] * c[:, 1] + sn[2] * c[:, 2])).astype(F)
        for i in range(len(proj)):
            if proj[i] > right_max:
                right_max, rmax = proj[i], m.xyz[i]
            elif proj[i] < left_min:
                left_min, lmin = proj[i], m.xyz[i]
    sl = np.array([sn[0], sn[1], sn[2], F(F(-1) * F(F(F(sn[0] * lmin[0]) + F(sn[1] * lmin[1])) + F(sn[2] * lmin[2])))], F)
    sr = np.array([sn[0], sn[1], sn[2], F(F(-1) * F(F(F(sn[0] * rmax[0]) + F(sn[1] * rmax[1])) + F(sn[2] * rmax[2])))], F)
    out["slide_left"], out["slide_right"] = sl, sr

    # computePlaneCounter (:2861-3075)
    for m in P:
        m.contour, m.cwidth = np.zeros((4, 3), F), 0
    nodes = out["nodes"]
    for ni, nd in enumerate(nodes):
        last = ni == len(nodes) - 1
        ln = np.asarray(nd["line"]["coeff"], F)
        mh, mv = nd["mh"], nd["mv"]
        with_dir = lambda pt: np.concatenate([np.asarray(pt, F), ln[3:6]])
        if nd["kind"] == 1:
            if mv.cwidth == 0:
                pmax, _ = _extreme(mv, [1, 0, 0])
                ld = with_dir(pmax)
                mv.contour[0], mv.contour[1], mv.contour[2], mv.contour[3] = cross_point(ld, sl), cross_point(ld, sr), cross_point(ln, sr), cross_point(ln, sl)
                mv.cwidth = 4
            elif mv.cwidth == 2:
                mv.contour[2], mv.contour[3] = cross_point(ln, sr), cross_point(ln, sl)
                mv.cwidth = 4
            if mh.cwidth == 0:
                mh.contour[0], mh.contour[1] = cross_point(ln, sl), cross_point(ln, sr)
                mh.cwidth = 2
                if last:
                    pmax, _ = _extreme(mh, vn)
                    lb = with_dir(pmax)
                    mh.contour[2], mh.contour[3] = cross_point(lb, sr), cross_point(lb, sl)
                    mh.cwidth = 4
        else:
            if mh.cwidth == 0:
                _, pmin = _extreme(mh, vn)
                lf = with_dir(pmin)
                mh.contour[0], mh.contour[1], mh.contour[2], mh.contour[3] = cross_point(lf, sl), cross_point(lf, sr), cross_point(ln, sr), cross_point(ln, sl)
                mh.cwidth = 4
            elif mh.cwidth == 2:
                mh.contour[2], mh.contour[3] = cross_point(ln, sr), cross_point(ln, sl)
                mh.cwidth = 4
            if mv.cwidth == 0:
                mv.contour[0], mv.contour[1] = cross_point(ln, sl), cross_point(ln, sr)
                mv.cwidth = 2
                if last:
                    _, pmin = _extreme(mv, [1, 0, 0])
                    lu = with_dir(pmin)
                    mv.contour[2], mv.contour[3] = cross_point(lu, sr), cross_point(lu, sl)
                    mv.cwidth = 4
    out["contour"] = {}          # hand-off plane index -> (width, 4 x 3); -1 = merged ground plane; -2 - k = virtual riser k
    vk = 0
    for m in P:
        if len(m.src) == 1:
            out["contour"][m.src[0]] = (m.cwidth, m.contour)
        elif len(m.src) > 1:
            out["contour"][-1] = (m.cwidth, m.contour)
        else:
            out["contour"][-2 - vk] = (m.cwidth, m.contour)
            vk += 1
    return out
