"""GPU-vs-oracle comparison used by the -m gpu parity tests and by tools/."""
import numpy as np

BIT_EXACT_TAPS = ["xyz", "normals", "minmax", "voxel_key", "voxel_idx", "nonan_idx", "noleg_idx", "h_idx", "v_idx",
                  "h_root", "v_root", "h_log", "v_log", "h_seg_coef", "h_seg_count", "h_seg_idx", "v_seg_coef", "v_seg_count",
                  "v_seg_idx", "h_merge_coef", "h_merge_count", "h_merge_idx", "v_merge_coef", "v_merge_count", "v_merge_idx",
                  "plane_idx"]

STAGE_OF_TAP = {  # lowest show_mode at which the tap exists
    "xyz": 1, "normals": 1, "minmax": 2, "voxel_key": 2, "voxel_idx": 2, "nonan_idx": 2, "noleg_idx": 3, "h_idx": 4, "v_idx": 4,
}


def same_bits(a, b):
    """equal including the NaN pattern (NaN payloads ignored), -0.0 == +0.0"""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        na, nb = np.isnan(a), np.isnan(b)
        return bool(np.array_equal(na, nb) and np.array_equal(a[~na], b[~nb]))
    return bool(np.array_equal(a, b))


def compare_frame(ctx, frame, res, oracle, taps=BIT_EXACT_TAPS, show_mode=10):
    """returns a list of mismatch descriptions (empty = parity)"""
    bad = []
    oc = oracle.counts()
    for k, v in oc.items():
        if int(res[k]) != v:
            bad.append("count %s: gpu %d oracle %d" % (k, int(res[k]), v))
    for name in taps:
        o = oracle.tap(name)
        if o is None:
            continue
        g = ctx.tap(frame, name)
        if name.endswith("_log"):
            # the kernel never launches segment() on clusters that cannot yield a plane (size < seg_rest_point);
            # the reference's vertical loop does (and gets < seg_rest_point inliers back): drop those rows
            o = o[o[:, 2] >= max(3, int(oracle.p.seg_rest_point))]
        if name in ("h_root", "v_root", "h_idx", "v_idx") and o.size == 0 and g.size == 0:
            continue
        if not same_bits(g.reshape(-1), o.reshape(-1)):
            n = min(g.size, o.size)
            gd, od = g.reshape(-1)[:n], o.reshape(-1)[:n]
            if gd.dtype.kind == "f":
                diff = ~((gd == od) | (np.isnan(gd) & np.isnan(od)))
            else:
                diff = gd != od
            bad.append("tap %s: sizes %d/%d, %d differing, first at %s" % (name, g.size, o.size, int(diff.sum()),
                                                                         np.flatnonzero(diff)[:5].tolist()))
    op = oracle.tap("planes")
    if op is not None:
        n = int(res["n_planes"])
        if n != len(op):
            bad.append("planes: %d vs %d" % (n, len(op)))
        else:
            for i in range(n):
                for fld in op.dtype.names:
                    if not same_bits(res["planes"][i][fld], op[i][fld]):
                        bad.append("plane %d field %s: gpu %s oracle %s" % (i, fld, res["planes"][i][fld], op[i][fld]))
    return bad
