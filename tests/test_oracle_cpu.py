"""CPU tests of the oracle (test infrastructure): PRNG known answers, fixture facts, stage invariants, the committed
golden digests, and the only external anchor the reference publishes -- the step table of samples/0001
(/root/reference/pic/screenshot2.png)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

TOOLS_TAPS = None


def _canon(a):
    a = np.ascontiguousarray(a)
    if a.dtype.kind == "f":
        a = a.copy()
        a[np.isnan(a)] = np.nan
        a[a == 0] = 0.0
    return a.tobytes()


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "oracle_digests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def oracles(sp, orc, samples):
    out = {}
    for name, (rec, w, h) in samples.items():
        o = orc.Oracle(sp.DEFAULT_PARAMS)
        o.process(rec, w, h)
        out[name] = o
    return out


def test_prng_known_answers(orc):
    """std::mt19937(12345) >> 1 (boost uniform_int(0, INT_MAX)) and glibc rand() from seed 1"""
    assert orc.mt19937_draws(12345, 5).tolist() == [1996335345, 1911592690, 679411342, 280691776, 394962642]
    assert orc.glibc_rand(1, 5).tolist() == [1804289383, 846930886, 1681692777, 1714636915, 1957747793]
    assert orc.glibc_rand(0, 5).tolist() == orc.glibc_rand(1, 5).tolist()      # glibc maps seed 0 to 1


def test_fixture_facts(samples, oracles):
    """SURVEY Appendix C, measured on the recorded frames"""
    facts = {"0000": (153921, 153921, 63167), "0001": (168885, 168884, 48203)}
    for name, (rec, w, h) in samples.items():
        assert (w, h) == (512, 424) and rec.shape == (217088,)
        rgba = rec["rgba"]
        invalid = ((rgba & 255) == 0) | (((rgba >> 8) & 255) == 0) | (((rgba >> 16) & 255) == 0)
        assert int(invalid.sum()) == facts[name][2]
        c = oracles[name].counts()
        assert (c["n_valid"], c["n_crop"]) == facts[name][:2]
        assert np.isfinite(rec["x"]).all()                                     # stale finite xyz under invalid pixels


def test_levelling_matrices(orc):
    """identity IMU pose maps camera (x, y, z) -> (y, z, x) (test/reader.cpp:233-248); Euler vs double-precision formula"""
    R = orc.rotation_from_euler(0, 0, 0).reshape(3, 3)
    assert np.allclose(R, [[0, 1, 0], [0, 0, 1], [1, 0, 0]], atol=1e-7)
    R = orc.rotation_from_quat(1, 0, 0, 0).reshape(3, 3)
    assert np.allclose(R, [[0, 1, 0], [0, 0, 1], [1, 0, 0]], atol=1e-7)
    p, r, y = 26.96, 0.12, 48.8
    a = np.deg2rad([r, p, y])
    Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
    Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
    Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
    m1 = np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]])
    m3 = np.array([[0, 0, -1], [1, 0, 0], [0, -1, 0]])
    assert np.allclose(orc.rotation_from_euler(p, r, y).reshape(3, 3), m3 @ Rx @ Ry @ Rz @ m1, atol=2e-6)
    # quaternion of the same rotation
    from scipy.spatial.transform import Rotation
    q = Rotation.from_matrix(Rx @ Ry @ Rz).as_quat()                            # x, y, z, w
    assert np.allclose(orc.rotation_from_quat(q[3], q[0], q[1], q[2]).reshape(3, 3), m3 @ Rx @ Ry @ Rz @ m1, atol=2e-6)


def test_eigen_solvers_against_numpy(orc):
    rng = np.random.default_rng(5)
    L = orc.lib()
    for _ in range(200):
        A = rng.normal(size=(3, 3)) * rng.uniform(1e-3, 1.0)
        cov = (A @ A.T).astype(np.float32)
        w, v = np.linalg.eigh(cov.astype(np.float64))
        ev = np.zeros(1, np.float32)
        vec = np.zeros(3, np.float32)
        L.orc_eigen33(cov.ctypes.data, 0, ev.ctypes.data, vec.ctypes.data)
        assert abs(ev[0] - w[0]) <= 2e-5 * max(1e-6, w[2])
        if (w[1] - w[0]) > 1e-3 * w[2]:
            assert abs(abs(float(vec @ v[:, 0])) - 1.0) < 1e-3
        evals = np.zeros(3, np.float32)
        evecs = np.zeros(9, np.float32)
        L.orc_jacobi3(cov.ctypes.data, evals.ctypes.data, evecs.ctypes.data)
        assert np.allclose(evals, w[::-1], rtol=1e-4, atol=1e-6 * w[2])


def test_stage_invariants(sp, oracles, samples):
    """properties that hold whatever the arithmetic: order preservation, membership, label consistency"""
    for name, o in oracles.items():
        c = o.counts()
        xyz, nrm = o.tap("xyz"), o.tap("normals")
        key, vidx = o.tap("voxel_key"), o.tap("voxel_idx")
        assert vidx.size == c["n_voxel"] == np.unique(key[key != 0xFFFFFFFF]).size
        assert (np.diff(key[vidx].astype(np.int64)) > 0).all()                 # output in ascending voxel key, one per voxel
        nonan, noleg = o.tap("nonan_idx"), o.tap("noleg_idx")
        assert np.isin(nonan, vidx).all() and np.isfinite(nrm[nonan]).all()
        assert np.isin(noleg, nonan).all()
        th = np.float32(sp.DEFAULT_PARAMS["noleg_distance"])
        assert ((xyz[noleg].astype(np.float64) ** 2).sum(1) > float(th) ** 2 * (1 - 1e-6)).all()
        h, v = o.tap("h_idx"), o.tap("v_idx")
        assert np.intersect1d(h, v).size == 0 and np.isin(h, noleg).all() and np.isin(v, noleg).all()
        t1 = np.deg2rad(sp.DEFAULT_PARAMS["parallel_angle_diff"])
        ang = np.arccos(np.clip(nrm[h, 0].astype(np.float64), -1, 1))
        assert (np.minimum(ang, np.pi - ang) < t1 + 1e-6).all()
        # unit normals, flipped toward the sensor origin
        n = nrm[nonan].astype(np.float64)
        assert np.allclose((n ** 2).sum(1), 1.0, atol=1e-5)
        assert ((n * -xyz[nonan]).sum(1) >= -1e-6).all()
        # clusters: same root <=> connected within tolerance is too costly to brute-force on 15k points; check roots are members
        for tag, idx in (("h", h), ("v", v)):
            root = o.tap(tag + "_root")
            assert root.size == idx.size
            ok = root >= 0
            assert (root[ok] <= np.flatnonzero(ok)).all()                       # root = smallest member position
            assert (root[root[ok]] == root[ok]).all()
        # planes partition a subset of H / V points; inliers within threshold of their RANSAC plane
        pidx = o.tap("plane_idx")
        assert pidx.size == c["n_plane_points"] == np.unique(pidx).size
        for tag, idx in (("h", h), ("v", v)):
            coef, cnt, sidx = o.tap(tag + "_seg_coef"), o.tap(tag + "_seg_count"), o.tap(tag + "_seg_idx")
            assert np.isin(sidx, idx).all() and cnt.sum() == sidx.size
            off = 0
            for k in range(cnt.size):
                p = xyz[sidx[off:off + cnt[k]]].astype(np.float64)
                dist = np.abs(p @ coef[k, :3].astype(np.float64) + coef[k, 3])
                assert dist.max() < sp.DEFAULT_PARAMS["seg_threshold"] * (1 + 1e-4)
                assert cnt[k] >= sp.DEFAULT_PARAMS["seg_rest_point"]
                off += cnt[k]


def test_cluster_labels_bruteforce(sp, orc, synth):
    """radius-graph components against an O(n^2) numpy restatement on a thinned frame"""
    rec, R = synth.make_frame(2)
    p = dict(sp.DEFAULT_PARAMS)
    p.update(voxel_x=0.04, voxel_y=0.04, voxel_z=0.04, cluster_tolerance=0.09, seg_rest_point=60)
    o = orc.Oracle(p)
    o.process(rec, 512, 424, R)
    xyz = o.tap("xyz")
    for tag in ("h", "v"):
        idx, root = o.tap(tag + "_idx"), o.tap(tag + "_root")
        n = idx.size
        assert 50 < n < 6000
        P = xyz[idx]
        d = P[:, None, :] - P[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]       # float32, flann::L2_Simple order
        r2 = np.float32(np.float64(np.float32(p["cluster_tolerance"])) ** 2)
        from scipy.sparse.csgraph import connected_components
        from scipy.sparse import csr_matrix
        ncomp, lab = connected_components(csr_matrix(d2 < r2), directed=False)
        first = np.full(ncomp, n, dtype=np.int64)
        np.minimum.at(first, lab, np.arange(n))
        size = np.bincount(lab, minlength=ncomp)
        want = np.where((size[lab] >= p["min_cluster_size"]) & (size[lab] <= p["max_cluster_size"]), first[lab], -1)
        assert np.array_equal(root, want)


def test_ransac_first_draw_rule(sp, oracles):
    """the first scored triple of a segment() call on an n-point cluster follows the index-shuffle sampler driven by
    mt19937(12345): row layout of the call log = cluster, call, n, scored, i0, i1, i2, best, inliers, emitted"""
    for o in oracles.values():
        for tag in ("h", "v"):
            log = o.tap(tag + "_log")
            assert log.shape[1] == 10 and (log[:, 3] <= 101).all() and (log[:, 3] >= 1).all()
            assert (log[:, 7] <= log[:, 2]).all()
            big = log[log[:, 9] == 1]
            assert (big[:, 8] >= sp.DEFAULT_PARAMS["seg_rest_point"]).all()


def _line_of_two_planes(c1, c2):
    """computeLineFrom2Planes (StairPerception.cpp:1096-1142): h = x of the intersection line's point closest to the
    origin, d = sqrt(y^2 + z^2) of that point"""
    n1, n2 = c1[:3].astype(np.float64), c2[:3].astype(np.float64)
    pt = np.linalg.solve(np.vstack([n1, n2, np.cross(n1, n2)]), [-float(c1[3]), -float(c2[3]), 0.0])
    return pt[0], float(np.hypot(pt[1], pt[2]))


def test_screenshot_anchor_sample_0001(oracles):
    """External anchor -- the only numbers the reference publishes for this path: pic/screenshot2.png lists, for
    samples/0001, four steps {Height, Depth, V_Distance, H_Distance}.  V/H_Distance are the (h, d) of
    the riser/tread intersection line of each step; Height_k = |h_step_k - h_concave_(k-1)|, Depth_k = |d_concave_k -
    d_step_k| (StairPerception.cpp:2622-2767).  Recomputed here from the planes the front-end hands over: every line agrees
    within 2 mm, every height/depth that needs only plane coefficients within 1 mm (north_star tolerance)."""
    published = [(0.110046, 0.280456, 0.490394, 0.526096), (0.110491, 0.276792, 0.397571, 0.807455),
                 (0.111913, 0.269715, 0.300275, 1.08765), (0.104439, 0.283345, 0.212566, 1.35878)]
    pl = oracles["0001"].tap("planes")
    H = sorted((p for p in pl if p["type"] == 1 and abs(p["center"][2]) < 0.3 and p["center"][1] < 1.7), key=lambda p: -p["center"][0])
    V = sorted((p for p in pl if p["type"] == 0 and abs(p["coef"][1]) > 0.98 and abs(p["coef"][3]) < 1.7), key=lambda p: abs(p["coef"][3]))
    assert len(H) == 4 and len(V) == 5
    step = [_line_of_two_planes(V[k]["coef"], H[k]["coef"]) for k in range(4)]           # V -> H: step edge
    concave = [_line_of_two_planes(H[k]["coef"], V[k + 1]["coef"]) for k in range(4)]    # H -> V: concave line
    for k in range(4):
        assert abs(step[k][0] - published[k][2]) < 2e-3 and abs(step[k][1] - published[k][3]) < 2e-3, (k, step[k])
        assert abs(abs(concave[k][1] - step[k][1]) - published[k][1]) < 1e-3, ("depth", k)
        if k > 0:
            assert abs(abs(step[k][0] - concave[k - 1][0]) - published[k][0]) < 1e-3, ("height", k)


def test_golden_digests(sp, orc, synth, samples, gold):
    """the oracle reproduces the committed goldens (tools/make_golden.py) bit for bit"""
    todo = [("sample_0000", samples["0000"][0], None), ("sample_0001", samples["0001"][0], None)]
    for fid in (3, 11):
        rec, R = synth.make_frame(fid)
        todo.append(("synth_%d" % fid, rec, R))
    for name, rec, R in todo:
        g = gold["frames"][name]
        sha = hashlib.sha256(rec.tobytes() + (b"" if R is None else np.asarray(R, dtype=np.float32).tobytes())).hexdigest()
        if sha != g["input_sha256"]:
            assert name.startswith("synth"), "recorded fixture changed"
            pytest.skip("numpy produced a different synthetic frame than the one the goldens were made from")
        o = orc.Oracle(sp.DEFAULT_PARAMS)
        o.process(rec, 512, 424, R)
        assert o.counts() == g["counts"], name
        for t, want in g["taps"].items():
            a = o.tap(t)
            assert a.size == want["n"] and hashlib.sha256(_canon(a)).hexdigest() == want["sha256"], (name, t)


def test_faithful_normals_mode_is_close(sp, orc, samples):
    """normal_arith=0 (PCL-style integral images, double accumulators) vs the portable box-sum definition the CUDA kernel
    reproduces: same NaN pattern, normals equal to float rounding"""
    rec, w, h = samples["0000"]
    a = orc.Oracle(sp.DEFAULT_PARAMS, normal_arith=0)
    b = orc.Oracle(sp.DEFAULT_PARAMS, normal_arith=1)
    a.process(rec, w, h, show_mode=1)
    b.process(rec, w, h, show_mode=1)
    na, nb = a.tap("normals"), b.tap("normals")
    assert np.array_equal(np.isnan(na), np.isnan(nb))
    ok = ~np.isnan(na)
    assert np.abs(na[ok] - nb[ok]).max() < 1e-5


def test_show_mode_original_does_nothing(sp, orc, samples):
    rec, w, h = samples["0000"]
    o = orc.Oracle(sp.DEFAULT_PARAMS)
    assert o.process(rec, w, h, show_mode=0) == 0


def test_determinism_and_thread_batch(sp, orc, synth):
    recs, Rs = synth.make_batch(4)
    tot1, c1 = orc.process_batch(sp.DEFAULT_PARAMS, recs, 512, 424, Rs, nthreads=1)
    tot4, c4 = orc.process_batch(sp.DEFAULT_PARAMS, recs, 512, 424, Rs, nthreads=4)
    assert tot1 == tot4 and np.array_equal(c1, c4) and (c1 >= 8).all()


PUBLISHED_0001 = [(0.110046, 0.280456, 0.490394, 0.526096), (0.110491, 0.276792, 0.397571, 0.807455),
                  (0.111913, 0.269715, 0.300275, 1.08765), (0.104439, 0.283345, 0.212566, 1.35878)]


def test_stair_model_reproduces_published_step_table(sp, oracles):
    """The oracle's stair model (oracle/stair_model.py, a restatement of stairModel, StairPerception.cpp:1621-2773) on the
    oracle's planes of samples/0001 finds the four steps of pic/screenshot2.png: 14 of the 16 published numbers within 0.5 mm,
    all within 1.6 mm (the first step hangs on a 330-point riser)."""
    from oracle import stair_model as sm
    o = oracles["0001"]
    res = sm.stair_model(sp.DEFAULT_PARAMS, o.tap("planes"), o.tap("xyz"), o.tap("plane_idx"))
    assert res["has_stair"] and len(res["steps"]) == 4
    err = np.abs(np.array(res["steps"]) - np.array(PUBLISHED_0001))
    assert err.max() < 1.6e-3 and (err < 0.5e-3).sum() >= 14, err
    assert err[1:, :2].max() < 1e-3                    # heights and depths of steps 2-4: the 1 mm north-star tolerance
    o0 = oracles["0000"]
    res0 = sm.stair_model(sp.DEFAULT_PARAMS, o0.tap("planes"), o0.tap("xyz"), o0.tap("plane_idx"))
    assert res0["has_stair"] and len(res0["steps"]) == 2


def test_stair_contours_are_consistent(sp, oracles):
    """computePlaneCounter (StairPerception.cpp:2861-3075) restated in the oracle: every closed contour of samples/0001 has its
    corners 0 / 3 on the left side plane and 1 / 2 on the right one, all four within the RANSAC threshold of its own plane, and
    the two side planes are parallel and about one stair width apart."""
    from oracle import stair_model as sm
    o = oracles["0001"]
    planes = o.tap("planes")
    res = sm.stair_model(sp.DEFAULT_PARAMS, planes, o.tap("xyz"), o.tap("plane_idx"))
    sl, sr = res["slide_left"].astype(np.float64), res["slide_right"].astype(np.float64)
    assert np.allclose(sl[:3], sr[:3]) and abs(np.linalg.norm(sl[:3]) - 1) < 1e-5
    width = abs(sl[3] - sr[3])
    assert 0.5 < width < 2.0, width
    closed = 0
    for k, (w, pts) in res["contour"].items():
        if w != 4:
            continue
        closed += 1
        p = pts.astype(np.float64)
        for c, side in ((0, sl), (3, sl), (1, sr), (2, sr)):
            assert abs(p[c] @ side[:3] + side[3]) < 1e-4, (k, c)
        if k >= 0:
            coef = planes[k]["coef"].astype(np.float64)
            assert np.abs(p @ coef[:3] + coef[3]).max() < 0.03, (k, np.abs(p @ coef[:3] + coef[3]))
            assert abs(np.linalg.norm(p[1] - p[0]) - width) < 0.02 and abs(np.linalg.norm(p[2] - p[3]) - width) < 0.02
    assert closed >= 7                                   # four treads' risers + treads of a four-step stair (the ground may stay open)


def test_stair_model_on_synthetic_ground_truth(sp, orc, synth):
    """synthetic frames have a known staircase: rise 0.11 m, run 0.28 m"""
    from oracle import stair_model as sm
    for fid in (0, 3):
        rec, R = synth.make_frame(fid)
        o = orc.Oracle(sp.DEFAULT_PARAMS)
        o.process(rec, 512, 424, R)
        res = sm.stair_model(sp.DEFAULT_PARAMS, o.tap("planes"), o.tap("xyz"), o.tap("plane_idx"))
        assert res["has_stair"] and len(res["steps"]) >= 2
        steps = np.array(res["steps"])
        assert np.abs(steps[:, 0] - 0.11).max() < 3e-3
        assert np.abs(steps[:-1, 1] - 0.28).max() < 3e-3


def test_shared_numerics_audited_against_independent_ones(sp, orc, samples, synth):
    """include/sp_detmath.h (eigen33, Jacobi, acos, log) is compiled into the product AND into the oracle's default mode, so the
    bit-exact parity tests cannot see a wrong restatement in it.  Two independent checks (oracle/indep_math.h: classical double
    Jacobi + libm, sharing nothing with that header):
      - the audit re-solves every eigen problem / acos / log of a frame and bounds the deviation of the shared-header result;
      - eig_math = 2 runs the whole pipeline on the independent numerics: same planes, coefficients within 1e-4 (the stated
        tolerance for plane coefficients), centres identical."""
    cases = [(samples["0000"][0], None), (samples["0001"][0], None)] + [synth.make_frame(f) for f in (2, 9)]
    for rec, R in cases:
        o = orc.Oracle(sp.DEFAULT_PARAMS)
        o.process(rec, 512, 424, R)
        a = o.audit()
        assert a["n_eigen33"] >= 5 and a["n_jacobi"] >= 5 and a["n_log"] >= 10
        assert a["eigen33_max_angle_wellsep"] < 1e-5 and a["eigen33_max_val"] < 1e-4
        assert a["jacobi_max_val"] < 1e-5 and a["jacobi_max_angle"] < 1e-4
        assert a["acos_max_abs"] < 1e-6 and a["log_max_rel"] < 1e-13
        o2 = orc.Oracle(sp.DEFAULT_PARAMS, eig_math=2)
        o2.process(rec, 512, 424, R)
        p0, p2 = o.tap("planes"), o2.tap("planes")
        assert len(p0) == len(p2) and len(p0) >= 6
        for q0, q2 in zip(p0, p2):
            sg = np.sign(np.dot(q0["coef"][:3], q2["coef"][:3]))
            assert int(q0["count"]) == int(q2["count"])
            assert np.abs(q0["coef"] - sg * q2["coef"]).max() < 1e-4
            assert np.abs(q0["center"] - q2["center"]).max() < 1e-6
            assert np.allclose(q0["eval"], q2["eval"], rtol=1e-3, atol=1e-9)
