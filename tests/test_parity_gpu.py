"""Parity of the CUDA path (through the C ABI) against the CPU oracle: bit-exact for every integer output (voxel
membership, representative picks, H/V labels, cluster labels, inlier masks, plane labels) and -- because the kernels
evaluate the same un-fused float expressions in the same order -- for the float outputs too."""
import numpy as np
import pytest

from parity import compare_frame, same_bits, BIT_EXACT_TAPS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx1(sp):
    c = sp.Context(device=0, width=512, height=424, max_batch=1)
    yield c
    c.close()


@pytest.mark.parametrize("name", ["0000", "0001"])
def test_samples_bit_exact(sp, orc, samples, ctx1, name):
    rec, w, h = samples[name]
    res = ctx1.process_frames(rec)[0]
    o = orc.Oracle(sp.DEFAULT_PARAMS)
    o.process(rec, w, h)
    bad = compare_frame(ctx1, 0, res, o)
    assert not bad, "\n".join(bad[:20])
    assert int(res["n_planes"]) > 0


@pytest.mark.parametrize("fid", [3, 11, 42])
def test_synthetic_with_levelling(sp, orc, synth, ctx1, fid):
    rec, R = synth.make_frame(fid)
    res = ctx1.process_frames(rec, R.reshape(1, 9))[0]
    o = orc.Oracle(sp.DEFAULT_PARAMS)
    o.process(rec, 512, 424, R)
    bad = compare_frame(ctx1, 0, res, o)
    assert not bad, "\n".join(bad[:20])
    assert int(res["n_planes"]) >= 8


def test_batch_equals_single_and_chunking(sp, orc, synth, samples):
    recs = [samples["0000"][0], samples["0001"][0]]
    Rs = [np.eye(3, dtype=np.float32).reshape(9)] * 2
    for fid in (1, 2, 5):
        r, R = synth.make_frame(fid)
        recs.append(r)
        Rs.append(R)
    clouds = np.stack(recs)
    R = np.stack(Rs)
    ctx = sp.Context(device=0, width=512, height=424, max_batch=3)      # 5 frames -> chunks of 3 + 2
    res = ctx.process_frames(clouds, R)
    assert len(res) == 5
    for i in (3, 4):                                                     # last chunk holds frames 3, 4
        o = orc.Oracle(sp.DEFAULT_PARAMS)
        o.process(recs[i], 512, 424, Rs[i])
        bad = compare_frame(ctx, i - 3, res[i], o)
        assert not bad, "frame %d\n" % i + "\n".join(bad[:20])
    for i in (0, 1, 2):
        o = orc.Oracle(sp.DEFAULT_PARAMS)
        o.process(recs[i], 512, 424, Rs[i])
        oc = o.counts()
        for k, v in oc.items():
            assert int(res[i][k]) == v, (i, k)
        op = o.tap("planes")
        for j in range(len(op)):
            for fld in op.dtype.names:
                assert same_bits(res[i]["planes"][j][fld], op[j][fld]), (i, j, fld)
    ctx.close()


def test_device_resident_input(sp, orc, synth):
    import torch
    rec, R = synth.make_frame(9)
    ctx = sp.Context(device=0, width=512, height=424, max_batch=2)
    dcloud = torch.from_numpy(np.stack([rec, rec]).view(np.uint8).reshape(2, -1)).cuda()
    dR = torch.from_numpy(np.stack([R, R])).cuda()
    torch.cuda.synchronize()
    res = ctx.process_frames(dcloud, dR, device_input=True, nframes=2)
    o = orc.Oracle(sp.DEFAULT_PARAMS)
    o.process(rec, 512, 424, R)
    for f in (0, 1):
        bad = compare_frame(ctx, f, res[f], o)
        assert not bad, "\n".join(bad[:20])
    ctx.close()


def test_edge_cases(sp, orc, synth, ctx1):
    rec, R = synth.make_frame(4)
    # empty frame: every pixel invalid -> "<100 points" guard (StairPerception.cpp:56)
    e = rec.copy()
    e["rgba"] = 0
    res = ctx1.process_frames(e, R.reshape(1, 9))[0]
    assert int(res["status"]) == 1 and int(res["n_planes"]) == 0 and int(res["n_valid"]) == 0
    o = orc.Oracle(sp.DEFAULT_PARAMS)
    o.process(e, 512, 424, R)
    assert o.counts()["status"] == 1
    # ragged: only 150 valid pixels in one corner patch
    e = rec.copy()
    keep = np.zeros((424, 512), dtype=bool)
    keep[200:210, 100:115] = True
    e["rgba"] = np.where(keep.reshape(-1), np.uint32(0xFF808080), np.uint32(0))
    res = ctx1.process_frames(e, R.reshape(1, 9))[0]
    o.process(e, 512, 424, R)
    bad = compare_frame(ctx1, 0, res, o)
    assert not bad, "\n".join(bad[:20])
    # NaN coordinates in the input with valid colour
    e = rec.copy()
    e["x"][1000:5000] = np.nan
    res = ctx1.process_frames(e, R.reshape(1, 9))[0]
    o.process(e, 512, 424, R)
    bad = compare_frame(ctx1, 0, res, o)
    assert not bad, "\n".join(bad[:20])


@pytest.mark.parametrize("mode", [1, 2, 3, 4, 6, 8, 9])
def test_show_mode_taps(sp, orc, synth, ctx1, mode):
    """process() returns early at every ShowModeDef tap (StairPerception.cpp:49-176)"""
    rec, R = synth.make_frame(6)
    res = ctx1.process_frames(rec, R.reshape(1, 9), show_mode=mode)[0]
    o = orc.Oracle(sp.DEFAULT_PARAMS)
    o.process(rec, 512, 424, R, show_mode=mode)
    assert int(res["n_planes"]) == 0
    taps = {1: ["xyz", "normals"], 2: ["voxel_idx", "nonan_idx"], 3: ["noleg_idx"], 4: ["h_idx", "v_idx"],
            6: ["h_seg_coef", "h_seg_idx"], 8: ["h_merge_coef", "h_merge_idx"], 9: ["v_merge_coef", "v_merge_idx"]}[mode]
    for t in taps:
        assert same_bits(ctx1.tap(0, t).reshape(-1), o.tap(t).reshape(-1)), t


def test_other_parameters(sp, orc, synth):
    """non-default knobs: tighter crop, coarser leaf, different thresholds"""
    rec, R = synth.make_frame(8)
    p = dict(sp.DEFAULT_PARAMS)
    p.update(x_min=0.2, x_max=1.2, y_min=0.3, y_max=2.5, z_min=-0.6, z_max=0.6, voxel_x=0.02, voxel_y=0.015, voxel_z=0.02,
             cluster_tolerance=0.05, seg_rest_point=120, seg_threshold=0.015, noleg_distance=0.5, parallel_angle_diff=15.0,
             perpendicular_angle_diff=20.0, seg_max_iters=50)
    ctx = sp.Context(p, device=0, width=512, height=424, max_batch=1)
    res = ctx.process_frames(rec, R.reshape(1, 9))[0]
    o = orc.Oracle(p)
    o.process(rec, 512, 424, R)
    bad = compare_frame(ctx, 0, res, o)
    assert not bad, "\n".join(bad[:20])
    ctx.close()


def test_facade(sp, samples):
    det = sp.StairDetection(sp.DEFAULT_PARAMS)
    rec, w, h = samples["0001"]
    has, planes, res = det.process(rec, "stair", width=w, height=h)
    assert has and len(planes) == int(res["n_planes"])
    for p in planes:
        assert p["xyz"].shape == (int(p["record"]["count"]), 3)
    has, planes, _ = det.process(rec, "original", width=w, height=h)
    assert not has


def test_cpp_facade_matches_oracle(sp, orc, samples):
    """examples/files_mode (C++ facade over the C ABI) hands over the same sorted planes as the oracle"""
    import os
    import subprocess
    from conftest import ROOT
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "-s"])
    for name in ("0000", "0001"):
        r = subprocess.run([os.path.join(ROOT, "examples", "files_mode"), os.path.join(ROOT, "tests", "golden", "samples", name + "_cloud.pcd")],
                           capture_output=True, text=True, check=True)
        lines = r.stdout.strip().splitlines()
        rec, w, h = samples[name]
        o = orc.Oracle(sp.DEFAULT_PARAMS)
        o.process(rec, w, h)
        op = o.tap("planes")
        assert lines[0].split()[:4] == ["has_planes", "1", "planes", str(len(op))]
        for ln, p in zip(lines[1:], op):
            t = ln.split()
            assert int(t[0]) == int(p["type"]) and int(t[1]) == int(p["count"])
            got = np.array([float.fromhex(v) for v in t[2:9]], dtype=np.float32)
            assert same_bits(got, np.concatenate([p["coef"], p["center"]]))
        if name == "0001":                                   # four steps: the contours of their treads and risers are closed
            assert any(l.startswith("side planes:") for l in lines) and sum(l.startswith("contour plane") for l in lines) >= 7
    # a tap through the facade: horizontal cloud
    r = subprocess.run([os.path.join(ROOT, "examples", "files_mode"), os.path.join(ROOT, "tests", "golden", "samples", "0000_cloud.pcd"), "4"],
                       capture_output=True, text=True, check=True)
    assert r.stdout.split()[:6] == ["has_planes", "0", "planes", "1", "cloud_show", str(int(o.counts()["n_h"]) if False else 10679)]


@pytest.mark.parametrize("n_points,seed", [(150_000, 1), (60_000, 2)])
def test_knn_normals_unorganised_parity(sp, orc, synth, n_points, seed):
    """a3' (normalEstimationOMP, StairPerception.cpp:524-541): unorganised cloud, kNN covariance normals (k = 69), then the
    same pipeline.  Neighbour lists in (distance, index) order make every float sum order-identical: bit-exact taps."""
    rec = synth.make_dense_scene(n_points, seed=seed)
    rec["x"][100:140] = np.nan                                   # a few non-finite points with valid colour
    rec[-1]["x"], rec[-1]["y"], rec[-1]["z"] = 3.0, 9.0, 4.0     # an isolated point: its neighbours are far away
    p = dict(sp.DEFAULT_PARAMS)
    p.update(synth.DENSE_PARAMS)
    ctx = sp.Context(p, device=0, width=n_points, height=1, max_batch=1)
    ctx.set_normal_mode(1)
    res = ctx.process_frames(rec)[0]
    o = orc.Oracle(p, knn_threads=16)
    o.process(rec, n_points, 1, normal_mode=1)
    bad = compare_frame(ctx, 0, res, o)
    assert not bad, "\n".join(bad[:20])
    assert int(res["status"]) == 0 and int(res["n_planes"]) >= 8
    nrm = ctx.tap(0, "normals")
    ok = np.isfinite(nrm[:, 0])
    assert ok.sum() == int(res["n_valid"])
    assert np.allclose((nrm[ok].astype(np.float64) ** 2).sum(1), 1.0, atol=1e-4)
    ctx.close()


def test_knn_config4_two_million_points(sp, orc, synth):
    """BASELINE.json config 4 at full size: 2 M-point dense scene, leaf 0.005, kNN normals; GPU vs oracle on every count,
    the normals and the final planes, plus size-independent properties (treads 0.11 m apart, risers 0.28 m apart)."""
    import os
    n_points = 2_000_000
    rec = synth.make_dense_scene(n_points, seed=0)
    p = dict(sp.DEFAULT_PARAMS)
    p.update(synth.DENSE_PARAMS)
    ctx = sp.Context(p, device=0, width=n_points, height=1, max_batch=1)
    ctx.set_normal_mode(1)
    res = ctx.process_frames(rec)[0]
    assert int(res["status"]) == 0
    planes = res["planes"][:int(res["n_planes"])]
    treads = sorted(-float(q["coef"][3]) for q in planes if q["type"] == 1 and 15000 < q["count"] < 40000)
    risers = sorted(float(q["coef"][3]) for q in planes if q["type"] == 0 and 5000 < q["count"] < 12000)
    assert len(treads) >= 3 and np.allclose(np.diff(treads), 0.11, atol=3e-3)
    assert len(risers) >= 4 and np.allclose(np.diff(risers), 0.28, atol=5e-3)
    o = orc.Oracle(p, knn_threads=max(8, len(os.sched_getaffinity(0))))
    o.process(rec, n_points, 1, normal_mode=1)
    bad = compare_frame(ctx, 0, res, o, taps=["normals", "voxel_idx", "h_idx", "v_idx", "h_root", "v_root", "plane_idx"])
    assert not bad, "\n".join(bad[:20])
    ctx.close()


def test_minmax_proj_on_device(sp, orc, samples, ctx1):
    """a15 findMinMaxProjPoint (StairPerception.cpp:2780-2808) as an on-device reduction over a plane's point list"""
    rec, w, h = samples["0001"]
    res = ctx1.process_frames(rec)[0]
    o = orc.Oracle(sp.DEFAULT_PARAMS)
    o.process(rec, w, h)
    rng = np.random.default_rng(1)
    for plane in range(int(res["n_planes"])):
        for v in ([1, 0, 0], [0, 1, 0], [0, 0, -1], rng.normal(size=3)):
            g = ctx1.minmax_proj(0, plane, v)
            e = o.minmax_proj(plane, v)
            assert same_bits(g[0], e[0]) and same_bits(g[1], e[1]) and g[2:] == e[2:], (plane, v)
        g = ctx1.minmax_proj(0, plane, [0, 1, 0], center=[0.1, 0.2, 0.3])
        e = o.minmax_proj(plane, [0, 1, 0], center=[0.1, 0.2, 0.3])
        assert same_bits(g[0], e[0]) and g[2:] == e[2:]


@pytest.mark.parametrize("dtype,with_colour", [(np.uint16, True), (np.float32, True), (np.uint16, False)])
def test_depth_image_ingest(sp, orc, synth, dtype, with_colour):
    """next-row f2: K2G::updateCloud (k2g.cpp:231-308) on the device in front of the path.  The records the kernel builds from
    (depth, colour, intrinsics, mirror) equal the oracle's back-projection bit for bit, and so does everything downstream."""
    ctx = sp.Context(device=0, width=512, height=424, max_batch=2)
    frames = [synth.make_depth_frame(fid) for fid in (21, 22, 23)]
    depth = np.stack([f[0] for f in frames]).astype(dtype)
    if dtype == np.float32:
        depth[0, 5, 7] = np.nan                         # libfreenect2 marks some invalid pixels NaN instead of 0
    bgrx = np.stack([f[1] for f in frames]) if with_colour else None
    R = np.stack([f[2] for f in frames])
    res = ctx.process_depth_frames(depth, synth.INTRINSICS, bgrx=bgrx, mirror=True, R=R)
    assert len(res) == 3
    i = 2                                               # the last chunk (frames 2) still sits in the context
    rec = orc.depth_to_cloud(depth[i], synth.INTRINSICS, None if bgrx is None else bgrx[i], mirror=True)
    o = orc.Oracle(sp.DEFAULT_PARAMS)
    o.process(rec, 512, 424, R[i])
    bad = compare_frame(ctx, 0, res[i], o)
    assert not bad, "\n".join(bad[:20])
    assert int(res[i]["n_planes"]) >= 8
    for j in (0, 1):
        rec = orc.depth_to_cloud(depth[j], synth.INTRINSICS, None if bgrx is None else bgrx[j], mirror=True)
        o.process(rec, 512, 424, R[j])
        for k, v in o.counts().items():
            assert int(res[j][k]) == v, (j, k)
    ctx.close()


def test_stair_model_matches_oracle_and_published_table(sp, orc, synth, samples, ctx1):
    """next-row f1: sp_stair_model (stairModel restated in C++ over the hand-off records, plane clouds left on the device)
    against the independent numpy restatement in oracle/stair_model.py, and against pic/screenshot2.png for samples/0001"""
    from oracle import stair_model as sm
    cases = [(samples["0001"][0], None), (samples["0000"][0], None)] + [synth.make_frame(f) for f in (0, 3, 7)]
    for rec, R in cases:
        res = ctx1.process_frames(rec, None if R is None else R.reshape(1, 9))[0]
        st = ctx1.stair_model(0)
        o = orc.Oracle(sp.DEFAULT_PARAMS)
        o.process(rec, 512, 424, R)
        want = sm.stair_model(sp.DEFAULT_PARAMS, o.tap("planes"), o.tap("xyz"), o.tap("plane_idx"))
        got = sp.stair_steps(st)
        assert bool(st["has_stair"]) == bool(want["has_stair"]) and len(got) == len(want["steps"])
        assert np.allclose(np.array(got), np.array(want["steps"]), rtol=0, atol=2e-6), (got, want["steps"])
        assert list(st["ptype"][:int(res["n_planes"])]) == want["ptype"]
        if want["has_stair"]:
            # the side planes of step 8 and computePlaneCounter's contour of every plane (what the GUI draws)
            assert np.allclose(st["slide_left"], want["slide_left"], rtol=0, atol=1e-5) and np.allclose(st["slide_right"], want["slide_right"], rtol=0, atol=1e-5)
            assert int(st["n_virtual"]) == sum(1 for k in want["contour"] if k <= -2)
            closed = 0
            for k, (w, pts) in want["contour"].items():
                gw, gp = ((st["contour_width"][k], st["contour"][k]) if k >= 0 else
                          (st["ground_contour_width"], st["ground_contour"]) if k == -1 else
                          (st["virtual_contour_width"][-2 - k], st["virtual_contour"][-2 - k]))
                assert int(gw) == w, (k, int(gw), w)
                assert np.allclose(gp, pts, rtol=0, atol=2e-5), (k, gp, pts)
                closed += w == 4
            assert closed >= 2 * len(want["steps"]) - 1
    res = ctx1.process_frames(samples["0001"][0])[0]
    got = np.array(sp.stair_steps(ctx1.stair_model(0)))
    published = np.array([(0.110046, 0.280456, 0.490394, 0.526096), (0.110491, 0.276792, 0.397571, 0.807455),
                          (0.111913, 0.269715, 0.300275, 1.08765), (0.104439, 0.283345, 0.212566, 1.35878)])
    assert got.shape == (4, 4) and np.abs(got - published).max() < 1.6e-3 and np.abs(got - published)[1:, :2].max() < 1e-3


@pytest.mark.parametrize("w,h", [(200, 150), (97, 61), (33, 13), (512, 11)])
def test_other_image_sizes(sp, orc, samples, w, h):
    """the path is not tied to 512 x 424: windows cut from a recorded frame, widths that are not a multiple of the tile, an
    image too small for any normal (PCL needs more than 10 rows), all against the oracle"""
    rec, W0, H0 = samples["0001"]
    img = rec.reshape(H0, W0)
    c0 = min(200, W0 - w)
    sub = np.ascontiguousarray(img[120:120 + h, c0:c0 + w]).reshape(-1)
    p = dict(sp.DEFAULT_PARAMS)
    p.update(seg_rest_point=60, min_cluster_size=5)
    ctx = sp.Context(p, device=0, width=w, height=h, max_batch=2)
    res = ctx.process_frames(np.stack([sub, sub]))
    o = orc.Oracle(p)
    o.process(sub, w, h)
    for f in (0, 1):
        bad = compare_frame(ctx, f, res[f], o)
        assert not bad, "\n".join(bad[:20])
    ctx.close()


def test_fuzzed_parameters_and_scenes(sp):
    """random ParameterDef values inside the GUI's ranges, random staircases and dropout, random show modes, with and
    without levelling: every count and every tap equals the oracle's (tools/fuzz_parity.py; 1 020 more cases were run in round 1, 540 of them on the final kernels)"""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(ROOT, "tools", "fuzz_parity.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    fails = fz.run(16, 2026, verbose=False)
    assert not fails, fails[:2]


def test_mixed_batch_with_empty_and_sparse_frames(sp, orc, synth):
    """one chunk holding ordinary frames next to an empty one (status 1: fewer than 100 points after the crop) and a sparse
    one: the compacted voxel sort and every per-frame offset must keep the frames apart"""
    frames = []
    for fid in (31, 32, 33, 34, 35):
        rec, R = synth.make_frame(fid)
        frames.append([rec.copy(), R])
    frames[1][0]["rgba"] = 0                                              # empty
    keep = np.zeros((424, 512), dtype=bool)
    keep[100:130, 200:260] = True
    frames[3][0]["rgba"] = np.where(keep.reshape(-1), frames[3][0]["rgba"], 0)   # 1 800 pixels at most
    ctx = sp.Context(device=0, width=512, height=424, max_batch=5)
    res = ctx.process_frames(np.stack([f[0] for f in frames]), np.stack([f[1] for f in frames]))
    assert int(res[1]["status"]) == 1 and int(res[1]["n_planes"]) == 0
    for f in range(5):
        o = orc.Oracle(sp.DEFAULT_PARAMS)
        o.process(frames[f][0], 512, 424, frames[f][1])
        bad = compare_frame(ctx, f, res[f], o)
        assert not bad, "frame %d\n" % f + "\n".join(bad[:20])
    ctx.close()


def test_voxel_sort_four_pass_path(sp, orc, synth, monkeypatch):
    """The per-frame voxel sort takes ceil(bits of the voxel grid / 9) stable passes: three for the usual grids, four when the
    grid (or wrapped voxel indices) needs more than 27 bits.  SP_VBITS=32 forces the four-pass path on an ordinary batch, and
    SP_VBITS=10 cannot shrink it below what the grid needs: same bits out.  (An even pass count also leaves the sorted pairs in
    the other buffer, which the voxel_key tap must follow.)"""
    recs, Rs = zip(*[synth.make_frame(fid) for fid in (4, 6, 8)])
    for bits in ("32", "10"):
        monkeypatch.setenv("SP_VBITS", bits)
        ctx = sp.Context(device=0, width=512, height=424, max_batch=3)
        monkeypatch.delenv("SP_VBITS")
        res = ctx.process_frames(np.stack(recs), np.stack(Rs))
        for i in range(3):
            o = orc.Oracle(sp.DEFAULT_PARAMS)
            o.process(recs[i], 512, 424, Rs[i])
            bad = compare_frame(ctx, i, res[i], o)
            assert not bad, "frame %d\n" % i + "\n".join(bad[:20])
        ctx.close()


def test_normals_without_the_exactness_guarantee(sp, orc, synth):
    """k_ingest_normals takes separable box sums only in tiles whose coordinates make every double sum exact (NaN, zero, or
    2^-21 <= |c| < 8); other tiles take the ordered sums.  A scene scaled by four (coordinates up to 18 m) and one sprinkled with
    tiny coordinates must still match the oracle bit for bit."""
    rec, R = synth.make_frame(12)
    big = rec.copy()
    for k in ("x", "y", "z"):
        big[k] = big[k] * np.float32(4.0)
    tiny = rec.copy()
    rng = np.random.default_rng(5)
    idx = rng.choice(tiny.shape[0], 4000, replace=False)
    tiny["x"][idx[:2000]] = np.float32(3e-9)
    tiny["z"][idx[2000:]] = np.float32(-7e-8)
    p = dict(sp.DEFAULT_PARAMS)
    p.update(x_min=-20.0, x_max=20.0, y_min=-20.0, y_max=20.0, z_min=-20.0, z_max=20.0)
    ctx = sp.Context(p, device=0, width=512, height=424, max_batch=2)
    res = ctx.process_frames(np.stack([big, tiny]), np.stack([R, R]))
    for i, r in enumerate((big, tiny)):
        o = orc.Oracle(p)
        o.process(r, 512, 424, R)
        bad = compare_frame(ctx, i, res[i], o)
        assert not bad, "frame %d\n" % i + "\n".join(bad[:20])
    ctx.close()


def test_two_lane_chunk_pipeline_keeps_order_and_bits(sp, synth, monkeypatch):
    """batches of two or more chunks alternate between a context and its twin lane (own workspace, own stream): eleven frames
    through chunks of two must give, record by record, what one big chunk gives -- from host clouds, from device-resident clouds,
    from depth images, and with the twin switched off (SP_NO_TWIN=1)"""
    import torch
    frames = [synth.make_frame(f) for f in (21, 22, 23)]
    recs = np.stack([frames[i % 3][0] for i in range(11)])
    Rs = np.stack([frames[i % 3][1] for i in range(11)])
    big = sp.Context(device=0, width=512, height=424, max_batch=11)
    want = big.process_frames(recs, Rs).copy()
    big.close()
    assert int(want["n_planes"].min()) >= 8
    for no_twin in (False, "3 lanes", True):
        if no_twin == "3 lanes":
            monkeypatch.setenv("SP_LANES", "3")
        elif no_twin:
            monkeypatch.delenv("SP_LANES")
            monkeypatch.setenv("SP_NO_TWIN", "1")
        ctx = sp.Context(device=0, width=512, height=424, max_batch=2)
        got = ctx.process_frames(recs, Rs)
        assert got.tobytes() == want.tobytes()
        d_c = torch.from_numpy(recs.view(np.uint8).reshape(11, -1)).cuda()
        d_R = torch.from_numpy(Rs).cuda()
        torch.cuda.synchronize()
        got = ctx.process_frames(d_c, d_R, device_input=True, nframes=11)
        assert got.tobytes() == want.tobytes()
        # the last chunk (frame 10 alone) ran on lane 0 or 1: its taps must be that frame's
        assert ctx.tap(0, "plane_idx").size == int(want[10]["n_plane_points"])
        ctx.close()
    monkeypatch.delenv("SP_NO_TWIN")
    dfr = [synth.make_depth_frame(f) for f in (21, 22, 23)]
    depth = np.stack([dfr[i % 3][0] for i in range(7)])
    bgrx = np.stack([dfr[i % 3][1] for i in range(7)])
    one = sp.Context(device=0, width=512, height=424, max_batch=7)
    want_d = one.process_depth_frames(depth, synth.INTRINSICS, bgrx=bgrx, mirror=True, R=Rs[:7], nframes=7).copy()
    one.close()
    ctx = sp.Context(device=0, width=512, height=424, max_batch=2)
    got_d = ctx.process_depth_frames(depth, synth.INTRINSICS, bgrx=bgrx, mirror=True, R=Rs[:7], nframes=7)
    assert got_d.tobytes() == want_d.tobytes()
    ctx.close()


def test_product_against_the_oracles_independent_numerics(sp, orc, samples, synth, ctx1):
    """the product and the oracle's default mode share include/sp_detmath.h; eig_math = 2 runs the oracle on its own numerics
    (oracle/indep_math.h: double Jacobi, libm).  Tolerance test next to the bit-exact ones: same plane count and sizes,
    coefficients within 1e-4 (sign of a normal is free), centres within 1e-6, eigenvalues within 1e-3 relative."""
    cases = [(samples["0000"][0], None), (samples["0001"][0], None)] + [synth.make_frame(f) for f in (2, 9)]
    for rec, R in cases:
        res = ctx1.process_frames(rec, None if R is None else R.reshape(1, 9))[0]
        o2 = orc.Oracle(sp.DEFAULT_PARAMS, eig_math=2)
        o2.process(rec, 512, 424, R)
        p2 = o2.tap("planes")
        assert int(res["n_planes"]) == len(p2) and len(p2) >= 6
        for i, q2 in enumerate(p2):
            q = res["planes"][i]
            sg = np.sign(np.dot(q["coef"][:3], q2["coef"][:3]))
            assert int(q["count"]) == int(q2["count"]) and int(q["type"]) == int(q2["type"])
            assert np.abs(q["coef"] - sg * q2["coef"]).max() < 1e-4
            assert np.abs(q["center"] - q2["center"]).max() < 1e-6
            assert np.allclose(q["eval"], q2["eval"], rtol=1e-3, atol=1e-9)


def test_batched_stair_model_equals_per_frame_model(sp, synth, samples):
    """sp_stair_model_batch: the plane clouds of a whole chunk packed by one kernel and copied once, the host model on several
    threads -- record for record what sp_stair_model returns frame by frame (which pulls points per findMinMaxProjPoint call)"""
    frames = [synth.make_frame(f) for f in (0, 3, 7, 12)]
    recs = [samples["0001"][0], samples["0000"][0]] + [f[0] for f in frames]
    Rs = [np.eye(3, dtype=np.float32).reshape(9)] * 2 + [f[1] for f in frames]
    recs.append(recs[2].copy())
    recs[-1]["rgba"] = 0                                      # an empty frame: no planes, no stair
    Rs.append(Rs[2])
    ctx = sp.Context(device=0, width=512, height=424, max_batch=len(recs))
    res = ctx.process_frames(np.stack(recs), np.stack(Rs))
    one = np.stack([ctx.stair_model(i) for i in range(len(recs))])
    for nthreads in (1, 4):
        many = ctx.stair_model_batch(0, len(recs), nthreads=nthreads)
        assert many.tobytes() == one.tobytes()
    part = ctx.stair_model_batch(2, 3)
    assert part.tobytes() == one[2:5].tobytes()
    assert int(one["has_stair"].sum()) >= 4 and int(one[-1]["has_stair"]) == 0 and int(res[-1]["n_planes"]) == 0
    ctx.close()


@pytest.mark.gpu
def test_tapered_tail_of_a_host_batch_keeps_order_and_bits(sp, synth, monkeypatch):
    """a host-input batch of several chunks ends in halved chunks (80, 80, 32, 8 for 200 frames of 80): record for record what
    uniform chunks (SP_TAPER=0) and the device-input path give"""
    import torch
    frames = [synth.make_frame(f) for f in (31, 32, 33, 34)]
    n = 200
    recs = np.stack([frames[i % 4][0] for i in range(n)])
    Rs = np.stack([frames[i % 4][1] for i in range(n)])
    ctx = sp.Context(device=0, width=512, height=424, max_batch=80)
    got = ctx.process_frames(recs, Rs).copy()
    d_c = torch.from_numpy(recs.view(np.uint8).reshape(n, -1)).cuda()
    d_R = torch.from_numpy(Rs).cuda()
    dev = ctx.process_frames(d_c, d_R, device_input=True, nframes=n).copy()
    ctx.close()
    assert int(got["n_planes"].min()) >= 8
    assert got.tobytes() == dev.tobytes()
    for i in range(4, n):                                                # the batch cycles four scenes
        assert got[i].tobytes() == got[i % 4].tobytes(), i
    monkeypatch.setenv("SP_TAPER", "0")
    ctx = sp.Context(device=0, width=512, height=424, max_batch=80)
    uniform = ctx.process_frames(recs, Rs).copy()
    ctx.close()
    assert uniform.tobytes() == got.tobytes()
