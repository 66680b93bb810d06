"""On-disk formats of the reference's files mode (test/reader.cpp:109-166, :449-505) and param.yaml
(modules/gui/gui.cpp:534-653)."""
import os

import numpy as np

from conftest import GOLDEN


def test_pcd_roundtrip_matches_recorded_file(sp, tmp_path, samples):
    rec, w, h = samples["0000"]
    out = tmp_path / "0007_cloud.pcd"
    sp.write_pcd(str(out), rec, w, h)
    rec2, w2, h2 = sp.read_pcd(str(out))
    assert (w2, h2) == (w, h) and np.array_equal(rec.view(np.uint8), rec2.view(np.uint8))
    # same 11-line header and payload as the recorded sample (which only adds page padding)
    a = open(os.path.join(GOLDEN, "samples", "0000_cloud.pcd"), "rb").read()
    b = out.read_bytes()
    assert a[:len(b)] == b and set(a[len(b):]) <= {0}


def test_imu_txt(sp):
    imu = sp.read_imu_txt(os.path.join(GOLDEN, "samples", "0001_imu.txt"))
    assert (imu["pitch"], imu["roll"], imu["yaw"]) == (26.96, 0.12, 48.8)
    imu = sp.read_imu_txt(os.path.join(GOLDEN, "samples", "0000_imu.txt"))
    assert (imu["pitch"], imu["roll"], imu["yaw"]) == (76.39, 16.88, 25.2)


def test_param_yaml_is_the_reference_file(sp, tmp_path):
    p = sp.read_param_yaml(os.path.join(GOLDEN, "param.yaml"))         # verbatim copy of /root/reference/param.yaml
    assert p == sp.DEFAULT_PARAMS
    out = tmp_path / "param.yaml"
    sp.write_param_yaml(str(out), p)
    assert sp.read_param_yaml(str(out)) == p
    assert out.read_text().startswith("%YAML:1.0\n---\n")


def test_synthetic_frame_shape(synth):
    rec, R = synth.make_frame(0)
    assert rec.shape == (512 * 424,) and rec.dtype.itemsize == 16 and R.shape == (9,)
    valid = (rec["rgba"] & 0xFFFFFF) != 0
    assert 0.5 < valid.mean() < 0.85 and np.isfinite(rec["x"]).all()
    rec2, R2 = synth.make_frame(0)
    assert np.array_equal(rec.view(np.uint8), rec2.view(np.uint8)) and np.array_equal(R, R2)       # seeded
    Rm = R.reshape(3, 3).astype(np.float64)
    assert np.allclose(Rm @ Rm.T, np.eye(3), atol=1e-6)


def test_result_filter_and_robot_packet(sp):
    """next rows f3/f4: DProcessor::resolve_result (test/processor.cpp:46-82) and the packed 19-byte union of
    test/processor.h:75-95 {0xFF, flag, v_depth, v_height, depth, height, 0xFE}; known answer = the status bar of
    pic/screenshot1.png for samples/0001 (height 110.046, width 280.456, v_depth 526.096; v_height keeps its stale value
    because 490.394 mm fails the 700..1150 gate)."""
    import struct
    r = sp.ResultDef()
    r.v_height = 883.672                                            # the stale value visible in the screenshot
    sp.resolve_result(r, True, 0.110046, 0.280456, 0.490394, 0.526096)
    assert r.has_stair == 1
    assert abs(r.height - 110.046) < 1e-3 and abs(r.width - 280.456) < 1e-3 and abs(r.v_depth - 526.096) < 1e-3
    assert abs(r.v_height - 883.672) < 1e-3
    pkt = sp.pack_robot(r)
    assert len(pkt) == 19 and pkt[0] == 0xFF and pkt[1] == 1 and pkt[18] == 0xFE
    v_depth, v_height, depth, height = struct.unpack("<ffff", pkt[2:18])
    assert (v_depth, v_height, depth, height) == (r.v_depth, r.v_height, r.width, r.height)
    sp.resolve_result(r, True, 0.30, 0.10, 0.9, 0.7)                # all out of range except v_height
    assert abs(r.height - 110.046) < 1e-3 and abs(r.width - 280.456) < 1e-3 and abs(r.v_height - 900.0) < 1e-3 and abs(r.v_depth - 526.096) < 1e-3
    sp.resolve_result(r, False, 0.12, 0.3, 0.8, 0.5)                # no stair: nothing but the flag changes
    assert r.has_stair == 0 and abs(r.height - 110.046) < 1e-3
    assert sp.pack_robot(r)[1] == 0
