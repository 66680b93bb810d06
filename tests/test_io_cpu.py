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
