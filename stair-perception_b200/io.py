"""On-disk formats of the reference's files mode: binary PCD v0.7 (x y z rgba, 16 B / point), NNNN_imu.txt and
param.yaml (OpenCV FileStorage YAML 1.0).  Mirrors Reader::run files mode, /root/reference/test/reader.cpp:449-505,
Reader::save_images :109-166 and gui::read_parameters, /root/reference/modules/gui/gui.cpp:534-597."""
import re
import numpy as np

PARAM_KEYS = [
    "x_min", "x_max", "y_min", "y_max", "z_min", "z_max", "normal_compute_points", "voxel_x", "voxel_y", "voxel_z",
    "parallel_angle_diff", "perpendicular_angle_diff", "cluster_tolerance", "min_cluster_size", "max_cluster_size",
    "seg_threshold", "seg_max_iters", "seg_rest_point", "seg_plane_angle_diff", "merge_threshold", "merge_angle_diff",
    "min_num_points", "min_length", "max_length", "min_width", "max_width", "max_g_height", "counter_max_distance",
    "min_height", "max_height", "vertical_angle_diff", "mcv_angle_diff", "center_vector_angle_diff",
    "vertical_plane_angle_diff", "noleg_distance",
]
INT_KEYS = {"normal_compute_points", "min_cluster_size", "max_cluster_size", "seg_max_iters", "seg_rest_point", "min_num_points"}

# /root/reference/param.yaml verbatim (the GUI's slider-quantised fixed points)
DEFAULT_PARAMS = {
    "x_min": -5.0, "x_max": 5.0, "y_min": -5.0, "y_max": 5.0, "z_min": -5.0, "z_max": 5.0,
    "normal_compute_points": 69, "voxel_x": 1.0000000000000000e-02, "voxel_y": 1.0000000000000000e-02,
    "voxel_z": 1.0000000000000000e-02, "parallel_angle_diff": 1.9577777777777779e+01,
    "perpendicular_angle_diff": 2.4466666666666665e+01, "cluster_tolerance": 2.9600000000000001e-02,
    "min_cluster_size": 9, "max_cluster_size": 100000, "seg_threshold": 1.0000000000000000e-02,
    "seg_max_iters": 100, "seg_rest_point": 236, "seg_plane_angle_diff": 7.0,
    "merge_threshold": 2.0000000000000000e-02, "merge_angle_diff": 7.0, "min_num_points": 150,
    "min_length": 1.4999999999999999e-01, "max_length": 1.0, "min_width": 5.0000000000000003e-02,
    "max_width": 3.8000000000000000e-01, "max_g_height": 5.0000000000000012e-02,
    "counter_max_distance": 2.5000000000000000e-01, "min_height": 5.0000000000000003e-02,
    "max_height": 2.0000000000000001e-01, "vertical_angle_diff": 20.0, "mcv_angle_diff": 80.0,
    "center_vector_angle_diff": 70.0, "vertical_plane_angle_diff": 20.0, "noleg_distance": 8.1100000000000005e-01,
}


def read_param_yaml(path):
    """OpenCV FileStorage YAML 1.0: `%YAML:1.0`, `---`, then `key: value` lines (values like `-5.` are valid)."""
    out = dict(DEFAULT_PARAMS)
    with open(path) as f:
        for line in f:
            m = re.match(r"^\s*([A-Za-z_0-9]+)\s*:\s*([-+0-9.eE]+)\s*$", line)
            if m and m.group(1) in out:
                k = m.group(1)
                out[k] = int(float(m.group(2))) if k in INT_KEYS else float(m.group(2))
    return out


def write_param_yaml(path, params):
    with open(path, "w") as f:
        f.write("%YAML:1.0\n---\n")
        for k in PARAM_KEYS:
            v = params[k]
            f.write(f"{k}: {v}\n" if k in INT_KEYS else f"{k}: {v:.16e}\n")


def read_pcd(path):
    """Binary PCD v0.7 with FIELDS x y z rgba / SIZE 4 4 4 4 (what Reader::save_images writes).  Returns
    (records, width, height) with records a (N,) structured array {x,y,z: f4, rgba: u4} of 16 B each."""
    with open(path, "rb") as f:
        data = f.read()
    hdr_end = data.index(b"DATA binary\n") + len(b"DATA binary\n")
    hdr = data[:hdr_end].decode("ascii", "replace")
    kv = {}
    for line in hdr.splitlines():
        parts = line.split()
        if parts and not parts[0].startswith("#"):
            kv[parts[0]] = parts[1:]
    if kv.get("FIELDS") != ["x", "y", "z", "rgba"] or kv.get("SIZE") != ["4", "4", "4", "4"]:
        raise ValueError("unsupported PCD layout: %r" % (kv.get("FIELDS"),))
    w, h = int(kv["WIDTH"][0]), int(kv["HEIGHT"][0])
    n = int(kv["POINTS"][0])
    dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")])
    rec = np.frombuffer(data, dtype=dt, count=n, offset=hdr_end).copy()
    return rec, w, h


def write_pcd(path, rec, width, height):
    n = width * height
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgba\nSIZE 4 4 4 4\nTYPE F F F U\n"
           "COUNT 1 1 1 1\nWIDTH %d\nHEIGHT %d\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (width, height, n))
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(np.ascontiguousarray(rec).tobytes())


def read_imu_txt(path):
    """Third line, first three comma-separated fields = pitch, roll, yaw (degrees), reader.cpp:462-482."""
    with open(path) as f:
        lines = f.read().splitlines()
    vals = [float(v) for v in lines[2].split(",") if v.strip() != ""]
    return {"pitch": vals[0], "roll": vals[1], "yaw": vals[2], "raw": vals}
