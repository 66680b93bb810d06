"""ctypes wrapper of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

PARAM_FIELDS = [
    ("x_min", C.c_double), ("x_max", C.c_double), ("y_min", C.c_double), ("y_max", C.c_double),
    ("z_min", C.c_double), ("z_max", C.c_double), ("normal_compute_points", C.c_int),
    ("voxel_x", C.c_double), ("voxel_y", C.c_double), ("voxel_z", C.c_double),
    ("parallel_angle_diff", C.c_double), ("perpendicular_angle_diff", C.c_double),
    ("cluster_tolerance", C.c_double), ("min_cluster_size", C.c_int), ("max_cluster_size", C.c_int),
    ("seg_threshold", C.c_double), ("seg_max_iters", C.c_int), ("seg_rest_point", C.c_int),
    ("seg_plane_angle_diff", C.c_double), ("merge_threshold", C.c_double), ("merge_angle_diff", C.c_double),
    ("min_num_points", C.c_int), ("min_length", C.c_double), ("max_length", C.c_double),
    ("min_width", C.c_double), ("max_width", C.c_double), ("max_g_height", C.c_double),
    ("counter_max_distance", C.c_double), ("min_height", C.c_double), ("max_height", C.c_double),
    ("vertical_angle_diff", C.c_double), ("mcv_angle_diff", C.c_double),
    ("center_vector_angle_diff", C.c_double), ("vertical_plane_angle_diff", C.c_double),
    ("noleg_distance", C.c_double),
]


class OrcParams(C.Structure):
    _fields_ = PARAM_FIELDS


PLANE_DTYPE = np.dtype([
    ("type", "<i4"), ("ptype", "<i4"), ("count", "<i4"), ("first", "<i4"),
    ("coef", "<f4", (4,)), ("center", "<f4", (3,)), ("pcenter", "<f4", (3,)), ("evec", "<f4", (9,)),
    ("eval", "<f4", (3,)), ("pmin", "<f4", (3,)), ("pmax", "<f4", (3,)),
    ("pt_min", "<f4", (3, 3)), ("pt_max", "<f4", (3, 3)), ("idx_min", "<i4", (3,)), ("idx_max", "<i4", (3,)),
])

COUNT_NAMES = ["status", "n_valid", "n_crop", "n_voxel", "n_nonan", "n_noleg", "n_h", "n_v", "n_clusters_h",
               "n_clusters_v", "n_seg_h", "n_seg_v", "n_planes_h", "n_planes_v", "n_planes", "n_plane_points"]

TAP_DTYPES = {
    "xyz": "<f4", "normals": "<f4", "minmax": "<f4", "voxel_key": "<u4", "h_seg_coef": "<f4", "v_seg_coef": "<f4",
    "h_merge_coef": "<f4", "v_merge_coef": "<f4", "planes": PLANE_DTYPE,
}

_lib = None


def _host_signature():
    """the oracle is built -march=native: a library built on another CPU model (it travels with the repo snapshot to the GPU
    box) must be rebuilt there.  Signature = the CPU feature flags of this host."""
    import hashlib
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    return hashlib.sha256(flags.encode()).hexdigest()[:16]


def build_flags():
    """the CXXFLAGS line of oracle/Makefile (recorded next to every CPU-baseline number)"""
    with open(os.path.join(_HERE, "Makefile")) as f:
        for line in f:
            if line.startswith("CXXFLAGS"):
                return "g++ " + line.split("=", 1)[1].strip()
    return "unknown"


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    sig_file = so + ".host"
    srcs = [os.path.join(_HERE, f) for f in ("sp_oracle.cpp", "sp_oracle.h", "indep_math.h", "Makefile")] + [os.path.join(_HERE, "..", "include", "sp_detmath.h")]
    sig = _host_signature()
    try:
        with open(sig_file) as f:
            same_host = f.read().strip() == sig
    except OSError:
        same_host = False
    if force or not same_host or not os.path.exists(so) or any(os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
        with open(sig_file, "w") as f:
            f.write(sig + "\n")
    return so


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcParams)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_mode.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.orc_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_process_batch.restype = C.c_int64
        L.orc_process_batch.argtypes = [C.POINTER(OrcParams), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_tap_bytes.restype = C.c_int64
        L.orc_tap_bytes.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_tap_copy.restype = C.c_int64
        L.orc_tap_copy.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
        L.orc_tap_names.restype = C.c_int64
        L.orc_tap_names.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        L.orc_stage_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        L.orc_audit.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        L.orc_stage_name.restype = C.c_char_p
        L.orc_stage_name.argtypes = [C.c_int]
        L.orc_rotation_from_euler.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_rotation_from_quat.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_mt19937_draws.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
        L.orc_glibc_rand.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
        L.orc_depth_to_cloud.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p]
        L.orc_minmax_proj.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_eigen33.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_jacobi3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def make_params(d):
    p = OrcParams()
    for name, _ in PARAM_FIELDS:
        setattr(p, name, d[name])
    return p


class Oracle:
    def __init__(self, params: dict, **modes):
        self.L = lib()
        self.p = make_params(params)
        self.h = self.L.orc_create(C.byref(self.p))
        for k, v in modes.items():
            assert self.L.orc_set_mode(self.h, k.encode(), int(v)) == 0, k

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def process(self, cloud, width, height, R=None, show_mode=10, keep_taps=True, normal_mode=0):
        """cloud: (N,4) float32/uint32 view or raw bytes of N x 16 B records.  Returns number of planes."""
        buf = np.ascontiguousarray(cloud)
        assert buf.nbytes == width * height * 16
        Rp = None
        if R is not None:
            R = np.ascontiguousarray(R, dtype=np.float32)
            Rp = R.ctypes.data
        return self.L.orc_process(self.h, buf.ctypes.data, width, height, Rp, show_mode, normal_mode, 1 if keep_taps else 0)

    def tap_names(self):
        n = self.L.orc_tap_names(self.h, None, 0)
        b = C.create_string_buffer(int(n))
        self.L.orc_tap_names(self.h, b, n)
        return [s for s in b.value.decode().split("\n") if s]

    def tap(self, name):
        n = self.L.orc_tap_bytes(self.h, name.encode())
        if n < 0:
            return None
        dt = np.dtype(TAP_DTYPES.get(name, "<i4"))
        out = np.empty(int(n) // dt.itemsize, dtype=dt)
        if n:
            self.L.orc_tap_copy(self.h, name.encode(), out.ctypes.data, n)
        if name in ("xyz", "normals"):
            out = out.reshape(-1, 3)
        if name.endswith("_coef"):
            out = out.reshape(-1, 4)
        if name.endswith("_log"):
            out = out.reshape(-1, 10)
        return out

    def minmax_proj(self, plane, direction, center=None):
        dv = np.ascontiguousarray(direction, dtype=np.float32)
        cv = None if center is None else np.ascontiguousarray(center, dtype=np.float32)
        mx, mn = np.zeros(3, np.float32), np.zeros(3, np.float32)
        pm, pn = C.c_int32(), C.c_int32()
        rc = self.L.orc_minmax_proj(self.h, plane, None if cv is None else cv.ctypes.data, dv.ctypes.data, mx.ctypes.data, mn.ctypes.data,
                                    C.byref(pm), C.byref(pn))
        assert rc == 0
        return mx, mn, int(pm.value), int(pn.value)

    def counts(self):
        return dict(zip(COUNT_NAMES, self.tap("counts").tolist()))

    AUDIT_NAMES = ["n_eigen33", "n_jacobi", "n_acos", "n_log", "eigen33_max_angle", "eigen33_max_angle_wellsep", "eigen33_max_val",
                   "jacobi_max_val", "jacobi_max_angle", "acos_max_abs", "log_max_rel"]

    def audit(self):
        """deviation of the shared-header numerics (include/sp_detmath.h) from the oracle's own (oracle/indep_math.h) over the
        last frame processed with keep_taps"""
        arr = (C.c_double * 16)()
        n = self.L.orc_audit(self.h, arr, 16)
        return {self.AUDIT_NAMES[i]: arr[i] for i in range(n)}

    def stage_ms(self):
        arr = (C.c_double * 16)()
        n = self.L.orc_stage_ms(self.h, arr, 16)
        return {self.L.orc_stage_name(i).decode(): arr[i] for i in range(n)}


def process_batch(params: dict, clouds, width, height, R=None, nthreads=1, normal_arith=1):
    L = lib()
    p = make_params(params)
    clouds = np.ascontiguousarray(clouds)
    nframes = clouds.nbytes // (width * height * 16)
    counts = np.zeros(nframes, dtype=np.int32)
    Rp = None
    if R is not None:
        R = np.ascontiguousarray(R, dtype=np.float32)
        Rp = R.ctypes.data
    tot = L.orc_process_batch(C.byref(p), clouds.ctypes.data, width, height, Rp, nframes, nthreads, normal_arith, counts.ctypes.data)
    return int(tot), counts


def rotation_from_euler(pitch, roll, yaw):
    R = np.zeros(9, dtype=np.float32)
    lib().orc_rotation_from_euler(pitch, roll, yaw, R.ctypes.data)
    return R


def rotation_from_quat(w, x, y, z):
    R = np.zeros(9, dtype=np.float32)
    lib().orc_rotation_from_quat(w, x, y, z, R.ctypes.data)
    return R


def mt19937_draws(seed, n):
    out = np.zeros(n, dtype=np.int32)
    lib().orc_mt19937_draws(seed, n, out.ctypes.data)
    return out


def glibc_rand(seed, n):
    out = np.zeros(n, dtype=np.int32)
    lib().orc_glibc_rand(seed, n, out.ctypes.data)
    return out


def depth_to_cloud(depth, intrinsics, bgrx=None, mirror=True):
    """K2G::updateCloud on the CPU: depth (H, W) float32 or uint16 millimetres -> records[H*W] {x,y,z,rgba}"""
    depth = np.ascontiguousarray(depth)
    H, W = depth.shape
    rec = np.zeros(W * H, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")])
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    cp = None
    if bgrx is not None:
        bgrx = np.ascontiguousarray(bgrx, dtype=np.uint8)
        cp = bgrx.ctypes.data
    lib().orc_depth_to_cloud(depth.ctypes.data, 0 if depth.dtype == np.float32 else 1, cp, W, H, fx, fy, cx, cy, 1 if mirror else 0, rec.ctypes.data)
    return rec
