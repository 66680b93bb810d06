"""stair-perception segmentation front-end for B200 -- host-side mirror of the reference interface.

The product is `libsp_b200.so` (hand-written sm_100a CUDA behind the C ABI of include/sp_b200.h).  This module is
the thin Python binding of that ABI plus a `StairDetection` facade with the reference class's entry points
(/root/reference/modules/stairperception/StairPerception.hpp:127-184).  There is no CPU path: creating a context
without a CUDA device raises.  The directory is `stair-perception_b200`; import it through
`__graft_entry__.load_package()` (a hyphen is not importable), which registers it as `stair_perception_b200`.
"""
import ctypes as C
import os
import numpy as np

from .io import DEFAULT_PARAMS, PARAM_KEYS, INT_KEYS, read_pcd, write_pcd, read_imu_txt, read_param_yaml, write_param_yaml  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsp_b200.so")

SP_MAX_PLANES = 64
SHOW_MODES = ["original", "range_limited", "voxel", "noleg", "horizontal_cloud", "vertical_cloud", "seg_horizontal_plane",
              "seg_vertical_plane", "merge_horizontal_plane", "merge_vertical_plane", "stair"]
SHOW = {n: i for i, n in enumerate(SHOW_MODES)}


class SpParams(C.Structure):
    _fields_ = [(k, C.c_int if k in INT_KEYS else C.c_double) for k in PARAM_KEYS]


PLANE_DTYPE = np.dtype([
    ("type", "<i4"), ("ptype", "<i4"), ("count", "<i4"), ("first", "<i4"),
    ("coef", "<f4", (4,)), ("center", "<f4", (3,)), ("pcenter", "<f4", (3,)), ("evec", "<f4", (9,)),
    ("eval", "<f4", (3,)), ("pmin", "<f4", (3,)), ("pmax", "<f4", (3,)),
    ("pt_min", "<f4", (3, 3)), ("pt_max", "<f4", (3, 3)), ("idx_min", "<i4", (3,)), ("idx_max", "<i4", (3,)),
])
COUNT_NAMES = ["status", "n_valid", "n_crop", "n_voxel", "n_nonan", "n_noleg", "n_h", "n_v", "n_clusters_h", "n_clusters_v",
               "n_seg_h", "n_seg_v", "n_planes_h", "n_planes_v", "n_planes", "n_plane_points"]
FRAME_DTYPE = np.dtype([(n, "<i4") for n in COUNT_NAMES] + [("planes", PLANE_DTYPE, (SP_MAX_PLANES,))])

TAP_DTYPES = {"xyz": "<f4", "normals": "<f4", "minmax": "<f4", "voxel_key": "<u4", "voxel_code": "u1"}

_lib = None


def lib():
    """Load libsp_b200.so; raises if the CUDA extension has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.sp_create.argtypes = [C.POINTER(SpParams), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.sp_destroy.argtypes = [C.c_void_p]
        L.sp_last_error.restype = C.c_char_p
        L.sp_last_error.argtypes = [C.c_void_p]
        L.sp_set_params.argtypes = [C.c_void_p, C.POINTER(SpParams)]
        L.sp_rotation_from_euler.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.sp_rotation_from_quat.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        for fn in (L.sp_process_frames, L.sp_process_frames_device):
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.sp_tap_bytes.restype = C.c_int64
        L.sp_tap_bytes.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.sp_tap_copy.restype = C.c_int64
        L.sp_tap_copy.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int64]
        L.sp_copy_plane_points.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.sp_stage_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.sp_stage_name.restype = C.c_char_p
        L.sp_stage_name.argtypes = [C.c_int]
        L.sp_launch_count.restype = C.c_int64
        L.sp_launch_count.argtypes = [C.c_void_p]
        L.sp_run_stage_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.sp_process_depth_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.sp_minmax_proj.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sp_set_normal_mode.argtypes = [C.c_void_p, C.c_int]
        L.sp_set_trace.argtypes = [C.c_void_p, C.c_int]
        L.sp_trace_ms.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]
        L.sp_process_frames_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                                C.POINTER(C.c_int64)]
        L.sp_compact_bytes_max.restype = C.c_int64
        L.sp_compact_bytes_max.argtypes = [C.c_int]
        L.sp_expand_compact.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.sp_stream.restype = C.c_void_p
        L.sp_stream.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def make_params(d=None):
    p = SpParams()
    src = dict(DEFAULT_PARAMS)
    if d:
        src.update(d)
    for k in PARAM_KEYS:
        setattr(p, k, src[k])
    return p


def rotation_from_euler(pitch, roll, yaw):
    R = np.zeros(9, dtype=np.float32)
    lib().sp_rotation_from_euler(pitch, roll, yaw, R.ctypes.data)
    return R


def rotation_from_quat(w, x, y, z):
    R = np.zeros(9, dtype=np.float32)
    lib().sp_rotation_from_quat(w, x, y, z, R.ctypes.data)
    return R


def _ptr(a):
    """host pointer of a numpy array / torch CPU tensor, device pointer of a torch CUDA tensor"""
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


STAIR_NODE_DTYPE = np.dtype([("kind", "<i4"), ("count", "<i4"), ("plane_h", "<i4"), ("plane_v", "<i4"), ("line_h", "<f4"), ("line_d", "<f4"),
                             ("line_coeff", "<f4", (6,)), ("height", "<f8"), ("depth", "<f8")])
SP_MAX_VIRTUAL_PLANES = 32
STAIR_DTYPE = np.dtype([("has_stair", "<i4"), ("n_steps", "<i4"), ("n_nodes", "<i4"), ("ptype", "<i4", (SP_MAX_PLANES,)), ("vnormal", "<f4", (3,)),
                        ("snormal", "<f4", (3,)), ("n_virtual", "<i4"), ("nodes", STAIR_NODE_DTYPE, (64,)),
                        ("slide_left", "<f4", (4,)), ("slide_right", "<f4", (4,)),
                        ("contour", "<f4", (SP_MAX_PLANES, 4, 3)), ("contour_width", "<i4", (SP_MAX_PLANES,)),
                        ("ground_contour", "<f4", (4, 3)), ("ground_contour_width", "<i4"),
                        ("virtual_contour_width", "<i4", (SP_MAX_VIRTUAL_PLANES,)), ("virtual_contour", "<f4", (SP_MAX_VIRTUAL_PLANES, 4, 3)),
                        ("pad", "<i4")])


class ResultDef(C.Structure):
    """ResultDef, /root/reference/include/common.h:133-141"""
    _fields_ = [("has_stair", C.c_int32), ("height", C.c_float), ("width", C.c_float), ("v_height", C.c_float), ("v_depth", C.c_float),
                ("time", C.c_double)]


def resolve_result(result, has_stair, step_height_m, step_depth_m, line_h_m, line_d_m):
    """DProcessor::resolve_result (test/processor.cpp:46-82) on a ResultDef held across frames"""
    L = lib()
    L.sp_resolve_result.argtypes = [C.POINTER(ResultDef), C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
    L.sp_resolve_result.restype = None
    L.sp_resolve_result(C.byref(result), 1 if has_stair else 0, step_height_m, step_depth_m, line_h_m, line_d_m)
    return result


def pack_robot(result):
    """the 19-byte packet of DProcessor::pack_and_send_robot (test/processor.cpp:85-100)"""
    L = lib()
    L.sp_pack_robot.argtypes = [C.POINTER(ResultDef), C.c_void_p]
    buf = (C.c_uint8 * 19)()
    assert L.sp_pack_robot(C.byref(result), buf) == 19
    return bytes(buf)


class SpError(RuntimeError):
    pass


class Context:
    """One sp_ctx: device workspaces for `max_batch` frames of width x height points on one GPU."""

    def __init__(self, params=None, device=0, width=512, height=424, max_batch=1):
        self.L = lib()
        self.width, self.height, self.N, self.max_batch = width, height, width * height, max_batch
        self.params = dict(DEFAULT_PARAMS)
        if params:
            self.params.update(params)
        self._p = make_params(self.params)
        h = C.c_void_p()
        rc = self.L.sp_create(C.byref(self._p), device, width, height, max_batch, C.byref(h))
        if rc != 0:
            raise SpError("sp_create failed (%d): %s" % (rc, self.L.sp_last_error(None).decode()))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.sp_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _check(self, rc, what):
        if rc < 0:
            raise SpError("%s failed (%d): %s" % (what, rc, self.L.sp_last_error(self.h).decode()))
        return rc

    def set_params(self, params):
        self.params.update(params)
        self._p = make_params(self.params)
        self._check(self.L.sp_set_params(self.h, C.byref(self._p)), "sp_set_params")

    def process_frames(self, clouds, R=None, show_mode=10, nframes=None, device_input=False, results=None):
        """clouds: N x 16 B records per frame (numpy structured / (.,4) float32 view / torch tensor).  R: (nframes, 9)
        float32 or None.  Returns a FRAME_DTYPE array of nframes records."""
        nbytes = clouds.nbytes if isinstance(clouds, np.ndarray) else clouds.numel() * clouds.element_size()
        if nframes is None:
            nframes = nbytes // (self.N * 16)
        if nbytes < nframes * self.N * 16:
            raise ValueError("cloud buffer too small")
        if results is None:
            results = np.zeros(nframes, dtype=FRAME_DTYPE)
        Rp = None
        if R is not None:
            Rp = _ptr(R)
        fn = self.L.sp_process_frames_device if device_input else self.L.sp_process_frames
        if device_input and hasattr(clouds, "is_cuda") and clouds.is_cuda:
            import torch                                     # the library's streams do not know torch's: finish the producer first
            torch.cuda.current_stream(clouds.device).synchronize()
        self._check(fn(self.h, _ptr(clouds), Rp, nframes, int(show_mode), results.ctypes.data), "sp_process_frames")
        return results

    def process_frames_compact(self, clouds, R=None, show_mode=10, nframes=None, device_input=False, out=None, offsets=None):
        """Same processing, compact records left ON THE DEVICE: `out` is a torch uint8 CUDA tensor (made here when None; any
        capacity >= compact_bytes_max(nframes) always suffices), `offsets` an optional int64 CUDA tensor of nframes + 1 entries.
        Returns (out, nbytes): out[:nbytes] holds {16 int32 counters, n_planes plane records} per frame, back to back."""
        import torch
        nbytes = clouds.nbytes if isinstance(clouds, np.ndarray) else clouds.numel() * clouds.element_size()
        if nframes is None:
            nframes = nbytes // (self.N * 16)
        if nbytes < nframes * self.N * 16:
            raise ValueError("cloud buffer too small")
        if out is None:
            out = torch.empty(int(self.L.sp_compact_bytes_max(nframes)), dtype=torch.uint8, device="cuda")
        if device_input and hasattr(clouds, "is_cuda") and clouds.is_cuda:
            torch.cuda.current_stream(clouds.device).synchronize()
        total = C.c_int64(0)
        self._check(self.L.sp_process_frames_compact(self.h, _ptr(clouds), None if R is None else _ptr(R), 1 if device_input else 0, nframes,
                                                     int(show_mode), out.data_ptr(), out.numel(), None if offsets is None else offsets.data_ptr(),
                                                     C.byref(total)), "sp_process_frames_compact")
        return out, int(total.value)

    def process_depth_frames(self, depth, intrinsics, bgrx=None, mirror=True, R=None, show_mode=10, nframes=None, results=None):
        """depth: float32 or uint16 millimetres, nframes x H x W (numpy or pinned torch CPU tensor); intrinsics = (fx, fy, cx, cy);
        bgrx: optional registered colour, nframes x H x W x 4 uint8.  Returns FRAME_DTYPE records."""
        itemsize = depth.dtype.itemsize if isinstance(depth, np.ndarray) else depth.element_size()
        if itemsize not in (2, 4):
            raise ValueError("depth must be float32 or uint16")
        nbytes = depth.nbytes if isinstance(depth, np.ndarray) else depth.numel() * itemsize
        if nframes is None:
            nframes = nbytes // (self.N * itemsize)
        if results is None:
            results = np.zeros(nframes, dtype=FRAME_DTYPE)
        K = (C.c_float * 4)(*[float(v) for v in intrinsics])
        self._check(self.L.sp_process_depth_frames(self.h, _ptr(depth), 0 if itemsize == 4 else 1, None if bgrx is None else _ptr(bgrx), K,
                                                   1 if mirror else 0, None if R is None else _ptr(R), nframes, int(show_mode), results.ctypes.data),
                    "sp_process_depth_frames")
        return results

    def run_stage_device(self, clouds_ptr, R_ptr, nframes, stage):
        self._check(self.L.sp_run_stage_device(self.h, clouds_ptr, R_ptr, nframes, stage), "sp_run_stage_device")

    def stream(self):
        return self.L.sp_stream(self.h)

    def tap(self, frame, name):
        n = self._check(self.L.sp_tap_bytes(self.h, frame, name.encode()), "sp_tap_bytes(%s)" % name)
        dt = np.dtype(TAP_DTYPES.get(name, "<f4" if name.endswith("_coef") else "<i4"))
        out = np.empty(int(n) // dt.itemsize, dtype=dt)
        if n:
            self._check(self.L.sp_tap_copy(self.h, frame, name.encode(), out.ctypes.data, n), "sp_tap_copy")
        if name in ("xyz", "normals"):
            out = out.reshape(-1, 3)
        elif name.endswith("_coef"):
            out = out.reshape(-1, 4)
        elif name.endswith("_log"):
            out = out.reshape(-1, 10)
        return out

    def plane_points(self, frame, plane, count):
        xyz = np.empty((count, 3), dtype=np.float32)
        idx = np.empty(count, dtype=np.int32)
        n = self._check(self.L.sp_copy_plane_points(self.h, frame, plane, xyz.ctypes.data, idx.ctypes.data, count), "sp_copy_plane_points")
        return xyz[:n], idx[:n]

    def minmax_proj(self, frame, plane, direction, center=None):
        """findMinMaxProjPoint on the device: returns (max_xyz, min_xyz, max_pixel, min_pixel)"""
        dv = np.ascontiguousarray(direction, dtype=np.float32)
        cv = None if center is None else np.ascontiguousarray(center, dtype=np.float32)
        mx, mn = np.zeros(3, np.float32), np.zeros(3, np.float32)
        pm, pn = C.c_int32(), C.c_int32()
        self._check(self.L.sp_minmax_proj(self.h, frame, plane, None if cv is None else cv.ctypes.data, dv.ctypes.data, mx.ctypes.data,
                                          mn.ctypes.data, C.byref(pm), C.byref(pn)), "sp_minmax_proj")
        return mx, mn, int(pm.value), int(pn.value)

    def stair_model(self, frame=0):
        """StairDetection::stairModel on frame `frame` of the last chunk: returns a STAIR_DTYPE record; steps() lists
        (height, depth, line.h, line.d) per step -- the columns of printStairEstParam"""
        out = np.zeros(1, dtype=STAIR_DTYPE)
        L = self.L
        L.sp_stair_model.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self._check(L.sp_stair_model(self.h, frame, out.ctypes.data), "sp_stair_model")
        return out[0]

    def stair_model_batch(self, first_frame=0, nframes=None, nthreads=0):
        """sp_stair_model for frames [first_frame, first_frame + nframes) of the last chunk: one pack kernel, three copies, the
        host model on `nthreads` threads.  Returns a STAIR_DTYPE array."""
        if nframes is None:
            nframes = self.max_batch - first_frame
        out = np.zeros(nframes, dtype=STAIR_DTYPE)
        L = self.L
        L.sp_stair_model_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        self._check(L.sp_stair_model_batch(self.h, first_frame, nframes, out.ctypes.data, nthreads), "sp_stair_model_batch")
        return out

    def stage_ms(self):
        arr = (C.c_float * 16)()
        n = self.L.sp_stage_ms(self.h, arr, 16)
        return {self.L.sp_stage_name(i).decode(): float(arr[i]) for i in range(n)}

    def set_normal_mode(self, mode):
        """0 = organised-image normals (live path), 1 = kNN covariance normals for unorganised clouds (a3')"""
        self._check(self.L.sp_set_normal_mode(self.h, int(mode)), "sp_set_normal_mode")

    def set_trace(self, on=True):
        self.L.sp_set_trace(self.h, 1 if on else 0)

    def trace_ms(self):
        """[(kernel name, ms)] of the last chunk, in launch order (needs set_trace / SP_TRACE=1)"""
        out = []
        for i in range(max(0, self.L.sp_trace_ms(self.h, -1, None, None))):
            nm, ms = C.c_char_p(), C.c_float()
            if self.L.sp_trace_ms(self.h, i, C.byref(nm), C.byref(ms)) == 0:
                out.append((nm.value.decode(), float(ms.value)))
        return out

    def launch_count(self):
        return int(self.L.sp_launch_count(self.h))


_shared = {}


def _shared_context(params, device, width, height):
    """The reference builds a StairDetection per frame (test/processor.cpp:156); device state is process-level."""
    key = (device, width, height)
    ctx = _shared.get(key)
    if ctx is None:
        ctx = _shared[key] = Context(params, device=device, width=width, height=height, max_batch=1)
    else:
        ctx.set_params(params)
    return ctx


class StairDetection:
    """Facade with the reference class's entry points (StairPerception.hpp:127-184): construct with the 35
    ParameterDef fields, call process(cloud_in, show_mode).  Returns (has_planes, planes, cloud_show_idx) where
    `planes` are the sorted hand-off records (vector_plane_sorted) with their point lists."""

    def __init__(self, parameters=None, device=0):
        self.parameters = dict(DEFAULT_PARAMS)
        if parameters:
            self.parameters.update(parameters)
        self.device = device

    def process(self, cloud_in, show_mode=SHOW["stair"], width=512, height=424, rotation=None):
        if isinstance(show_mode, str):
            show_mode = SHOW[show_mode]
        ctx = _shared_context(self.parameters, self.device, width, height)
        cloud = np.ascontiguousarray(cloud_in)
        res = ctx.process_frames(cloud, None if rotation is None else np.ascontiguousarray(rotation, dtype=np.float32), show_mode, nframes=1)[0]
        planes = []
        for i in range(int(res["n_planes"])):
            rec = res["planes"][i]
            xyz, idx = ctx.plane_points(0, i, int(rec["count"]))
            planes.append({"record": rec, "xyz": xyz, "pixel_idx": idx})
        self.last_result = res
        self.last_context = ctx
        self.last_stair = None
        if show_mode == SHOW["stair"] and int(res["n_planes"]) > 0:
            # the reference's process() answers "stair found" (StairPerception.cpp:190)
            self.last_stair = ctx.stair_model(0)
            return bool(self.last_stair["has_stair"]), planes, res
        return False, planes, res


def expand_compact(buf, nframes):
    """host-side: compact records (numpy uint8 / bytes) -> FRAME_DTYPE array of nframes fixed-size records"""
    b = np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.view(np.uint8).reshape(-1))
    out = np.zeros(nframes, dtype=FRAME_DTYPE)
    rc = lib().sp_expand_compact(b.ctypes.data, b.nbytes, nframes, out.ctypes.data)
    if rc != 0:
        raise SpError("sp_expand_compact: malformed compact buffer (%d)" % rc)
    return out


def compact_records(results):
    """host-side inverse of expand_compact (numpy): FRAME_DTYPE records -> compact bytes"""
    parts = []
    for r in results:
        raw = np.frombuffer(r.tobytes(), dtype=np.uint8)
        parts.append(raw[:64 + int(r["n_planes"]) * PLANE_DTYPE.itemsize])
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


def stair_steps(stair):
    """[(height, depth, V_Distance, H_Distance)] of a STAIR_DTYPE record, as StairDetection::printStairEstParam prints them"""
    return [(float(n["height"]), float(n["depth"]), float(n["line_h"]), float(n["line_d"])) for n in stair["nodes"][:int(stair["n_nodes"])] if int(n["kind"]) == 1]
