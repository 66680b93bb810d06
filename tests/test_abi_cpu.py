"""The C-ABI library without a GPU: it loads, exports every symbol include/sp_b200.h declares, its structs have the
layout the Python binding assumes, its host-only entry points work, and every compute entry point FAILS LOUDLY when no
CUDA device exists (there is no CPU path in the product)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, has_cuda

HEADER = os.path.join(ROOT, "include", "sp_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sp_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(sp):
    L = sp.lib()
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "libsp_b200.so does not export " + n


def test_header_is_plain_c_and_layout_matches_binding(sp, tmp_path):
    """compile a C (not C++) translation unit against the header and compare sizeof / offsetof with the numpy dtypes"""
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sp_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(sp_params), sizeof(sp_plane_record), sizeof(sp_frame_result),'
                   ' offsetof(sp_frame_result, planes), offsetof(sp_plane_record, pt_min), offsetof(sp_params, noleg_distance)); return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got[0] == C.sizeof(sp.SpParams)
    assert got[1] == sp.PLANE_DTYPE.itemsize
    assert got[2] == sp.FRAME_DTYPE.itemsize
    assert got[3] == sp.FRAME_DTYPE.fields["planes"][1]
    assert got[4] == sp.PLANE_DTYPE.fields["pt_min"][1]
    assert got[5] == sp.SpParams.noleg_distance.offset


def test_parameter_struct_mirrors_parameterdef(sp):
    """same 35 fields, same order and int/double types as ParameterDef (/root/reference/include/common.h:94-131)"""
    assert len(sp.PARAM_KEYS) == 35
    assert [k for k in sp.PARAM_KEYS if k in sp.INT_KEYS] == ["normal_compute_points", "min_cluster_size", "max_cluster_size",
                                                             "seg_max_iters", "seg_rest_point", "min_num_points"]
    assert sp.PARAM_KEYS[0] == "x_min" and sp.PARAM_KEYS[-1] == "noleg_distance"
    assert sp.SHOW["stair"] == 10 and sp.SHOW["original"] == 0 and len(sp.SHOW_MODES) == 11      # ShowModeDef, common.h:79-92


def test_host_rotation_entry_points_match_oracle(sp, orc):
    rng = np.random.default_rng(3)
    for _ in range(50):
        p, r, y = rng.uniform(-90, 90), rng.uniform(-45, 45), rng.uniform(-180, 180)
        assert np.array_equal(sp.rotation_from_euler(p, r, y), orc.rotation_from_euler(p, r, y))
        q = rng.normal(size=4)
        assert np.array_equal(sp.rotation_from_quat(*q), orc.rotation_from_quat(*q))
        R = sp.rotation_from_quat(*q).reshape(3, 3).astype(np.float64)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-5) and abs(np.linalg.det(R) - 1) < 1e-5
        # a check that shares no restatement with the oracle: the double-precision closed forms of test/reader.cpp:233-271
        a = np.deg2rad([r, p, y])
        Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
        Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
        Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
        m1 = np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]])
        m3 = np.array([[0, 0, -1], [1, 0, 0], [0, -1, 0]])
        assert np.allclose(sp.rotation_from_euler(p, r, y).reshape(3, 3), m3 @ Rx @ Ry @ Rz @ m1, rtol=0, atol=5e-7)
        qn = q / np.linalg.norm(q)
        w, x, yy, z = qn
        M = np.array([[1 - 2 * (yy * yy + z * z), 2 * (x * yy - z * w), 2 * (x * z + yy * w)],
                      [2 * (x * yy + z * w), 1 - 2 * (x * x + z * z), 2 * (yy * z - x * w)],
                      [2 * (x * z - yy * w), 2 * (yy * z + x * w), 1 - 2 * (x * x + yy * yy)]])
        assert np.allclose(R, m3 @ M @ m1, rtol=0, atol=5e-7)


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(sp):
    """without a CUDA device the product refuses to work instead of computing on the CPU"""
    with pytest.raises(sp.SpError) as e:
        sp.Context(device=0, width=512, height=424, max_batch=1)
    assert "CUDA" in str(e.value) or "device" in str(e.value)
    L = sp.lib()
    h = C.c_void_p()
    p = sp.make_params()
    assert L.sp_create(C.byref(p), 0, 512, 424, 1, C.byref(h)) == -2          # SP_ERR_CUDA
    assert L.sp_create(None, 0, 512, 424, 1, C.byref(h)) == -1                # SP_ERR_ARG
    assert L.sp_process_frames(None, None, None, 1, 10, None) == -1
    det = sp.StairDetection(sp.DEFAULT_PARAMS)
    with pytest.raises(sp.SpError):
        det.process(np.zeros(512 * 424, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")]))


def test_product_does_not_import_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/"""
    pkg = os.path.join(ROOT, "stair-perception_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liborc" not in txt and "sp_oracle" not in txt and not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
    for f in os.listdir(os.path.join(ROOT, "include")):
        assert "oracle" not in open(os.path.join(ROOT, "include", f)).read().lower() or f == "sp_detmath.h"


def test_cpp_facade_builds_and_fails_loudly_without_gpu(sp):
    """the C++ host facade (include/StairPerception_b200.hpp) compiles with plain g++, links against the C ABI only, and
    -- without a CUDA device -- reports the error instead of computing anything on the CPU"""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "-s"])
    exe = os.path.join(ROOT, "examples", "files_mode")
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "samples", "0000_cloud.pcd")], capture_output=True, text=True)
    if has_cuda():
        assert r.returncode == 0 and r.stdout.startswith("has_planes 1")
    else:
        assert r.returncode == 3 and "sp_create failed" in r.stderr
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "samples", "0000_cloud.pcd"), "0"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("has_planes 0 planes 0")       # show_mode original: no work at all
