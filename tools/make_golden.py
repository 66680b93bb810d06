"""Regenerates tests/golden/oracle_digests.json: counts, SHA-256 of every stage tap and the plane records of the CPU
oracle on the two recorded sample frames and on seeded synthetic frames.  The reference itself cannot be built in
this image (PCL/Eigen/Boost/Qt absent), so these goldens pin the ORACLE (and, through the -m gpu tests, the CUDA path)
against regressions; they are not outputs of the reference binary.   usage: python tools/make_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

sp = load_package()
import importlib  # noqa: E402

synth = importlib.import_module("stair_perception_b200.synth")
from oracle import orc  # noqa: E402

TAPS = ["xyz", "normals", "minmax", "voxel_key", "voxel_idx", "nonan_idx", "noleg_idx", "h_idx", "v_idx", "h_root", "v_root",
        "h_log", "v_log", "h_seg_coef", "h_seg_count", "h_seg_idx", "v_seg_coef", "v_seg_count", "v_seg_idx",
        "h_merge_coef", "h_merge_count", "h_merge_idx", "v_merge_coef", "v_merge_count", "v_merge_idx", "plane_idx"]


def canon(a):
    """canonical bytes: NaN payloads and the sign of zero do not count"""
    a = np.ascontiguousarray(a)
    if a.dtype.kind == "f":
        a = a.copy()
        a[np.isnan(a)] = np.nan
        a[a == 0] = 0.0
    return a.tobytes()


def digest_frame(o, rec, R=None):
    out = {"counts": o.counts(), "taps": {}, "input_sha256": hashlib.sha256(rec.tobytes() + (b"" if R is None else np.asarray(R, dtype=np.float32).tobytes())).hexdigest()}
    for t in TAPS:
        a = o.tap(t)
        out["taps"][t] = {"n": int(a.size), "sha256": hashlib.sha256(canon(a)).hexdigest()}
    pl = o.tap("planes")
    out["planes"] = [{"type": int(p["type"]), "count": int(p["count"]), "coef": [float(v).hex() for v in p["coef"]],
                      "center": [float(v).hex() for v in p["center"]], "extent": [float(p["pmax"][k] - p["pmin"][k]).hex() for k in range(3)]}
                     for p in pl]
    return out


def main():
    gold = {"params": "param.yaml (reference defaults)", "frames": {}}
    for name in ("0000", "0001"):
        rec, w, h = sp.read_pcd(os.path.join(ROOT, "tests", "golden", "samples", name + "_cloud.pcd"))
        o = orc.Oracle(sp.DEFAULT_PARAMS)
        o.process(rec, w, h)
        gold["frames"]["sample_" + name] = digest_frame(o, rec)
    for fid in (3, 11):
        rec, R = synth.make_frame(fid)
        o = orc.Oracle(sp.DEFAULT_PARAMS)
        o.process(rec, 512, 424, R)
        gold["frames"]["synth_%d" % fid] = digest_frame(o, rec, R)
    with open(os.path.join(ROOT, "tests", "golden", "oracle_digests.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", len(gold["frames"]), "frames")


if __name__ == "__main__":
    main()
