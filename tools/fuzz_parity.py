"""Parity fuzzing on a GPU: random parameter sets (inside the GUI's ranges), random synthetic scenes, random show modes,
GPU vs oracle on every tap.  usage: python tools/fuzz_parity.py [cases=40] [seed=0]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402

sp = load_package()
synth = importlib.import_module("stair_perception_b200.synth")
from oracle import orc  # noqa: E402
from parity import compare_frame  # noqa: E402



def run(cases, seed, verbose=True):
    """returns the list of (case, mismatch descriptions) that differ"""
    rng = np.random.default_rng(seed)
    failures = []
    for case in range(cases):
        _one(case, rng, failures, verbose)
    return failures


def _one(case, rng, failures, verbose):
    p = dict(sp.DEFAULT_PARAMS)
    p.update(
        x_min=float(rng.uniform(-5, 0.2)), x_max=float(rng.uniform(0.9, 5)), y_min=float(rng.uniform(-5, 0.5)), y_max=float(rng.uniform(1.5, 5)),
        z_min=float(rng.uniform(-5, -0.3)), z_max=float(rng.uniform(0.3, 5)),
        voxel_x=float(rng.uniform(0.01, 0.03)), voxel_y=float(rng.uniform(0.01, 0.03)), voxel_z=float(rng.uniform(0.01, 0.03)),
        parallel_angle_diff=float(rng.uniform(5, 40)), perpendicular_angle_diff=float(rng.uniform(5, 40)),
        cluster_tolerance=float(rng.uniform(0.02, 0.08)), min_cluster_size=int(rng.integers(3, 40)), max_cluster_size=int(rng.integers(3000, 100000)),
        seg_threshold=float(rng.uniform(0.004, 0.03)), seg_max_iters=int(rng.integers(5, 200)), seg_rest_point=int(rng.integers(30, 600)),
        seg_plane_angle_diff=float(rng.uniform(2, 20)), merge_threshold=float(rng.uniform(0.005, 0.05)), merge_angle_diff=float(rng.uniform(2, 20)),
        noleg_distance=float(rng.uniform(0.1, 1.0)))
    fid = int(rng.integers(1000, 100000))
    rec, R = synth.make_frame(fid, rise=float(rng.uniform(0.08, 0.2)), run=float(rng.uniform(0.2, 0.4)), width=float(rng.uniform(0.6, 1.6)),
                              n_steps=int(rng.integers(2, 7)), dropout=float(rng.uniform(0.05, 0.45)))
    mode = int(rng.choice([10, 10, 10, 10, 9, 8, 6, 4, 3, 2, 1]))
    use_R = rng.random() < 0.8
    if not use_R:                                   # already levelled cloud: apply R on the host (any float result is fine as input)
        xyz = np.stack([rec["x"], rec["y"], rec["z"]], 1).astype(np.float64) @ R.reshape(3, 3).astype(np.float64).T
        rec = rec.copy()
        rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    ctx = sp.Context(p, device=0, width=512, height=424, max_batch=1)
    res = ctx.process_frames(rec, R.reshape(1, 9) if use_R else None, show_mode=mode)[0]
    o = orc.Oracle(p)
    o.process(rec, 512, 424, R if use_R else None, show_mode=mode)
    bad = compare_frame(ctx, 0, res, o)
    flag = int(res["status"]) >> 8
    if verbose:
        print("case %3d frame %6d mode %2d R %d planes %2d status %#x %s" % (case, fid, mode, use_R, int(res["n_planes"]), int(res["status"]), "OK" if not bad else "MISMATCH"))
    if bad:
        failures.append((case, bad))
        if verbose:
            print("   params:", {k: p[k] for k in p if p[k] != sp.DEFAULT_PARAMS[k]})
            for b in bad[:6]:
                print("   ", b)
    ctx.close()


if __name__ == "__main__":
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    fails = run(cases, seed)
    print("mismatching cases: %d of %d" % (len(fails), cases))
    sys.exit(1 if fails else 0)
