"""Config 4 (2 M-point unorganised dense scene, kNN normals) once through the pipeline with the per-kernel trace.
usage: python tools/profile_knn.py [n_points=2000000] [passes=3]"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

sp = load_package()
synth = importlib.import_module("stair_perception_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rec = synth.make_dense_scene(n, seed=0)
p = dict(sp.DEFAULT_PARAMS)
p.update(synth.DENSE_PARAMS)
ctx = sp.Context(p, device=0, width=n, height=1, max_batch=1)
ctx.set_normal_mode(1)
ctx.set_trace(True)
for _ in range(passes):
    t0 = time.perf_counter()
    res = ctx.process_frames(rec)[0]
    dt = time.perf_counter() - t0
print("config 4: %d points, %d planes, host-to-host %.2f ms" % (n, int(res["n_planes"]), dt * 1e3))
tr = ctx.trace_ms()
tot = sum(ms for _, ms in tr)
for nm, ms in tr:
    print("  %-24s %9.3f ms  %5.1f%%" % (nm, ms, 100 * ms / max(tot, 1e-9)))
print("  %-24s %9.3f ms" % ("total", tot))
ctx.close()
