"""single-frame latency (sp_process_frames, max_batch = 1, host cloud in, host record out) with and without the CUDA graph of
the chunk pipeline:  python tools/latency_b1.py   (GPU box)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

sp = load_package()
import importlib  # noqa: E402

synth = importlib.import_module("stair_perception_b200.synth")
rec, R = synth.make_frame(0)
pcd, w, h = sp.read_pcd(os.path.join(ROOT, "tests", "golden", "samples", "0001_cloud.pcd"))
for graph in ("0", "8"):
    os.environ["SP_GRAPH"] = graph
    ctx = sp.Context(device=0, width=512, height=424, max_batch=1)
    for name, c, r in (("synthetic", rec, R.reshape(1, 9)), ("samples/0001", pcd, None)):
        for _ in range(5):
            ref = ctx.process_frames(c, r).copy()
        ts = []
        for _ in range(50):
            t0 = time.perf_counter()
            out = ctx.process_frames(c, r)
            ts.append((time.perf_counter() - t0) * 1e3)
        assert out.tobytes() == ref.tobytes()
        print("SP_GRAPH=%s %-14s median %.3f ms  min %.3f ms  planes %d" % (graph, name, float(np.median(ts)), min(ts), int(out[0]["n_planes"])))
    ctx.close()

# where the millisecond goes: the same frame from pinned host memory and from device memory
import torch  # noqa: E402

os.environ["SP_GRAPH"] = "8"
ctx = sp.Context(device=0, width=512, height=424, max_batch=1)
ctx.set_trace(True)
pin = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy()).pin_memory()
dev = pin.cuda()
Rd = torch.from_numpy(R.reshape(1, 9).copy()).cuda()
torch.cuda.synchronize()
for name, fn in (("pageable host", lambda: ctx.process_frames(rec, R.reshape(1, 9))),
                 ("pinned host", lambda: ctx.process_frames(pin.numpy().view(rec.dtype), R.reshape(1, 9))),
                 ("device", lambda: ctx.process_frames(dev, Rd, device_input=True, nframes=1))):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(50):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    print("input in %-14s median %.3f ms  min %.3f ms" % (name, float(np.median(ts)), min(ts)))
print("stage ms:", ctx.stage_ms())
tr = ctx.trace_ms()
print("kernel sum %.3f ms: " % sum(ms for _, ms in tr) + ", ".join("%s %.0f us" % (n, 1e3 * ms) for n, ms in tr if ms > 0.008))
ctx.close()
