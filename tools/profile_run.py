"""Small fixed workload for ncu / compute-sanitizer: one chunk of synthetic frames, device-resident, whole pipeline.
usage: python tools/profile_run.py [frames=64] [passes=2] [stage=-1]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

sp = load_package()
import importlib  # noqa: E402

synth = importlib.import_module("stair_perception_b200.synth")
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
stage = int(sys.argv[3]) if len(sys.argv) > 3 else -1
nd = min(frames, 8)
recs, Rs = synth.make_batch(nd)
Nb = 512 * 424 * 16
clouds = torch.from_numpy(np.tile(recs.view(np.uint8).reshape(nd, Nb), ((frames + nd - 1) // nd, 1))[:frames]).cuda()
R = torch.from_numpy(np.tile(Rs, ((frames + nd - 1) // nd, 1))[:frames].copy()).cuda()
torch.cuda.synchronize()
ctx = sp.Context(device=0, width=512, height=424, max_batch=frames)
ctx.set_trace(True)
for _ in range(passes):
    if stage < 0:
        res = ctx.process_frames(clouds, R, device_input=True, nframes=frames)
    else:
        ctx.run_stage_device(clouds.data_ptr(), R.data_ptr(), frames, stage)
        ctx.stage_ms()
print("stage ms:", ctx.stage_ms())
tr = ctx.trace_ms()
tot = sum(ms for _, ms in tr)
for nm, ms in tr:
    print("  %-24s %9.3f ms  %5.1f%%" % (nm, ms, 100 * ms / max(tot, 1e-9)))
print("  %-24s %9.3f ms" % ("total", tot))
if stage < 0:
    print("planes:", res["n_planes"][:8].tolist(), "status:", res["status"][:8].tolist())
ctx.close()
