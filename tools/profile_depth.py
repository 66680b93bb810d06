"""Depth-image ingest (row f2) for ncu: a batch of uint16 depth + registered colour frames through sp_process_depth_frames.
usage: python tools/profile_depth.py [frames=64] [passes=2]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

sp = load_package()
synth = importlib.import_module("stair_perception_b200.synth")
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
one = [synth.make_depth_frame(i) for i in range(4)]
depth = np.stack([one[i % 4][0] for i in range(frames)])
bgrx = np.stack([one[i % 4][1] for i in range(frames)])
R = np.stack([one[i % 4][2].reshape(9) for i in range(frames)]).astype(np.float32)
ctx = sp.Context(device=0, width=512, height=424, max_batch=frames)
for _ in range(passes):
    res = ctx.process_depth_frames(depth, synth.INTRINSICS, bgrx=bgrx, mirror=True, R=R)
print("depth ingest: %d frames, planes of frame 0: %d, status %d" % (frames, int(res[0]["n_planes"]), int(res[0]["status"])))
ctx.close()
