"""A/B of the two K1K2 kernels (row-marching TMA kernel vs the tile kernel, SP_K1=tile) on the same inputs: where do the levelled
points, the normals and the crop statistics differ?   python tools/debug_k1.py   (GPU box)"""
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


VARIANT = os.environ.get("K1_VARIANT", "tma")             # the variant compared against the plain tile kernel


def ctx_pair(w, h, b, params=None, variant=None):
    os.environ["SP_K1"] = "tile"
    a = sp.Context(params, device=0, width=w, height=h, max_batch=b)
    os.environ["SP_K1"] = variant or VARIANT
    n = sp.Context(params, device=0, width=w, height=h, max_batch=b)
    del os.environ["SP_K1"]
    return a, n


def diff(name, clouds, R, w, h, params=None, show=1):
    nf = len(clouds)
    old, new = ctx_pair(w, h, nf, params)
    ro = old.process_frames(clouds, R, show_mode=show).copy()
    rn = new.process_frames(clouds, R, show_mode=show).copy()
    bad = 0
    for f in range(nf):
        for k in ("status", "n_valid", "n_crop"):
            if int(ro[f][k]) != int(rn[f][k]):
                print("  %s frame %d %s: tile %d other %d" % (name, f, k, int(ro[f][k]), int(rn[f][k])))
                bad += 1
        for tap in ("xyz", "normals", "minmax"):
            a, b = old.tap(f, tap), new.tap(f, tap)
            na, nb = np.isnan(a), np.isnan(b)
            neq = (na != nb) | (~na & ~nb & (a != b))
            if neq.any():
                bad += 1
                if tap == "minmax":
                    print("  %s frame %d minmax: %s vs %s" % (name, f, a, b))
                    continue
                rows = np.flatnonzero(neq.any(axis=1))
                rr, cc = rows // w, rows % w
                print("  %s frame %d %s: %d pixels differ (nan pattern %d); rows %d..%d cols %d..%d; first %s" % (
                    name, f, tap, rows.size, int((na != nb).any(axis=1).sum()), rr.min(), rr.max(), cc.min(), cc.max(),
                    [(int(r), int(c)) for r, c in zip(rr[:6], cc[:6])]))
                for p in rows[:3]:
                    print("      (%d,%d) tile %s strip %s" % (p // w, p % w, a[p], b[p]))
    print("%-28s %s" % (name, "OK" if bad == 0 else "%d MISMATCHES" % bad))
    old.close()
    new.close()
    return bad


def main():
    total = 0
    rec, w, h = sp.read_pcd(os.path.join(ROOT, "tests", "golden", "samples", "0001_cloud.pcd"))
    total += diff("samples/0001", rec[None], None, w, h)
    rec0, _, _ = sp.read_pcd(os.path.join(ROOT, "tests", "golden", "samples", "0000_cloud.pcd"))
    total += diff("samples/0000", rec0[None], None, w, h)
    frames = [synth.make_frame(i) for i in (3, 11, 12)]
    clouds = np.stack([f[0] for f in frames])
    Rs = np.stack([f[1] for f in frames])
    total += diff("synthetic x3 levelled", clouds, Rs, 512, 424)
    img = rec.reshape(h, w)
    for ww, hh in ((200, 150), (97, 61), (33, 13), (512, 11), (300, 40)):
        sub = np.ascontiguousarray(img[120:120 + hh, min(200, w - ww):min(200, w - ww) + ww]).reshape(-1)
        total += diff("window %dx%d" % (ww, hh), np.stack([sub, sub]), None, ww, hh)
    big = frames[2][0].copy()
    for k in ("x", "y", "z"):
        big[k] = big[k] * np.float32(4.0)
    tiny = frames[2][0].copy()
    rng = np.random.default_rng(5)
    idx = rng.choice(tiny.shape[0], 4000, replace=False)
    tiny["x"][idx[:2000]] = np.float32(3e-9)
    tiny["z"][idx[2000:]] = np.float32(-7e-8)
    p = dict(sp.DEFAULT_PARAMS)
    p.update(x_min=-20.0, x_max=20.0, y_min=-20.0, y_max=20.0, z_min=-20.0, z_max=20.0)
    total += diff("inexact coordinates", np.stack([big, tiny]), np.stack([frames[2][1]] * 2), 512, 424, p)
    # timing of the kernel alone on a 256-frame chunk
    import torch
    fr = [synth.make_frame(i) for i in range(32)]
    base = torch.from_numpy(np.stack([f[0] for f in fr]).view(np.uint8).reshape(32, -1)).cuda()
    d_c = base.repeat(8, 1).contiguous()
    d_R = torch.from_numpy(np.stack([f[1] for f in fr])).cuda().repeat(8, 1).contiguous()
    torch.cuda.synchronize()
    old, new = ctx_pair(512, 424, 256, variant="tma")
    os.environ["SP_K1"] = "strip"
    third = sp.Context(device=0, width=512, height=424, max_batch=256)
    del os.environ["SP_K1"]
    for nm, c in (("tile", old), ("tma", new), ("strip", third)):
        ts = []
        for it in range(8):
            c.run_stage_device(d_c.data_ptr(), d_R.data_ptr(), 256, 0)
            ts.append(c.stage_ms()["ingest_normals"])
        print("k_ingest_normals (%s): %.3f ms per 256 frames (min %.3f)" % (nm, float(np.mean(ts[3:])), min(ts[3:])))
    print("TOTAL MISMATCHES", total)
    return 1 if total else 0


if __name__ == "__main__":
    sys.exit(main())
