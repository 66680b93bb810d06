#!/usr/bin/env python
"""Headline benchmark: frames/s of 512x424 synthetic Kinect-V2 stair clouds through the segmentation front-end.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the whole front-end (levelling .. sorted plane records) over ONE batch of 8192 frames (BASELINE.json config 5),
sharded frame-wise across the N GPUs (stair-perception_b200/shard.py): STRONG scaling.  Every GPU runs whole frames; the only
exchange is the all-gather of the compact plane records (64-byte header + n_planes x 224 B per frame), device to device.
`value`  : the batch resident in HBM; the step ends with the gathered records of all 8192 frames in host memory on every rank.
           The library runs the chunks of a rank's block on two lanes (context + twin workspace), see DESIGN.md section 2.
`e2e`    : the same step from pinned HOST clouds (H2D of every cloud and D2H of the records inside the timed region).
`roofline`: the fused ingest+normals kernel (K1K2) timed alone with CUDA events on the library's stream.
`config4`, `latency_b1` (N = 1): the dense 2 M-point kNN scene and the reference's own one-frame-per-call mode.
`cpu_baseline` / `--impl reference`: the CPU oracle (a port: the reference's PCL build cannot be compiled in this
image) on the box's host cores, on a bounded sample of the same synthetic frames.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 512, 424
N = W * H
K1K2_BYTES_PER_POINT = 16 + 12 + 12          # read xyzrgba, write levelled xyz + normals (fused K1+K2); the kernel stores both as float4 (48 B moved)
N_DISTINCT = 256                             # distinct synthetic frames (seeds 0..255); the 8192-frame batch cycles through them


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bind_to_gpu_numa_node(index):
    """Pin this process (and so its first-touch pinned host buffers) to the CPUs of the GPU's NUMA node: with one rank per GPU
    the H2D copies of the end-to-end leg then read local memory.  Best effort; returns the node or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        if isinstance(bus, bytes):
            bus = bus.decode()
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                       # nvml pads the PCI domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def load_traffic(kernel, frames_per_launch):
    """DRAM bytes of one launch of `kernel` from the newest committed ncu --set full capture (profiles/r*_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum); ncu cannot run inside the timed run, so the line says which capture"""
    import glob
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            with open(p) as f:
                t = json.load(f)
            if t["kernel"] == kernel and int(t["frames_per_launch"]) == int(frames_per_launch):
                return {"bytes": int(t["dram_bytes_read"]) + int(t["dram_bytes_write"]),
                        "source": "%s (%s)" % (os.path.basename(p), t.get("captured", "round 1 capture"))}
        except Exception:
            pass
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region"""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """call at the start of the timed region (the sampler is started earlier, during the warm-up, so that it is already
        delivering rows: nvidia-smi needs a few hundred ms to print its first one and an 8-GPU step lasts 27 ms)"""
        self.t0 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.perf_counter()
        t0 = getattr(self, "t0", 0.0)
        deadline = t1 + 1.5
        while time.perf_counter() < deadline and not any(t >= t1 for t, _ in self.rows):
            time.sleep(0.02)                                   # one row after the region closes it from the right
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        good = [(t, r) for t, r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        inside = [r for t, r in good if t0 <= t <= t1]
        where = "inside the timed region"
        if not inside:                                         # region shorter than the sampling period: the rows that bracket it
            before = [r for t, r in good if t < t0][-1:]
            after = [r for t, r in good if t > t1][:1]
            inside = before + after
            where = "the rows bracketing the timed region (it is shorter than the 100 ms sampling period); the GPU ran the same steps (warm-up) when the earlier one was taken"
        sm = [float(r[0]) for r in inside]
        mx = [float(r[1]) for r in inside if r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            for k, nm in enumerate(names):
                if r[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "sampled": where}


def _one_frame(i):
    import importlib
    synth = importlib.import_module("stair_perception_b200.synth")
    return synth.make_frame(i)


def make_frames(n, workers=1):
    """frames 0 .. n-1 of the seeded synthetic generator (identical on every rank); a fork pool when the host has the cores
    (call it before CUDA is initialised)"""
    from __graft_entry__ import load_package
    load_package()
    if workers > 1 and n > 8:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            out = pool.map(_one_frame, range(n), chunksize=max(1, n // (4 * workers)))
    else:
        out = [_one_frame(i) for i in range(n)]
    return np.stack([o[0] for o in out]), np.stack([o[1] for o in out])


def cpu_baseline(sp, recs, Rs, seconds_budget=12.0):
    """oracle on all host cores over a bounded sample of the same synthetic frames"""
    from oracle import orc
    cores = host_cores()
    nthreads = max(1, cores)
    reps = 1
    n = len(recs)
    t0 = time.perf_counter()
    orc.process_batch(sp.DEFAULT_PARAMS, recs[:min(n, nthreads)], W, H, Rs[:min(n, nthreads)], nthreads=nthreads)
    t_one = time.perf_counter() - t0
    reps = int(max(1, min(8, seconds_budget / max(t_one, 1e-3) / max(1, n // max(1, min(n, nthreads))))))
    t0 = time.perf_counter()
    frames = 0
    for _ in range(reps):
        orc.process_batch(sp.DEFAULT_PARAMS, recs, W, H, Rs, nthreads=nthreads)
        frames += n
    dt = time.perf_counter() - t0
    # the reference's live path is single-threaded (one frame at a time; its screenshots show 137 ms per frame): latency of the
    # same oracle on ONE core, per stage, over a few frames
    single = {}
    try:
        o = orc.Oracle(sp.DEFAULT_PARAMS)
        per = []
        for i in range(min(4, n)):
            t1 = time.perf_counter()
            o.process(recs[i], W, H, Rs[i], keep_taps=False)
            per.append((time.perf_counter() - t1) * 1e3)
            st = o.stage_ms()
            for k, v in st.items():
                single[k] = single.get(k, 0.0) + v / min(4, n)
        o.close()
        single = {"ms_per_frame": float(np.mean(per)), "frames": len(per), "stage_ms": {k: round(v, 3) for k, v in single.items()}}
    except Exception as e:                                   # the throughput figure above stands on its own
        single = {"error": str(e)}
    return {"value": frames / dt, "unit": "frames/s", "cores": nthreads, "kind": "port",
            "sample": "%d synthetic 512x424 stair frames x %d passes, oracle/liborc.so, one frame per thread, %d threads" % (n, reps, nthreads),
            "build": orc.build_flags(), "single_core": single}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path (PCL itself cannot be built here) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from __graft_entry__ import load_package
    sp = load_package()
    from oracle import orc
    orc.build()
    cores = host_cores()
    n = max(8, min(32, cores))
    recs, Rs = make_frames(n)
    passes = 8                                                  # one step = 8 passes over the distinct frames (about half a second of CPU work)
    for _ in range(max(0, args.warmup)):
        orc.process_batch(sp.DEFAULT_PARAMS, recs, W, H, Rs, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(passes):
            orc.process_batch(sp.DEFAULT_PARAMS, recs, W, H, Rs, nthreads=cores)
    dt = time.perf_counter() - t0
    val = n * passes * args.steps / dt
    sample = "%d synthetic 512x424 stair frames x %d passes per step, CPU oracle port (oracle/liborc.so), %d host threads" % (n, passes, cores)
    line = {"impl": "reference", "metric": "frames/sec of 512x424 stair clouds through segmentation; % HBM roofline", "value": val, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config 5: batch of 8192 synthetic Kinect-V2 4-step stair frames, 512x424 (217088 pts), levelling + full segmentation front-end",
                       "sample": "bounded sample of the 8192-frame batch (the CPU arm's throughput does not depend on the batch length)", "frames_per_step": n * passes},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample, "build": orc.build_flags()},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def bench_config4(sp, dev, peak):
    """BASELINE.json config 4: the dense 2 M-point unorganised scene (leaf 0.005, kNN normals, k = 69), one frame per call,
    device-resident input.  Roofline of k_knn_normals on both bases of SURVEY 8(d): compulsory 24 B/point (16 in, 12 out: the
    neighbours come from L2) and gather-inclusive 852 B/point (k x 12 B of neighbour coordinates + 24)."""
    import importlib
    import torch
    synth = importlib.import_module("stair_perception_b200.synth")
    n_points = 2_000_000
    rec = synth.make_dense_scene(n_points, seed=0)
    p = dict(sp.DEFAULT_PARAMS)
    p.update(synth.DENSE_PARAMS)
    ctx = sp.Context(p, device=dev.index, width=n_points, height=1, max_batch=1)
    ctx.set_normal_mode(1)
    d_rec = torch.from_numpy(rec.view(np.uint8).reshape(1, -1)).to(dev)
    torch.cuda.synchronize()
    res = np.zeros(1, dtype=sp.FRAME_DTYPE)
    for _ in range(3):
        ctx.process_frames(d_rec, None, device_input=True, nframes=1, results=res)
    t0 = time.perf_counter()
    reps = 8
    for _ in range(reps):
        ctx.process_frames(d_rec, None, device_input=True, nframes=1, results=res)
    ms = (time.perf_counter() - t0) * 1e3 / reps
    ctx.set_trace(True)
    ctx.process_frames(d_rec, None, device_input=True, nframes=1, results=res)
    tr = ctx.trace_ms()
    ctx.set_trace(False)
    knn_ms = dict(tr).get("k_knn_normals", 0.0)
    n_valid = int(res[0]["n_valid"])
    out = {"workload": "config 4: dense 2 M-point unorganised stair scene, leaf 0.005, kNN normals (k = %d), one frame per call" % p["normal_compute_points"],
           "value": 1e3 / ms, "unit": "frames/s", "ms_per_frame": ms, "points_per_s": n_points * 1e3 / ms, "n_planes": int(res[0]["n_planes"]),
           "kernels": [{"kernel": k, "ms": round(t, 4)} for k, t in sorted(tr, key=lambda kv: -kv[1])[:6]]}
    if knn_ms > 0:
        k = int(p["normal_compute_points"])
        for name, bpp in (("compulsory", 24), ("with_neighbour_gathers", 24 + 12 * k)):
            ach = bpp * n_valid / (knn_ms * 1e-3) / 1e9
            out["roofline_knn_" + name] = {"bound": "hbm", "kernel": "k_knn_normals", "bytes_per_point": bpp, "achieved": ach, "peak": peak,
                                           "unit": "GB/s", "frac": ach / peak, "ms_per_launch": knn_ms, "points": n_valid}
    ctx.close()
    return out


def bench_latency_b1(sp, dev, recs, Rs):
    """The reference's own mode of use: one frame per call (test/processor.cpp:156-174), host cloud in, host record out, through
    sp_process_frames on a max_batch = 1 context; median over 30 calls after 5 warm-up calls.  Beside it: the published 137 ms
    (pic/screenshot1.png, unknown CPU) and the oracle's single-core time in cpu_baseline.single_core."""
    ctx = sp.Context(device=dev.index, width=W, height=H, max_batch=1)
    out = {"published_reference_ms": 137.0, "api": "sp_process_frames, max_batch = 1, pageable host cloud -> host record"}
    cases = [("synthetic_frame_0", recs[0], Rs[0].reshape(1, 9))]
    pcd = os.path.join(ROOT, "tests", "golden", "samples", "0001_cloud.pcd")
    if os.path.exists(pcd):
        rec, w, h = sp.read_pcd(pcd)
        cases.append(("samples_0001", rec, None))
    for name, rec, R in cases:
        for _ in range(5):
            ctx.process_frames(rec, R)
        ts = []
        for _ in range(30):
            t0 = time.perf_counter()
            r = ctx.process_frames(rec, R)
            ts.append((time.perf_counter() - t0) * 1e3)
        t1 = []
        for _ in range(10):
            t0 = time.perf_counter()
            ctx.process_frames(rec, R)
            ctx.stair_model(0)
            t1.append((time.perf_counter() - t0) * 1e3)
        t2 = []
        for _ in range(10):
            t0 = time.perf_counter()
            ctx.process_frames(rec, R)
            ctx.stair_model_batch(0, 1, nthreads=1)
            t2.append((time.perf_counter() - t0) * 1e3)
        out[name] = {"ms_median": float(np.median(ts)), "ms_min": float(np.min(ts)), "ms_with_stair_model_median": float(np.median(t1)),
                     "ms_with_packed_stair_model_median": float(np.median(t2)), "n_planes": int(r[0]["n_planes"])}
    ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--frames", type=int, default=8192, help="frames of the WHOLE batch per step (config 5: 8192), sharded frame-wise across the GPUs")
    ap.add_argument("--e2e-frames", type=int, default=8192, help="frames of the whole batch per end-to-end step (pinned host input, 28.5 GB over all ranks)")
    ap.add_argument("--max-batch", type=int, default=0, help="frames per chunk of the library (0 = 512 when a rank runs at least 1024 frames, else 256)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip config 4, single-frame latency, the weak-scaling figure and the depth-image leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries ONE JSON line: everything libraries print there while the run lasts (NCCL's version banner, for one) goes to
    # stderr instead; the descriptor is handed back just before the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # the distinct frames first (fork pool, before CUDA exists in this process); identical on every rank (seeded)
    recs, Rs = make_frames(N_DISTINCT, workers=max(1, min(32, host_cores() // max(1, world))))

    import importlib
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    sp = load_package()
    shard = importlib.import_module("stair_perception_b200.shard")

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"                 # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # ---- the batch: frame i of the 8192 = distinct frame i % 256; this rank holds (and runs) its contiguous block ---------------
    F = args.frames
    lo, hi = shard.shard_bounds(F, rank, world)
    Fl = hi - lo
    if args.max_batch <= 0:                                   # the chunk size is a library knob, not part of the workload: 512-frame chunks
        args.max_batch = 512 if Fl >= 1024 else 256           # 2-3 % faster than 256 from two chunks per rank on (38.3 k against 37.6 k frames/s at 1024 frames)
    base = torch.from_numpy(recs.view(np.uint8).reshape(N_DISTINCT, N * 16)).to(dev)
    base_R = torch.from_numpy(Rs).to(dev)
    idx = (torch.arange(lo, hi, device=dev) % N_DISTINCT)
    d_clouds = base.index_select(0, idx).contiguous()          # Fl x 3.47 MB resident in HBM (>> L2)
    d_R = base_R.index_select(0, idx).contiguous()
    torch.cuda.synchronize()

    ctx = sp.Context(device=local, width=W, height=H, max_batch=args.max_batch)
    cap = int(sp.lib().sp_compact_bytes_max(shard.shard_sizes(F, world)[0]))
    cbuf = torch.empty(cap, dtype=torch.uint8, device=dev)     # this rank's compact records (device)
    h_all = torch.empty(int(sp.lib().sp_compact_bytes_max(F)), dtype=torch.uint8, pin_memory=True)   # the whole batch's, on the host
    last = {}

    def step_device():
        # shard -> whole frames on this GPU, records stay on the device -> all-gather of the compact records -> host
        _, nb = ctx.process_frames_compact(d_clouds, d_R, device_input=True, nframes=Fl, out=cbuf)
        # (every rank ends with the whole batch's records on its GPU; rank 0, the consumer, also copies them to its host)
        last["packed"], last["sizes"] = shard.gather_compact(cbuf, nb, host_out=h_all, to_host=rank == 0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = max(e0.elapsed_time(e1), 0.0)
        ms = max(ms, wall * 0.999) if ms < wall * 0.5 else ms     # the library syncs its own streams inside fn
        barrier()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                        # already running through the warm-up: see ClockSampler.mark
    for _ in range(max(3, args.warmup)):
        step_device()
    sampler.mark()
    l0 = ctx.launch_count()
    ms = timed(step_device, args.steps)
    launches = ctx.launch_count() - l0
    stage_full = ctx.stage_ms()
    clocks = sampler.stop() if rank == 0 else None
    value = F * args.steps / (ms / 1e3)
    # what came back: every frame of the batch, in batch order, on every rank (ranks other than 0 hold it on their GPU: copy it out now, untimed)
    if not isinstance(last["packed"], np.ndarray):
        last["packed"] = last["packed"].cpu().numpy()
    results = sp.expand_compact(last["packed"], F)
    assert int(results["n_planes"].min()) >= 1 and int(results["status"].max()) == 0, "unexpected empty results"
    assert np.array_equal(results[:N_DISTINCT].view(np.uint8), results[N_DISTINCT:2 * N_DISTINCT].view(np.uint8)) or F < 2 * N_DISTINCT
    # the exchange alone (sizes + compact records + copy to the host), device-timed
    gather_ms = None
    gather_all_ms = None
    gather_parts = {}
    if world > 1:
        _, nb = ctx.process_frames_compact(d_clouds, d_R, device_input=True, nframes=Fl, out=cbuf)
        gather_ms = timed(lambda: shard.gather_compact(cbuf, nb, host_out=h_all, to_host=rank == 0), 5) / 5
        shard.gather_compact(cbuf, nb, host_out=h_all, parts=gather_parts, to_host=rank == 0)
        gather_all_ms = timed(lambda: shard.gather_compact(cbuf, nb, host_out=h_all), 5) / 5

    # the dominant streaming kernel alone (K1K2 = ingest + levelling + masking + organised normals), CUDA events on its stream
    nb = min(256, args.max_batch, Fl)                         # 256 frames per launch: what the committed ncu captures ran
    k_ms = []
    for i in range(3 + 10):
        off = (i % max(1, Fl // nb)) * nb
        ctx.run_stage_device(d_clouds.data_ptr() + off * N * 16, d_R.data_ptr() + off * 36, nb, 0)
        t = ctx.stage_ms()["ingest_normals"]
        if i >= 3:
            k_ms.append(t)
    k1 = float(np.mean(k_ms))
    peak, peak_src = load_peaks()
    achieved = K1K2_BYTES_PER_POINT * N * nb / (k1 * 1e-3) / 1e9
    vox_ms = []
    for i in range(2 + 5):
        ctx.run_stage_device(d_clouds.data_ptr(), d_R.data_ptr(), nb, 1)
        if i >= 2:
            vox_ms.append(ctx.stage_ms()["voxel"])
    # per-kernel times of one chunk (the library's own trace: a CUDA event after every launch on its stream), outside any timed region
    res_nb = np.zeros(nb, dtype=sp.FRAME_DTYPE)
    ctx.set_trace(True)
    ctx.process_frames(d_clouds, d_R, device_input=True, nframes=nb, results=res_nb)
    trace = ctx.trace_ms()
    ctx.set_trace(False)
    chunk_ms = sum(t for _, t in trace)
    # the voxel stage's streaming kernel: reads every levelled point (16 B), writes (key, pixel) = 8 B per cropped point
    tdict = {}
    for k, t in trace:
        tdict[k] = tdict.get(k, 0.0) + t
    n_crop = int(res_nb["n_crop"].sum())
    n_vox = int(res_nb["n_voxel"].sum())
    roof_voxel = None
    vk_ms = tdict.get("k_vx_key", 0.0)
    if vk_ms > 0:
        vk_bytes = 16 * N * nb + 8 * n_crop
        roof_voxel = {"bound": "hbm", "kernel": "k_vx_key", "achieved": vk_bytes / (vk_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                      "frac": vk_bytes / (vk_ms * 1e-3) / 1e9 / peak, "ms_per_launch": vk_ms, "algorithmic_bytes_per_launch": vk_bytes,
                      "note": "timed inside a traced chunk (one CUDA event before and after the launch on the library's stream)"}
    # the whole voxel stage against its compulsory bytes (12 B per point read + 4 B per voxel written)
    vstage_ms = float(np.mean(vox_ms))
    vs_bytes = 12 * N * nb + 4 * n_vox
    roof_vstage = {"bound": "hbm", "kernel": "voxel stage (grid + key + sort + pick + lists)", "achieved": vs_bytes / (vstage_ms * 1e-3) / 1e9, "peak": peak,
                   "unit": "GB/s", "frac": vs_bytes / (vstage_ms * 1e-3) / 1e9 / peak, "ms_per_launch": vstage_ms, "algorithmic_bytes_per_launch": vs_bytes}
    kernels = [{"kernel": k, "ms": round(t, 4), "share": round(t / max(chunk_ms, 1e-9), 4)} for k, t in sorted(tdict.items(), key=lambda kv: -kv[1])[:10]]

    # ---- end to end: the same batch from pinned HOST memory (H2D of the clouds and D2H of the records inside the timed region) ----
    Fe = min(args.e2e_frames, F)
    elo, ehi = shard.shard_bounds(Fe, rank, world)
    Fel = ehi - elo
    try:
        h_clouds = torch.empty((Fel, N * 16), dtype=torch.uint8, pin_memory=True)
    except RuntimeError:                                       # not enough lockable host memory for 28.5 GB: half the batch
        Fe = Fe // 2
        elo, ehi = shard.shard_bounds(Fe, rank, world)
        Fel = ehi - elo
        h_clouds = torch.empty((Fel, N * 16), dtype=torch.uint8, pin_memory=True)
    h_R = torch.empty((Fel, 9), dtype=torch.float32, pin_memory=True)
    base_h = torch.from_numpy(recs.view(np.uint8).reshape(N_DISTINCT, N * 16))
    R_h = torch.from_numpy(Rs)
    for i in range(Fel):
        h_clouds[i].copy_(base_h[(elo + i) % N_DISTINCT])
        h_R[i].copy_(R_h[(elo + i) % N_DISTINCT])

    def step_e2e():
        _, nbytes = ctx.process_frames_compact(h_clouds, h_R, device_input=False, nframes=Fel, out=cbuf)
        last["packed_e"], _ = shard.gather_compact(cbuf, nbytes, host_out=h_all, to_host=rank == 0)

    for _ in range(3):
        step_e2e()
    ms_e = timed(step_e2e, args.steps)
    e2e_val = Fe * args.steps / (ms_e / 1e3)
    if not isinstance(last["packed_e"], np.ndarray):
        last["packed_e"] = last["packed_e"].cpu().numpy()
    res_e = sp.expand_compact(last["packed_e"], Fe)
    assert np.array_equal(res_e.view(np.uint8), results[:Fe].view(np.uint8)), "host-input records differ from device-input records"
    d2h_e = int(last["packed_e"].size)

    extras = {}
    if not args.no_extras:
        # weak scaling for reference next to the strong figure: every GPU runs the whole 8192-frame batch (no exchange)
        if world > 1:
            del d_clouds
            torch.cuda.empty_cache()
            idx = (torch.arange(0, F, device=dev) % N_DISTINCT)
            d_all = base.index_select(0, idx).contiguous()
            d_Rall = base_R.index_select(0, idx).contiguous()
            cbig = torch.empty(int(sp.lib().sp_compact_bytes_max(F)), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()

            def step_weak():
                ctx.process_frames_compact(d_all, d_Rall, device_input=True, nframes=F, out=cbig)

            step_weak()
            ms_w = timed(step_weak, 2)
            extras["weak_scaling"] = {"value": F * world * 2 / (ms_w / 1e3), "unit": "frames/s", "frames_per_gpu_per_step": F, "ms_per_step": ms_w / 2,
                                      "note": "every GPU runs its own 8192-frame batch, no exchange (round 1's figure)"}
            del d_all, cbig
            torch.cuda.empty_cache()
        # the same frames entering as what the sensor delivers (uint16 depth in mm + registered colour, 6 B / pixel instead of 16):
        # sp_process_depth_frames back-projects on the device (K2G::updateCloud), everything else is identical
        synth = importlib.import_module("stair_perception_b200.synth")
        Fd = min(4096, Fe)
        dlo, dhi = shard.shard_bounds(Fd, rank, world)
        Fdl = dhi - dlo
        nd = min(32, N_DISTINCT)
        dframes = [synth.make_depth_frame(i) for i in range(nd)]
        h_depth = torch.empty((Fdl, H, W), dtype=torch.uint16, pin_memory=True)
        h_bgrx = torch.empty((Fdl, H, W, 4), dtype=torch.uint8, pin_memory=True)
        h_Rd = torch.empty((Fdl, 9), dtype=torch.float32, pin_memory=True)
        for i in range(Fdl):
            h_depth[i].copy_(torch.from_numpy(dframes[(dlo + i) % nd][0]))
            h_bgrx[i].copy_(torch.from_numpy(dframes[(dlo + i) % nd][1]))
            h_Rd[i].copy_(torch.from_numpy(dframes[(dlo + i) % nd][2]))
        res_d = np.zeros(Fdl, dtype=sp.FRAME_DTYPE)

        def step_depth():
            ctx.process_depth_frames(h_depth, synth.INTRINSICS, bgrx=h_bgrx, mirror=True, R=h_Rd, nframes=Fdl, results=res_d)

        for _ in range(2):
            step_depth()
        ms_d = timed(step_depth, args.steps)
        assert int(res_d["n_planes"].min()) >= 1 and int(res_d["status"].max()) == 0
        extras["e2e_depth_input"] = {"value": Fd * args.steps / (ms_d / 1e3), "unit": "frames/s", "h2d_bytes_per_step": int(Fd * (N * 6 + 36)),
                                     "d2h_bytes_per_step": int(Fd * sp.FRAME_DTYPE.itemsize), "frames_per_step": Fd, "ms_per_step": ms_d / args.steps,
                                     "note": "%d distinct scenes entering as uint16 depth + registered colour (6 B/pixel) through sp_process_depth_frames, "
                                             "sharded like the batch, fixed-size records to the host, no exchange" % nd}
        del h_depth, h_bgrx
        # the consumer on the measured path: chunk by chunk, the front-end on device-resident clouds, then the batched stair model
        # (sp_stair_model_batch: one pack kernel + one copy of the chunk's plane clouds, host model on the box's cores)
        nsm = min(args.max_batch, Fl)
        res_sm = np.zeros(nsm, dtype=sp.FRAME_DTYPE)
        ctx_sm = sp.Context(device=local, width=W, height=H, max_batch=nsm)

        def step_stair():
            for c0 in range(0, 4 * nsm, nsm):
                o = (c0 % max(nsm, Fl - nsm + 1))
                ctx_sm.process_frames(d_all_or_mine[o:o + nsm], d_R_all_or_mine[o:o + nsm], device_input=True, nframes=nsm, results=res_sm)
                last["stairs"] = ctx_sm.stair_model_batch(0, nsm)

        d_all_or_mine, d_R_all_or_mine = (d_clouds, d_R) if world == 1 else (base.index_select(0, (torch.arange(lo, hi, device=dev) % N_DISTINCT)), base_R.index_select(0, (torch.arange(lo, hi, device=dev) % N_DISTINCT)))
        step_stair()
        ms_s = timed(step_stair, 2)
        serial = 4 * nsm * world * 2 / (ms_s / 1e3)
        # the same with the host model of chunk i running while the GPU works on chunk i + 1 (two contexts, one worker thread)
        from concurrent.futures import ThreadPoolExecutor
        ctx_sm2 = sp.Context(device=local, width=W, height=H, max_batch=nsm)
        res_sm2 = np.zeros(nsm, dtype=sp.FRAME_DTYPE)
        pair = ((ctx_sm, res_sm), (ctx_sm2, res_sm2))
        pool = ThreadPoolExecutor(1)
        NCH = 8

        def step_stair_overlapped():
            futs = [None, None]
            for ci in range(NCH):
                cx, rs = pair[ci & 1]
                if futs[ci & 1] is not None:
                    last["stairs"] = futs[ci & 1].result()
                o = ((ci * nsm) % max(nsm, Fl - nsm + 1))
                cx.process_frames(d_all_or_mine[o:o + nsm], d_R_all_or_mine[o:o + nsm], device_input=True, nframes=nsm, results=rs)
                futs[ci & 1] = pool.submit(cx.stair_model_batch, 0, nsm)
            for fu in futs:
                if fu is not None:
                    last["stairs"] = fu.result()

        step_stair_overlapped()
        ms_o = timed(step_stair_overlapped, 2)
        pool.shutdown()
        extras["e2e_with_stair_model"] = {"value": NCH * nsm * world * 2 / (ms_o / 1e3), "unit": "frames/s", "frames_per_chunk": nsm, "ms_per_chunk": ms_o / (2 * NCH),
                                          "stairs_found": int(last["stairs"]["has_stair"].sum()), "host_threads": host_cores(),
                                          "one_chunk_at_a_time": {"value": serial, "ms_per_chunk": ms_s / 8},
                                          "note": "chunk by chunk: front-end on device-resident clouds, records to the host, then sp_stair_model_batch "
                                                  "(StairDetection::stairModel for every frame of the chunk) on the host cores WHILE the GPU runs the next chunk "
                                                  "(two contexts, one worker thread); one_chunk_at_a_time = the same without that overlap"}
        ctx_sm2.close()
        ctx_sm.close()
        if rank == 0 and world == 1:
            extras["config4"] = bench_config4(sp, dev, peak)
            extras["latency_b1"] = bench_latency_b1(sp, dev, recs, Rs)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        nc = max(8, min(32, host_cores()))
        cpu = cpu_baseline(sp, recs[:nc], Rs[:nc])

    if rank == 0:
        traffic = load_traffic("k_ingest_normals", nb)
        line = {
            "metric": "frames/sec of 512x424 stair clouds through segmentation; % HBM roofline",
            "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config 5: batch of 8192 synthetic Kinect-V2 4-step stair frames, 512x424 (217088 pts), levelling + full segmentation front-end",
                       "frames_per_step": F, "frames_per_gpu_per_step": Fl, "distinct_frames": N_DISTINCT, "max_batch": args.max_batch,
                       "lanes": 1 if os.environ.get("SP_NO_TWIN", "") not in ("", "0") else 2,
                       "l2": "inputs larger than L2 (%.1f GB resident per GPU)" % (Fl * N * 16 / 1e9),
                       "parallelism": "ONE %d-frame batch sharded frame-wise across %d GPU(s) (shard.shard_bounds), whole frames per GPU, "
                                      "all-gather of the compact plane records only (shard.gather_compact)" % (F, world),
                       "api": "sp_process_frames_compact (records stay on the device until the exchange) + all-gather on the GPUs + one D2H of the gathered records on rank 0",
                       "host_numa_binding": numa_node},
            "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": int(Fe * (N * 16 + 36)),
                    "d2h_bytes_per_step": d2h_e, "frames_per_step": Fe, "ms_per_step": ms_e / args.steps,
                    "api": "sp_process_frames_compact from pinned host clouds (H2D inside), exchange, compact records to the host (D2H inside)"},
            "gather": {"ms": gather_ms, "parts_ms_rank0": {k: round(v, 4) for k, v in gather_parts.items()}, "bytes_per_rank": [int(v) for v in last["sizes"]],
                       "ms_with_host_copy_on_every_rank": gather_all_ms,
                       "note": "sizes all-gather + padded all-gather of the compact records (NCCL, device buffers): the whole batch's records end on every GPU, and rank 0 (the consumer) copies them to its host inside the step; ms_with_host_copy_on_every_rank = the same exchange with N copies of the same bytes to the one host, as the mid-round lines measured it; device-timed"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_ingest_normals", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic["bytes"] if traffic else None, "traffic_source": traffic["source"] if traffic else None,
                         "peak_source": peak_src, "ms_per_launch": k1, "frames_per_launch": nb,
                         "algorithmic_bytes_per_launch": K1K2_BYTES_PER_POINT * N * nb},
            "roofline_voxel_key": roof_voxel,
            "roofline_voxel_stage": roof_vstage,
            "kernels_of_a_chunk": {"frames": nb, "ms": round(chunk_ms, 4), "top": kernels},
            "stage_ms_last_chunk": stage_full, "voxel_stage_ms": vstage_ms,
            "cpu_baseline": cpu,
        }
        line.update(extras)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
