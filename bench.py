#!/usr/bin/env python
"""Headline benchmark: frames/s of 512x424 synthetic Kinect-V2 stair clouds through the segmentation front-end.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the whole front-end (levelling .. sorted plane records) over this rank's batch of frames.
`value`  : device-resident input (HBM), includes the D2H of the per-frame plane records and, for N > 1, the gather.  The
           library runs the chunks of a batch on two lanes (context + twin workspace), see DESIGN.md section 2.
`e2e`    : the same metric through sp_process_frames with pinned HOST buffers (H2D + D2H inside the timed region).
`roofline`: the fused ingest+normals kernel (K1K2) timed with CUDA events on the library's stream.
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
N_DISTINCT = 32                              # distinct synthetic frames; the batch tiles them


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
    """dram bytes of one launch from the committed ncu capture (profiles/r01_traffic.json), or None"""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(p) as f:
            t = json.load(f)
        if t["kernel"] == kernel and int(t["frames_per_launch"]) == int(frames_per_launch):
            return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
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
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 8:
                for k, nm in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_frames(n):
    from __graft_entry__ import load_package
    load_package()
    import importlib
    synth = importlib.import_module("stair_perception_b200.synth")
    return synth.make_batch(n)


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
            "single_core": single}


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
    n = max(8, min(N_DISTINCT, cores))
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
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config 5: batch of synthetic Kinect-V2 4-step stair frames, 512x424 (217088 pts), levelling + full segmentation front-end",
                       "sample": "bounded sample of the 8192-frame batch", "frames_per_step": n * passes},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--frames", type=int, default=8192, help="frames per GPU per step (config 5: 8192)")
    ap.add_argument("--e2e-frames", type=int, default=4096, help="frames per GPU per end-to-end step (pinned host input, 14.2 GB)")
    ap.add_argument("--max-batch", type=int, default=256)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries ONE JSON line: everything libraries print there while the run lasts (NCCL's version banner, for one) goes to
    # stderr instead; the descriptor is handed back just before the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    sp = load_package()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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

    F = args.frames
    recs, Rs = make_frames(N_DISTINCT)                         # identical on every rank (seeded)
    base = torch.from_numpy(recs.view(np.uint8).reshape(N_DISTINCT, N * 16)).to(dev)
    reps = (F + N_DISTINCT - 1) // N_DISTINCT
    d_clouds = base.repeat(reps, 1)[:F].contiguous()           # F x 3.47 MB resident in HBM (>> L2)
    d_R = torch.from_numpy(Rs).to(dev).repeat(reps, 1)[:F].contiguous()
    del base
    Fe = min(args.e2e_frames, F, max(256, 8192 // world))      # pinned host memory: 14 GB at N=1, 3.6 GB per rank at N=8
    h_clouds = torch.empty((Fe, N * 16), dtype=torch.uint8).pin_memory()
    h_R = torch.empty((Fe, 9), dtype=torch.float32).pin_memory()
    base_h = torch.from_numpy(recs.view(np.uint8).reshape(N_DISTINCT, N * 16))
    R_h = torch.from_numpy(Rs)
    for i0 in range(0, Fe, N_DISTINCT):                        # tile the distinct frames without a 14 GB temporary
        n = min(N_DISTINCT, Fe - i0)
        h_clouds[i0:i0 + n].copy_(base_h[:n])
        h_R[i0:i0 + n].copy_(R_h[:n])
    torch.cuda.synchronize()

    ctx = sp.Context(device=local, width=W, height=H, max_batch=args.max_batch)
    results = np.zeros(F, dtype=sp.FRAME_DTYPE)
    res_e = np.zeros(Fe, dtype=sp.FRAME_DTYPE)
    gather_buf = None
    if world > 1:
        gather_buf = torch.empty((world, F * sp.FRAME_DTYPE.itemsize), dtype=torch.uint8, device=dev)

    def step_device():
        ctx.process_frames(d_clouds, d_R, device_input=True, nframes=F, results=results)
        if world > 1:   # gather of the per-frame plane records (the only inter-GPU exchange of this path)
            mine = torch.from_numpy(results.view(np.uint8)).to(dev, non_blocking=True)
            dist.all_gather_into_tensor(gather_buf.view(-1), mine)

    def step_e2e():
        ctx.process_frames(h_clouds, h_R, device_input=False, nframes=Fe, results=res_e)

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
        ms = max(ms, wall * 0.999) if ms < wall * 0.5 else ms     # the library syncs its own stream inside fn
        barrier()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(3, args.warmup)):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ms = timed(step_device, args.steps)
    launches = ctx.launch_count() - l0
    stage_full = ctx.stage_ms()
    clocks = sampler.stop() if rank == 0 else None
    value = F * world * args.steps / (ms / 1e3)

    # the dominant streaming kernel alone (K1K2 = ingest + levelling + masking + organised normals), CUDA events on its stream
    nb = args.max_batch
    k_ms = []
    for i in range(3 + 10):
        ctx.run_stage_device(d_clouds.data_ptr() + (i % max(1, F // nb)) * nb * N * 16, d_R.data_ptr(), nb, 0)
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
    ctx.set_trace(True)
    ctx.process_frames(d_clouds, d_R, device_input=True, nframes=nb, results=results[:nb])
    trace = ctx.trace_ms()
    ctx.set_trace(False)
    chunk_ms = sum(t for _, t in trace)
    # the voxel stage's streaming kernel: reads every levelled point (16 B), writes (key, pixel) = 8 B per cropped point
    tdict = dict(trace)
    n_crop = int(results[:nb]["n_crop"].sum())
    vk_bytes = 16 * N * nb + 8 * n_crop
    vk_ms = tdict.get("k_voxel_key", 0.0)
    roof_voxel = None
    if vk_ms > 0:
        roof_voxel = {"bound": "hbm", "kernel": "k_voxel_key", "achieved": vk_bytes / (vk_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                      "frac": vk_bytes / (vk_ms * 1e-3) / 1e9 / peak, "ms_per_launch": vk_ms, "algorithmic_bytes_per_launch": vk_bytes,
                      "note": "timed inside a traced chunk (one CUDA event before and after the launch on the library's stream)"}
    kernels = [{"kernel": k, "ms": round(t, 4), "share": round(t / max(chunk_ms, 1e-9), 4)} for k, t in sorted(trace, key=lambda kv: -kv[1])[:8]]

    for _ in range(3):
        step_e2e()
    ms_e = timed(step_e2e, args.steps)
    e2e_val = Fe * world * args.steps / (ms_e / 1e3)

    # the same frames entering as what the sensor delivers (uint16 depth in mm + registered colour, 6 B / pixel instead of 16):
    # sp_process_depth_frames back-projects on the device (K2G::updateCloud), everything else is identical
    import importlib
    synth = importlib.import_module("stair_perception_b200.synth")
    dframes = [synth.make_depth_frame(i) for i in range(N_DISTINCT)]
    h_depth = torch.empty((Fe, H, W), dtype=torch.uint16).pin_memory()
    h_bgrx = torch.empty((Fe, H, W, 4), dtype=torch.uint8).pin_memory()
    for i in range(Fe):
        h_depth[i].copy_(torch.from_numpy(dframes[i % N_DISTINCT][0]))
        h_bgrx[i].copy_(torch.from_numpy(dframes[i % N_DISTINCT][1]))
    res_d = np.zeros(Fe, dtype=sp.FRAME_DTYPE)

    def step_depth():
        ctx.process_depth_frames(h_depth, synth.INTRINSICS, bgrx=h_bgrx, mirror=True, R=h_R, nframes=Fe, results=res_d)

    for _ in range(3):
        step_depth()
    ms_d = timed(step_depth, args.steps)
    depth_val = Fe * world * args.steps / (ms_d / 1e3)
    assert int(res_d["n_planes"].min()) >= 1 and int(res_d["status"].max()) == 0

    # sanity: the batch really produced planes
    assert int(results["n_planes"].min()) >= 1 and int(results["status"].max()) == 0, "unexpected empty results"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(sp, recs[:max(8, min(N_DISTINCT, host_cores()))], Rs[:max(8, min(N_DISTINCT, host_cores()))])

    if rank == 0:
        line = {
            "metric": "frames/sec of 512x424 stair clouds through segmentation; % HBM roofline",
            "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config 5: batch of synthetic Kinect-V2 4-step stair frames, 512x424 (217088 pts), levelling + full segmentation front-end",
                       "frames_per_gpu_per_step": F, "distinct_frames": N_DISTINCT, "max_batch": args.max_batch,
                       "lanes": 1 if os.environ.get("SP_NO_TWIN", "") not in ("", "0") else 2,
                       "l2": "inputs larger than L2 (%.1f GB resident per GPU)" % (F * N * 16 / 1e9), "parallelism": "frames sharded across %d GPU(s)" % world,
                       "host_numa_binding": numa_node},
            "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": int(Fe * (N * 16 + 36)),
                    "d2h_bytes_per_step": int(Fe * sp.FRAME_DTYPE.itemsize), "frames_per_gpu_per_step": Fe, "ms_per_step": ms_e / args.steps},
            "e2e_depth_input": {"value": depth_val, "unit": "frames/s", "h2d_bytes_per_step": int(Fe * (N * 6 + 36)),
                                "d2h_bytes_per_step": int(Fe * sp.FRAME_DTYPE.itemsize), "frames_per_gpu_per_step": Fe, "ms_per_step": ms_d / args.steps,
                                "note": "same scenes entering as uint16 depth + registered colour (6 B/pixel) through sp_process_depth_frames"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_ingest_normals", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic("k_ingest_normals", nb), "peak_source": peak_src, "ms_per_launch": k1, "frames_per_launch": nb,
                         "algorithmic_bytes_per_launch": K1K2_BYTES_PER_POINT * N * nb},
            "roofline_voxel_key": roof_voxel,
            "kernels_of_a_chunk": {"frames": nb, "ms": round(chunk_ms, 4), "top": kernels,
                                   "note": "k_cl_union (radius-graph union-find) and k_ransac are issue / latency bound on a few MB per chunk: no HBM roofline applies; "
                                           "k_ingest_normals is the largest kernel that streams the whole input"},
            "stage_ms_last_chunk": stage_full, "voxel_stage_ms": float(np.mean(vox_ms)),
            "cpu_baseline": cpu,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
