"""Multi-GPU exchange on real devices: one batch cut by shard.shard_bounds, every rank runs whole frames through
sp_process_frames_compact (records stay on the device), shard.gather_compact all-gathers the compact records over NCCL, and the
expanded result must equal, byte for byte, what ONE GPU returns for the whole batch through sp_process_frames.
With fewer than two GPUs the NCCL leg is skipped and the same cut is checked inside one process (compact buffers concatenated
on the host instead of gathered)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

FRAMES = (41, 42, 43, 44, 45)


def _batch(synth):
    frames = [synth.make_frame(f) for f in FRAMES]
    recs = np.stack([frames[i % len(frames)][0] for i in range(7)])
    Rs = np.stack([frames[i % len(frames)][1] for i in range(7)])
    recs[3]["rgba"] = 0                                        # an empty frame in the middle: a header-only compact record
    return recs, Rs


def test_compact_records_equal_fixed_records_and_survive_a_cut(sp, synth):
    import importlib
    import torch
    shard = importlib.import_module("stair_perception_b200.shard")
    recs, Rs = _batch(synth)
    ctx = sp.Context(device=0, width=512, height=424, max_batch=2)
    want = ctx.process_frames(recs, Rs).copy()
    assert int(want[3]["status"]) == 1 and int(want["n_planes"].max()) >= 8
    # whole batch, host input and device input
    out, nb = ctx.process_frames_compact(recs, Rs)
    assert nb == 64 * len(want) + 224 * int(want["n_planes"].sum())
    assert sp.expand_compact(out[:nb].cpu().numpy(), len(want)).tobytes() == want.tobytes()
    d_c = torch.from_numpy(recs.view(np.uint8).reshape(len(recs), -1)).cuda()
    d_R = torch.from_numpy(Rs).cuda()
    offs = torch.zeros(len(recs) + 1, dtype=torch.int64, device="cuda")
    out, nb2 = ctx.process_frames_compact(d_c, d_R, device_input=True, nframes=len(recs), offsets=offs)
    assert nb2 == nb and int(offs[-1]) == nb and int(offs[4]) - int(offs[3]) == 64
    assert sp.expand_compact(out[:nb].cpu().numpy(), len(want)).tobytes() == want.tobytes()
    # the taps of the last chunk still work when the records live in the batch array
    assert ctx.tap(0, "plane_idx").size == int(want[6]["n_plane_points"])
    # a cut in two and in three blocks, concatenated in rank order
    for world in (2, 3):
        parts = []
        for r in range(world):
            lo, hi = shard.shard_bounds(len(recs), r, world)
            o, n = ctx.process_frames_compact(d_c[lo:hi], d_R[lo:hi], device_input=True, nframes=hi - lo)
            parts.append(o[:n].cpu().numpy())
        assert sp.expand_compact(np.concatenate(parts), len(want)).tobytes() == want.tobytes()
    # a buffer that is too small is an error, not a silent truncation
    small = torch.empty(1000, dtype=torch.uint8, device="cuda")
    with pytest.raises(sp.SpError):
        ctx.process_frames_compact(d_c, d_R, device_input=True, nframes=len(recs), out=small)
    ctx.close()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import importlib
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    sp = load_package()
    shard = importlib.import_module("stair_perception_b200.shard")
    synth = importlib.import_module("stair_perception_b200.synth")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    recs, Rs = _batch(synth)
    lo, hi = shard.shard_bounds(len(recs), rank, world)
    ctx = sp.Context(device=rank, width=512, height=424, max_batch=2)
    cap = int(sp.lib().sp_compact_bytes_max(max(shard.shard_sizes(len(recs), world))))
    buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
    _, nb = ctx.process_frames_compact(recs[lo:hi], Rs[lo:hi], out=buf)
    packed, sizes = shard.gather_compact(buf, nb)
    # the same exchange with the gathered batch left on the GPU (what a rank that does not consume the records on its host asks for)
    on_dev, sizes2 = shard.gather_compact(buf, nb, to_host=False)
    assert on_dev.is_cuda and sizes2 == sizes and on_dev.cpu().numpy().tobytes() == packed.tobytes()
    np.save(os.path.join(outdir, "rank%d.npy" % rank), sp.expand_compact(packed, len(recs)).view(np.uint8))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_nccl_gather_equals_single_gpu(sp, synth, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2); the one-GPU cut test above covers the compact path")
    import torch.multiprocessing as mp
    recs, Rs = _batch(synth)
    ctx = sp.Context(device=0, width=512, height=424, max_batch=2)
    want = ctx.process_frames(recs, Rs).copy()
    ctx.close()
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        assert got.tobytes() == want.tobytes()
