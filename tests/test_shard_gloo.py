"""Multi-GPU host logic on CPU: world_size-2 `gloo` processes shard a batch frame-wise, each produces result records
for its block, and the all-gather reassembles the batch in order (the only inter-GPU exchange of this path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_bounds(sp):
    import importlib
    shard = importlib.import_module("stair_perception_b200.shard")
    for n in (0, 1, 7, 8, 8192, 1000):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard.shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == shard.shard_sizes(n, world)
    with pytest.raises(ValueError):
        shard.shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nframes, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import importlib
    import torch.distributed as dist
    from __graft_entry__ import load_package
    sp = load_package()
    shard = importlib.import_module("stair_perception_b200.shard")
    synth = importlib.import_module("stair_perception_b200.synth")
    from oracle import orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_bounds(nframes, rank, world)
    local = np.zeros(hi - lo, dtype=sp.FRAME_DTYPE)
    for i in range(lo, hi):                                   # stand-in for the CUDA path: the oracle fills the records
        rec, R = synth.make_frame(100 + i)
        o = orc.Oracle(sp.DEFAULT_PARAMS)
        o.process(rec, 512, 424, R)
        for k, v in o.counts().items():
            local[i - lo][k] = v
        pl = o.tap("planes")
        for j in range(len(pl)):
            for fld in pl.dtype.names:
                local[i - lo]["planes"][j][fld] = pl[j][fld]
    full = shard.gather_results(local, nframes)
    np.save(os.path.join(outdir, "rank%d.npy" % rank), full.view(np.uint8))
    # the compact exchange of the GPU path (sp_process_frames_compact + shard.gather_compact), here on CPU tensors over gloo:
    # header + n_planes records per frame, a buffer sized like sp_compact_bytes_max, then expanded again
    import torch
    mine = sp.compact_records(local)
    buf = torch.zeros(int(sp.lib().sp_compact_bytes_max(max(shard.shard_sizes(nframes, world)))), dtype=torch.uint8)
    buf[:mine.size] = torch.from_numpy(mine)
    packed, sizes = shard.gather_compact(buf, int(mine.size))
    assert len(sizes) == world and sizes[rank] == mine.size and packed.size == sum(sizes)
    np.save(os.path.join(outdir, "compact%d.npy" % rank), sp.expand_compact(packed, nframes).view(np.uint8))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(sp, orc, synth, tmp_path):
    nframes, world = 5, 2                                     # uneven split: 3 + 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nframes, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(os.path.join(str(tmp_path), "rank%d.npy" % r)).view(sp.FRAME_DTYPE) for r in range(world)]
    assert np.array_equal(got[0].view(np.uint8), got[1].view(np.uint8)) and len(got[0]) == nframes
    for r in range(world):                                    # the compact exchange reassembles the very same records
        assert np.array_equal(np.load(os.path.join(str(tmp_path), "compact%d.npy" % r)), got[0].view(np.uint8).reshape(-1))
    for i in range(nframes):
        rec, R = synth.make_frame(100 + i)
        o = orc.Oracle(sp.DEFAULT_PARAMS)
        o.process(rec, 512, 424, R)
        for k, v in o.counts().items():
            assert int(got[0][i][k]) == v, (i, k)
        pl = o.tap("planes")
        for j in range(len(pl)):
            assert np.array_equal(got[0][i]["planes"][j]["coef"], pl[j]["coef"])
