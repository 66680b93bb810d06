"""Frame-wise sharding of a batch across the GPUs of one box, and the gather of the per-frame result records.

Frames are independent (the reference builds a fresh StairDetection per frame, /root/reference/test/processor.cpp:156),
so the batch is cut into contiguous blocks, one per rank, every rank runs WHOLE frames on its own GPU and the only
exchange is one all-gather of the fixed-size `sp_frame_result` records (a few KB per frame).  No point, voxel or
hypothesis ever crosses a GPU.  Works with any torch.distributed backend: `nccl` on the GPU box (records staged through
a device tensor), `gloo` in the CPU tests.
"""
import numpy as np


def shard_bounds(nframes, rank, world):
    """Contiguous block [lo, hi) of rank `rank`; the first `nframes % world` ranks get one extra frame."""
    if world < 1 or not (0 <= rank < world) or nframes < 0:
        raise ValueError("bad shard arguments")
    q, r = divmod(nframes, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_sizes(nframes, world):
    return [shard_bounds(nframes, r, world)[1] - shard_bounds(nframes, r, world)[0] for r in range(world)]


def gather_results(local, nframes, group=None, device=None):
    """All-gather of the per-rank result records into the full batch order.

    local   : numpy structured array (FRAME_DTYPE) of this rank's block
    nframes : total frames of the batch
    returns : numpy structured array of `nframes` records, identical on every rank
    """
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if len(local) != nframes:
            raise ValueError("single-process gather needs the whole batch")
        return local.copy()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(nframes, world)
    if len(local) != sizes[rank]:
        raise ValueError("rank %d holds %d records, its shard has %d" % (rank, len(local), sizes[rank]))
    item = local.dtype.itemsize
    cap = max(sizes) * item                                   # equal-sized slots: one fixed-size collective
    mine = torch.zeros(cap, dtype=torch.uint8)
    mine[: len(local) * item] = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).reshape(-1))
    if device is not None:
        mine = mine.to(device, non_blocking=True)
    out = torch.empty(world * cap, dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    flat = out.cpu().numpy()
    parts = [flat[r * cap: r * cap + sizes[r] * item].view(local.dtype) for r in range(world)]
    return np.concatenate(parts)
