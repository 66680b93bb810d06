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


_slots = {}                                                      # the receive buffer of the last gather, reused while the size stays


def gather_compact(local, nbytes, group=None, host_out=None, parts=None, to_host=True):
    """All-gather of COMPACT records (sp_process_frames_compact: 64-byte header + n_planes plane records per frame) straight
    from device memory: no host bounce, about 3 KB per frame on the wire instead of the 14 400-byte fixed record.

    local    : torch uint8 tensor holding this rank's compact records in [:nbytes] (CUDA with nccl, CPU with gloo); its
               capacity must cover the largest rank's byte count (sp_compact_bytes_max(frames of the largest shard) does)
    returns  : (numpy uint8 array with the compact records of the WHOLE batch in batch order, [bytes per rank]); identical on
               every rank.  Shards are contiguous blocks in rank order, so concatenating the ranks' buffers is batch order.
    host_out : optional pinned torch uint8 CPU tensor to receive the result (avoids a pageable copy).
    parts    : optional dict; with CUDA tensors it receives the device-timed milliseconds of the three steps
               (sizes exchange, all-gather of the records, copy of the gathered batch to this rank's host).
    to_host  : False (CUDA tensors only) leaves the gathered batch on this rank's GPU and returns it as a torch tensor: with N
               ranks on one box, N copies of the same bytes to the same host share one PCIe / memory path (2.0 of the 2.3 ms of
               the 8-GPU exchange), so only the rank that consumes the records on the host should ask for them there."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if host_out is not None and local.is_cuda:
            host_out[:nbytes].copy_(local[:nbytes], non_blocking=True)
            torch.cuda.current_stream(local.device).synchronize()
            return host_out[:nbytes].numpy(), [int(nbytes)]
        return local[:nbytes].cpu().numpy(), [int(nbytes)]
    world = dist.get_world_size(group)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if parts is not None and local.is_cuda else None
    if ev:
        ev[0].record()
    sizes_t = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(sizes_t, torch.tensor([int(nbytes)], dtype=torch.int64, device=local.device), group=group)
    sizes = [int(v) for v in sizes_t.tolist()]
    if ev:
        ev[1].record()
    pad = (max(sizes) + 15) // 16 * 16                          # equal-sized slots: one fixed-size collective
    if local.numel() < pad:
        raise ValueError("compact buffer of %d bytes cannot hold the largest shard (%d bytes)" % (local.numel(), pad))
    key = (str(local.device), world * pad)
    slots = _slots.get(key)
    if slots is None:
        _slots.clear()
        slots = _slots[key] = torch.empty(world * pad, dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(slots, local[:pad], group=group)
    if all(sz == pad for sz in sizes):                          # records are multiples of 16 bytes: equal shards leave no gaps
        packed = slots
    else:
        packed = torch.cat([slots[r * pad: r * pad + sizes[r]] for r in range(world)])
    total = int(packed.numel())
    if ev:
        ev[2].record()
    if not to_host and packed.is_cuda:
        if ev:
            ev[3].record()
            ev[3].synchronize()
            parts.update(sizes_ms=ev[0].elapsed_time(ev[1]), all_gather_ms=ev[1].elapsed_time(ev[2]), to_host_ms=0.0)
        return packed[:total], sizes
    if host_out is not None and packed.is_cuda:
        host_out[:total].copy_(packed, non_blocking=True)
        if ev:
            ev[3].record()
        torch.cuda.current_stream(local.device).synchronize()
        if ev:
            parts.update(sizes_ms=ev[0].elapsed_time(ev[1]), all_gather_ms=ev[1].elapsed_time(ev[2]), to_host_ms=ev[2].elapsed_time(ev[3]))
        return host_out[:total].numpy(), sizes
    return packed.cpu().numpy(), sizes
