"""Image sharding across the GPUs of one box: weights replicated by ONE broadcast of the
parameter arena, images split into contiguous per-rank blocks, detections gathered once at the end.
There is no collective on the per-layer path.  Works over NCCL (GPU tensors) and gloo (CPU tensors; the
world_size-2 tests in tests/test_shard_gloo.py)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """contiguous block [lo, hi) of `total` images owned by `rank`; blocks differ by at most one image"""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_arena(arena, src=0):
    """replicate the folded/repacked parameter arena (a flat uint8 tensor on this rank's device)"""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(arena, src=src)
    return arena


def gather_records(records, image_offset, device="cpu"):
    """variable-length gather of per-rank detection records (structured numpy array with an 'image' field that is
    LOCAL to the rank).  Returns, on every rank, the concatenation in rank order with global image ids."""
    rec = records.copy()
    if len(rec):
        rec["image"] += image_offset
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return rec
    world = dist.get_world_size()
    itemsize = rec.dtype.itemsize
    n = torch.tensor([len(rec)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    buf = torch.zeros(cap * itemsize, dtype=torch.uint8, device=device)
    if len(rec):
        raw = torch.from_numpy(np.frombuffer(rec.tobytes(), dtype=np.uint8).copy())
        buf[: raw.numel()] = raw.to(device)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    parts = [np.frombuffer(b.cpu().numpy().tobytes()[: c * itemsize], dtype=rec.dtype) for b, c in zip(bufs, counts)]
    return np.concatenate(parts) if parts else rec
