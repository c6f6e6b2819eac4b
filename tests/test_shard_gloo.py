"""world_size-2 gloo test of the multi-GPU host logic (image sharding, arena broadcast, detection gather)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from yolo_tensorflow_b200.shard import broadcast_arena, gather_records, shard_range  # noqa: E402

DT = np.dtype([("image", "<i4"), ("cls", "<i4"), ("box_id", "<i4"), ("prob", "<f4"), ("objectness", "<f4"),
               ("x", "<f4"), ("y", "<f4"), ("w", "<f4"), ("h", "<f4")])


def test_shard_range_covers_everything():
    for total in (1, 7, 64, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arena = torch.arange(1000, dtype=torch.uint8) if rank == 0 else torch.zeros(1000, dtype=torch.uint8)
    broadcast_arena(arena)
    ok_arena = bool((arena == torch.arange(1000, dtype=torch.uint8)).all())
    lo, hi = shard_range(5, rank, world)
    rec = np.zeros(3 if rank == 0 else 0, dtype=DT)          # ragged: rank 1 has nothing
    if rank == 0:
        rec["image"] = [0, 0, 2]; rec["cls"] = [1, 2, 3]; rec["prob"] = [.9, .8, .7]
    allrec = gather_records(rec, lo)
    q.put((rank, ok_arena, (lo, hi), allrec.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_broadcast_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ok0, span0, raw0), (r1, ok1, span1, raw1) = res
    assert ok0 and ok1
    assert span0 == (0, 3) and span1 == (3, 5)
    a0, a1 = np.frombuffer(raw0, dtype=DT), np.frombuffer(raw1, dtype=DT)
    assert np.array_equal(a0, a1) and len(a0) == 3
    assert a0["image"].tolist() == [0, 0, 2] and a0["cls"].tolist() == [1, 2, 3]
