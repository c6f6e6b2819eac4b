"""Two ranks on two GPUs through the library's own NCCL calls (csrc/dev/comm.cu): rank 0 alone loads the weights,
b200_comm_broadcast_weights replicates the parameter arena, every rank detects on ITS images and b200_comm_set_gather delivers
all records to rank 0 with global image numbers.  The gathered result must equal what one process computes for all images.
Needs two CUDA devices (skipped otherwise); the world_size-2 gloo tests in tests/test_shard_gloo.py cover the host logic."""
import ctypes
import os
import sys
import time

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from conftest import model_files  # noqa: E402

pytestmark = pytest.mark.gpu
PER_RANK, SIZE, MODEL = 3, 160, "yolov3-tiny"


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _quiet_net(dn, cfg, wpath):
    fd = os.dup(2); dv = os.open(os.devnull, os.O_WRONLY); os.dup2(dv, 2)
    try:
        return dn.Network(cfg, wpath, precision=dn.PREC_BF16)
    finally:
        os.dup2(fd, 2); os.close(fd); os.close(dv)


def _rank_main(rank, world, workdir, outdir):
    from yolo_tensorflow_b200 import synth, darknet as dn
    dn.set_gpu(rank)
    cfg, wpath = model_files(MODEL, PER_RANK, SIZE, workdir)
    net = _quiet_net(dn, cfg, wpath if rank == 0 else None)          # only rank 0 reads the .weights file
    ident = (ctypes.c_ubyte * 128)()
    id_path = os.path.join(outdir, "nccl_id")
    if rank == 0:
        dn.lib.b200_comm_unique_id(ident, 128)
        with open(id_path + ".tmp", "wb") as f:
            f.write(bytes(ident))
        os.replace(id_path + ".tmp", id_path)
    else:
        t0 = time.time()
        while not os.path.exists(id_path):
            assert time.time() - t0 < 120
            time.sleep(.05)
        ident = (ctypes.c_ubyte * 128).from_buffer_copy(open(id_path, "rb").read())
    dn.lib.b200_comm_init(net.ptr, ident, rank, world)
    dn.lib.b200_comm_broadcast_weights(net.ptr, 0)
    dn.lib.b200_comm_set_gather(net.ptr, 0, rank * PER_RANK, 20000)
    x = synth.make_images(PER_RANK, 3, SIZE, SIZE, 500 + rank)
    for rep in range(3):                                             # several batches: the slots are reused
        rec, counts = net.detect_batch(x, SIZE, SIZE, .3, .45)
    np.save(os.path.join(outdir, f"rec{rank}.npy"), rec)
    np.save(os.path.join(outdir, f"counts{rank}.npy"), counts)
    dn.lib.b200_comm_destroy(net.ptr)
    net.close()


@pytest.mark.skipif(_device_count() < 2, reason="needs two CUDA devices")
def test_two_ranks_broadcast_weights_and_gather_records(workdir, tmp_path):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_rank_main, args=(r, 2, workdir, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    from yolo_tensorflow_b200 import synth, darknet as dn
    cfg, wpath = model_files(MODEL, PER_RANK, SIZE, workdir)
    net = _quiet_net(dn, cfg, wpath)
    want = []
    for rank in range(2):
        rec, _ = net.detect_batch(synth.make_images(PER_RANK, 3, SIZE, SIZE, 500 + rank), SIZE, SIZE, .3, .45)
        rec = rec.copy(); rec["image"] += rank * PER_RANK
        want.append(rec)
    net.close()
    order = lambda r: r[np.lexsort((r["cls"], r["box_id"], r["image"]))]
    gathered = np.load(tmp_path / "rec0.npy")
    local1 = np.load(tmp_path / "rec1.npy")
    assert len(want[1]) > 0
    assert order(gathered).tobytes() == order(np.concatenate(want)).tobytes()       # rank 0 holds everything, global image ids
    assert order(local1).tobytes() == order(want[1]).tobytes()                      # rank 1 still sees its own, with global image numbers
