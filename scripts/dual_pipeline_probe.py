#!/usr/bin/env python3
"""Experiment: do two half-batch pipelines on two streams fill each other's inter-layer bubbles?  K networks of batch 64/K, each
with its own stream, driven by K host threads through the resident serving loop; total images/s against one batch-64 network.
Usage: dual_pipeline_probe.py [steps]"""
import ctypes, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tensorflow_b200 import synth, darknet as dn
import torch

WORK = "/tmp/b200_bench"
RES = ctypes.c_void_p(1)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30


def open_net(batch):
    cfg = synth.make_cfg("yolov3", WORK, batch=batch, width=416, height=416)
    wpath = os.path.join(WORK, "yolov3_seed0_damped.weights")
    if not os.path.exists(wpath):
        synth.write_weights(cfg, wpath, seed=0, damp_heads=True)
    fd = os.dup(2); dv = os.open(os.devnull, os.O_WRONLY); os.dup2(dv, 2)
    try:
        net = dn.Network(cfg, wpath, precision=dn.PREC_BF16)
    finally:
        os.dup2(fd, 2); os.close(fd); os.close(dv)
    net.set_head_sync(0)
    x = synth.make_images(batch, 3, 416, 416, 1002)
    net.predict(x)                      # leaves the batch resident in the device input buffer
    return net


def loop(net, n, out, counts, barrier, result, idx):
    dn.lib.b200_submit_batch(net.ptr, RES)
    for _ in range(3):
        dn.lib.b200_detect_submitted(net.ptr, RES, 416, 416, .5, .45, 1, out, len(out), counts)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(n):
        dn.lib.b200_detect_submitted(net.ptr, RES, 416, 416, .5, .45, 1, out, len(out), counts)
    torch.cuda.synchronize()
    result[idx] = time.perf_counter() - t0
    dn.lib.b200_detect_submitted(net.ptr, None, 416, 416, .5, .45, 1, out, len(out), counts)


for k in (1, 2, 4):
    b = 64 // k
    nets = [open_net(b) for _ in range(k)]
    outs = [(dn.B200_DET * (1 << 19))() for _ in range(k)]
    cnts = [(ctypes.c_int * b)() for _ in range(k)]
    barrier = threading.Barrier(k)
    result = [0.] * k
    th = [threading.Thread(target=loop, args=(nets[i], steps, outs[i], cnts[i], barrier, result, i)) for i in range(k)]
    for t in th: t.start()
    for t in th: t.join()
    dt = max(result)
    print(f"{k} pipeline(s) of batch {b}: {64 * steps / dt:9.1f} images/s  ({1e3 * dt / steps:.3f} ms per 64 images)", flush=True)
    for n in nets: n.close()
