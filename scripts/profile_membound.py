#!/usr/bin/env python3
"""One pass over every bandwidth- or latency-bound kernel of the path (maxpool, reorg, route copy, upsample, shortcut,
yolo/region/detection forward, decode, NMS, collect, letterbox) so that an `ncu --set full -k regex:...` capture of this
script holds one row per kernel and shape.  Usage (one GPU, under ncu):
  ncu --set full --clock-control none -k regex:'maxpool|reorg|forward_kernel|decode_|nms_|collect|class_count|upsample|shortcut|copy_channels|letterbox|zero_suppressed'
      -o gpurun_out/r2_membound python scripts/profile_membound.py
Nothing printed here is a bench value."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tensorflow_b200 import synth, darknet as dn  # noqa: E402

WORK = "/tmp/b200_bench"


def open_net(model, batch, size, fuse=None, damp=True):
    cfg = synth.make_cfg(model, WORK, batch=batch, width=size, height=size)
    wpath = os.path.join(WORK, f"{model}_seed0_{'damped' if damp else 'plain'}.weights")
    if not os.path.exists(wpath):
        synth.write_weights(cfg, wpath, seed=0, damp_heads=damp)
    fd = os.dup(2); dv = os.open(os.devnull, os.O_WRONLY); os.dup2(dv, 2)
    try:
        return dn.Network(cfg, wpath, precision=dn.PREC_BF16, fuse=fuse)
    finally:
        os.dup2(fd, 2); os.close(fd); os.close(dv)


def run(model, batch, size, thresh, nms, fuse=None, damp=True, env=None):
    for k in (env or {}):
        os.environ[k] = env[k]
    try:
        net = open_net(model, batch, size, fuse, damp)
    finally:
        for k in (env or {}):
            os.environ.pop(k, None)
    x = synth.make_images(batch, 3, size, size, 1002)
    net.predict(x)                                           # head sync on: yolo/region/detection forward kernels run
    w = h = 1 if model == "yolov1" else size
    rec, counts = net.detect_batch(x, w, h, thresh, nms)
    print(f"{model} {size} b{batch} fuse={fuse} damp={damp}: mean candidates {float(np.mean(counts)):.1f}, records {len(rec)}", flush=True)
    net.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["v2", "tiny", "v3", "v3plain", "v3-608", "v1", "stress", "letterbox"]
    if "v2" in which:
        run("yolov2", 64, 416, .5, .45)                      # maxpool x5 (one fused), reorg, route copy, region_forward, decode, nms
    if "tiny" in which:
        run("yolov3-tiny", 64, 416, .5, .45)                 # maxpool incl. size 2 / stride 1, upsample, route
    if "v3" in which:
        run("yolov3", 64, 416, .5, .45)                      # yolo_forward x3, decode, nms at the headline config
    if "v3plain" in which:                                   # unfused plan: standalone shortcut / upsample / route-copy kernels
        run("yolov3", 64, 416, .5, .45, fuse=False, env={"B200_NO_ZERO_COPY_ROUTE": "1", "B200_NO_UPSAMPLE_FUSION": "1"})
    if "v3-608" in which:
        run("yolov3", 32, 608, .5, .45)                      # NMS general kernel at 608
    if "v1" in which:
        run("yolov1", 64, 448, .2, .4)                       # detection_forward, decode (detection)
    if "stress" in which:
        run("yolov3", 8, 416, .5, .45, damp=False)           # undamped heads: ~6.7k candidates per image
    if "letterbox" in which:
        net = open_net("yolov3", 64, 416)
        rng = np.random.default_rng(3)
        ims = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(64)]
        net.letterbox_batch_u8(ims)
        rec, counts = net.detect_batch(None, 0, 0, .5, .45)
        print("letterbox: records", len(rec), flush=True)
        net.close()
