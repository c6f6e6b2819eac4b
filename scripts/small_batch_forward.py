#!/usr/bin/env python3
"""Forward-pass device time at small batches (YOLOv3 batch 1 and 4, YOLOv3-tiny and YOLOv2 batch 1) and how many layers the
planner split along K; run it again with B200_NO_SPLITK=1 for the A/B.  Usage: small_batch_forward.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tensorflow_b200 import synth, darknet as dn
import numpy as np
work = "/tmp/b200_bench"
for model, batch in (("yolov3", 1), ("yolov3", 4), ("yolov3-tiny", 1), ("yolov2", 1)):
    cfg = synth.make_cfg(model, work, batch=batch, width=416, height=416)
    wpath = os.path.join(work, f"{model}_seed0_damped.weights")
    if not os.path.exists(wpath): synth.write_weights(cfg, wpath, seed=0, damp_heads=True)
    fd = os.dup(2); dv = os.open(os.devnull, os.O_WRONLY); os.dup2(dv, 2)
    net = dn.Network(cfg, wpath, precision=dn.PREC_BF16)
    os.dup2(fd, 2)
    x = synth.make_images(batch, 3, 416, 416, 1002)
    net.set_head_sync(0); net.predict(x); net.predict(x)
    t = [net.profile_forward(100)[0] for _ in range(3)]
    ns = sum(1 for i in range(net.n) if "splitK" in dn.lib.b200_layer_plan(net.ptr, i).decode())
    print("%s batch %d: forward %.4f ms (split-K layers: %d)" % (model, batch, min(t), ns))
    net.close()
