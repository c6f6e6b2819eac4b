#!/usr/bin/env python3
"""per-layer device time / TFLOP/s table for a model (CUDA events between layers). Usage: layer_times.py [model] [batch] [size]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tensorflow_b200 import synth, darknet as dn
model = sys.argv[1] if len(sys.argv) > 1 else "yolov3"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
size = int(sys.argv[3]) if len(sys.argv) > 3 else 416
work = "/tmp/b200_bench"
cfg = synth.make_cfg(model, work, batch=batch, width=size, height=size)
wpath = os.path.join(work, f"{model}_seed0_damped.weights")
if not os.path.exists(wpath): synth.write_weights(cfg, wpath, seed=0, damp_heads=True)
fd = os.dup(2); dv = os.open(os.devnull, os.O_WRONLY); os.dup2(dv, 2)
net = dn.Network(cfg, wpath, precision=dn.PREC_BF16)
os.dup2(fd, 2)
x = synth.make_images(batch, 3, size, size, 1002)
net.set_head_sync(0)
net.predict(x); net.predict(x)
ms = net.profile_layers(5)
tot = 0; totf = 0
print("%3s %-13s %-10s %8s %9s %8s  %s" % ("i", "type", "kernel", "ms", "GFLOP", "TFLOP/s", "plan"))
for i, L in enumerate(net.layers):
    fl = 0
    if L["type_name"] == "CONVOLUTIONAL":
        fl = 2.0 * L["n"] * L["size"] ** 2 * L["c"] * L["out_h"] * L["out_w"] * batch
    tot += ms[i]; totf += fl
    print("%3d %-13s %-10s %8.4f %9.2f %8.1f  %s" % (i, L["type_name"], net.kernel(i), ms[i], fl / 1e9, fl / (ms[i] * 1e-3) / 1e12 if ms[i] > 0 and fl else 0,
          dn.lib.b200_layer_plan(net.ptr, i).decode().replace("conv_tc ", "")))
print("total %.3f ms, %.1f TFLOP/s overall" % (tot, totf / (tot * 1e-3) / 1e12))
th, nms = (0.2, 0.4) if model == "yolov1" else (0.5, 0.45)
tail = net.profile_tail(size, size, th, nms, 5)
print("tail: decode %.4f ms, nms %.4f ms, collect %.4f ms" % tuple(tail))
