#!/usr/bin/env python3
"""device time of every single forward pass after an idle rest (how fast the power management reacts). Usage: pass_times.py [flow 0/1] [passes] [rest s]"""
import os, sys, time
os.environ.setdefault("B200_FLOW", "1")
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_tensorflow_b200 import synth, darknet as dn
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rest = float(sys.argv[2]) if len(sys.argv) > 2 else 3.
work = "/tmp/b200_bench"
cfg = synth.make_cfg("yolov3", work, batch=64, width=416, height=416)
wpath = os.path.join(work, "yolov3_seed0_damped.weights")
if not os.path.exists(wpath): synth.write_weights(cfg, wpath, seed=0, damp_heads=True)
fd = os.dup(2); dv = os.open(os.devnull, os.O_WRONLY); os.dup2(dv, 2)
net = dn.Network(cfg, wpath, precision=dn.PREC_BF16)
os.dup2(fd, 2)
x = synth.make_images(64, 3, 416, 416, 1002)
net.set_head_sync(0)
net.predict(x); net.predict(x)
stream = torch.cuda.ExternalStream(net.stream_ptr())
eng = dn.lib.b200_engine_of(net.ptr)
for on in (0, 1, 0, 1):
    net.set_flow(on)
    time.sleep(rest)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(passes + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for i in range(passes):
            dn.lib.b200_engine_forward_resident(eng, net.ptr)
            ev[i + 1].record(stream)
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(passes)]
    print("flow %d:" % on, " ".join("%.2f" % v for v in t))
