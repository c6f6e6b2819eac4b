#!/usr/bin/env python3
"""Developer probe (not a test): runs a model through the engine on the GPU and prints per-layer error against
the oracle (reference .so when present, else the numpy port).  Usage: dev_check.py MODEL [fp32|bf16] [batch] [size]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tensorflow_b200 import synth, darknet as dn
from oracle import np_darknet as P

model = sys.argv[1] if len(sys.argv) > 1 else "yolov3-tiny"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
size = int(sys.argv[4]) if len(sys.argv) > 4 else None
work = "/tmp/b200_work"
cfg = synth.make_cfg(model, work, batch=batch, width=size, height=size)
wpath = os.path.join(work, os.path.basename(cfg).replace(".cfg", ".weights"))
if not os.path.exists(wpath):
    synth.write_weights(cfg, wpath, seed=0, damp_heads=True)
port = P.Net(cfg, wpath)
x = synth.make_images(batch, 3, port.h, port.w, 1000)
t = time.time(); outs = port.forward(x); print("oracle forward %.2fs" % (time.time() - t))
net = dn.Network(cfg, wpath, precision=dn.PREC_FP32 if prec == "fp32" else dn.PREC_BF16, fuse=False)
t = time.time(); net.predict(x); print("engine first predict %.3fs" % (time.time() - t))
t = time.time(); net.predict(x); print("engine second predict %.3fs" % (time.time() - t))
SUMMARY = os.environ.get("SUMMARY", "0") == "1"
LIMIT = 2e-2 if prec == "bf16" else 1e-4
print("== end-to-end (free-running) error per layer")
worst = {}
for i, o in enumerate(outs):
    a = net.layer_output(i).reshape(-1); r = o.reshape(-1)
    scale = np.abs(r).max() + 1e-30
    e = np.abs(a - r).max() / scale
    worst[net.kernel(i)] = max(worst.get(net.kernel(i), 0), e)
    if not SUMMARY or e > 5 * LIMIT or not np.isfinite(e):
        print("%3d %-14s %-12s max|d|/max|r| %.3e  rms rel %.3e" % (i, port.layers[i].type, net.kernel(i), e,
              np.sqrt(((a - r) ** 2).mean()) / (np.sqrt((r ** 2).mean()) + 1e-30)))
print("   worst free-running by kernel:", {k: "%.2e" % v for k, v in worst.items()})
print("== teacher-forced error per layer")
worst = {}
for i, L in enumerate(port.layers):
    if i == 0: continue
    srcs = [i - 1]
    if L.type == "route": srcs = L.src
    if L.type == "shortcut": srcs = [i - 1, L.src]
    for j in srcs: net.set_layer_output(j, outs[j].reshape(batch, -1))
    net.run_layers(i, i + 1)
    a = net.layer_output(i).reshape(-1); r = outs[i].reshape(-1)
    scale = np.abs(r).max() + 1e-30
    e = np.abs(a - r).max() / scale
    worst[net.kernel(i)] = max(worst.get(net.kernel(i), 0), e)
    if not SUMMARY or e > LIMIT or not np.isfinite(e):
        print("%3d %-14s %-12s max|d|/max|r| %.3e" % (i, L.type, net.kernel(i), e))
print("   worst teacher-forced by kernel:", {k: "%.2e" % v for k, v in worst.items()})
th, nms = (0.2, 0.4) if model == "yolov1" else (0.5, 0.45)
w_, h_ = (1, 1) if model == "yolov1" else (port.w, port.h)
net.predict(x)
for b in range(min(batch, 2)):
    dets, n = net.boxes(b, w_, h_, th)
    eb, eo, ep = dn.dets_to_arrays(dets, n, net.layers[-1]["classes"])
    pb, po, pp, pid = P.get_network_boxes(port, outs, b, w_, h_, th)
    print("image", b, "engine dets", n, "oracle dets", len(pb))
    if n == len(pb) and n:
        print("   box err", np.abs(eb - pb).max(), "prob err", np.abs(ep - pp).max())
    dn.free_detections(dets, n)
rec, counts = net.detect_batch(x, w_, h_, th, nms)
print("detect_batch records", len(rec), "counts", counts[:4])
pb, po, pp, pid = P.get_network_boxes(port, outs, 0, w_, h_, th)
after = P.do_nms_sort(pb, po, pp, nms)
print("oracle kept pairs image0:", int((after > 0).sum()), "engine kept image0:", int((rec["image"] == 0).sum()) if len(rec) else 0)
print("launches", dn.lib.b200_launch_count())
