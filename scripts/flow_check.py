#!/usr/bin/env python3
"""flow kernel vs one launch per layer: bit-identical layer outputs, and the forward-pass time of both.
Usage: flow_check.py [model] [batch] [size] [passes]"""
import os, sys, time
os.environ.setdefault("B200_FLOW", "1")
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tensorflow_b200 import synth, darknet as dn
model = sys.argv[1] if len(sys.argv) > 1 else "yolov3"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
size = int(sys.argv[3]) if len(sys.argv) > 3 else 416
passes = int(sys.argv[4]) if len(sys.argv) > 4 else 3
work = "/tmp/b200_bench"
cfg = synth.make_cfg(model, work, batch=batch, width=size, height=size)
wpath = os.path.join(work, f"{model}_seed0_damped.weights")
if not os.path.exists(wpath): synth.write_weights(cfg, wpath, seed=0, damp_heads=True)
fd = os.dup(2); dv = os.open(os.devnull, os.O_WRONLY); os.dup2(dv, 2)
net = dn.Network(cfg, wpath, precision=dn.PREC_BF16)
os.dup2(fd, 2)
for f in net.flows(): print("flow", f)
x = synth.make_images(batch, 3, size, size, 1002)
net.set_head_sync(0)
NOT_MAT = ("conv_tc+shortcut", "conv_tc(block)", "conv_tc+upsample", "conv_stem+maxpool")
layers = [i for i in range(net.n) if net.kernel(i) not in NOT_MAT and net.layers[i]["type_name"] in ("CONVOLUTIONAL", "SHORTCUT", "YOLO", "REGION")]
net.set_flow(0); net.predict(x)
ref = {i: net.layer_output(i).copy() for i in layers}
bad = 0
for p in range(passes):
    net.set_flow(1); net.predict(x)
    for i in layers:
        a = net.layer_output(i)
        if not np.array_equal(a, ref[i]):
            d = np.abs(a - ref[i]); bad += 1
            print("pass %d layer %d (%s) differs: max %.4g, %d elements, nan %d" % (p, i, net.kernel(i), np.nanmax(d), int((d > 0).sum()), int(np.isnan(a).sum())))
            break
print("bit-identical" if not bad else "MISMATCH", "over", len(layers), "layers,", passes, "passes")
net.flow_stats()
for on in (0, 1, 0, 1):
    net.set_flow(on)
    t, first = net.profile_forward(20)
    st = net.flow_stats()
    print("flow %d: forward %.4f ms (first layer %.4f); per pass: producers blocked %.1f us-pair-CTA, residual loaders %.1f, blocking waits %d" %
          (on, t, first, st[0] / 21e3, st[1] / 21e3, st[2] // 21))
# the fused detection path under both modes
recs = {}
for on in (0, 1, 0, 1):
    net.set_flow(on)
    r, c = net.detect_batch(x, size, size, .5, .45)
    recs.setdefault(on, []).append((r, c))
a, b = recs[0][0][0], recs[1][0][0]
print("records: flow 0 %d %d, flow 1 %d %d" % (len(recs[0][0][0]), len(recs[0][1][0]), len(recs[1][0][0]), len(recs[1][1][0])))
print("same bytes 0/0 %s 1/1 %s 0/1 %s" % (recs[0][0][0].tobytes() == recs[0][1][0].tobytes(), recs[1][0][0].tobytes() == recs[1][1][0].tobytes(), a.tobytes() == b.tobytes()))
if len(a) == len(b) and a.tobytes() != b.tobytes():
    ka = np.lexsort((a["cls"], a["box_id"], a["image"])); kb = np.lexsort((b["cls"], b["box_id"], b["image"]))
    sa, sb = a[ka], b[kb]
    print("same after sorting:", sa.tobytes() == sb.tobytes())
    bad = [i for i in range(len(sa)) if sa[i].tobytes() != sb[i].tobytes()][:5]
    for i in bad: print(sa[i], sb[i])
