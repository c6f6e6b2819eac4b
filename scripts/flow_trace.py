#!/usr/bin/env python3
"""Per-layer anatomy of a flow kernel launch from its device-clock trace (B200_FLOW_TRACE=1).
Usage: flow_trace.py [model] [batch] [size] [flow index]"""
import os, sys
os.environ["B200_FLOW_TRACE"] = "1"
os.environ.setdefault("B200_FLOW", "1")
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tensorflow_b200 import synth, darknet as dn
model = sys.argv[1] if len(sys.argv) > 1 else "yolov3"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
size = int(sys.argv[3]) if len(sys.argv) > 3 else 416
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
work = "/tmp/b200_bench"
cfg = synth.make_cfg(model, work, batch=batch, width=size, height=size)
wpath = os.path.join(work, f"{model}_seed0_damped.weights")
if not os.path.exists(wpath): synth.write_weights(cfg, wpath, seed=0, damp_heads=True)
fd = os.dup(2); dv = os.open(os.devnull, os.O_WRONLY); os.dup2(dv, 2)
net = dn.Network(cfg, wpath, precision=dn.PREC_BF16)
os.dup2(fd, 2)
x = synth.make_images(batch, 3, size, size, 1002)
net.set_head_sync(0)
for _ in range(3): net.predict(x)
first, last, desc = net.flows()[which]
print(desc, "layers", first, "..", last)
st = net.flow_stats()
if st[4]: print("SM clock while the first flow ran: %.0f MHz" % (st[3] / st[4] * 1e3))
raw, item0 = net.flow_trace(which)
pair = (raw[:, 4] & np.uint64(0xffff)).astype(np.int64); pos = (raw[:, 4] >> np.uint64(16)).astype(np.int64)
tr = raw[:, :4].astype(np.float64) / 1e3                      # us
t_begin = tr[:, 0].min()
tr -= t_begin
print("flow span %.1f us (first inputs-complete stamp to last tile stored)" % tr[:, 3].max())
members = [i for i in range(first, last + 1) if net.kernel(i).startswith("conv_tc")]
print("%3s %5s %6s | %8s %8s | %7s %7s %7s %7s | %s" % ("lyr", "items", "kxN", "start", "end", "load", "mma", "drain", "mma sum", "per-layer kernel plan"))
tot_mma = 0
for k, li in enumerate(members):
    a, b = item0[k], item0[k + 1]
    t = tr[a:b]
    L = net.layers[li]
    load = (t[:, 1] - t[:, 0]).mean(); mma = (t[:, 2] - t[:, 1]).mean(); drain = (t[:, 3] - t[:, 2]).mean()
    tot_mma += (t[:, 2] - t[:, 1]).sum()
    print("%3d %5d %2dx%-4d | %8.1f %8.1f | %7.2f %7.2f %7.2f %7.0f | %s" % (li, b - a, L["size"], L["n"], t[:, 1].min(), t[:, 3].max(), load, mma, drain,
          (t[:, 2] - t[:, 1]).sum() / 74, dn.lib.b200_layer_plan(net.ptr, li).decode()[:60]))
print("tensor-issue busy: %.1f us per pair of %.1f us = %.0f %%" % (tot_mma / 74, tr[:, 3].max(), 100 * tot_mma / 74 / tr[:, 3].max()))

# per-pair anatomy: tensor-issue idle between consecutive items of a pair, by (kind of the previous item -> kind of the next)
layer_of = np.zeros(len(tr), dtype=np.int64)
for k, li in enumerate(members): layer_of[item0[k]:item0[k + 1]] = li
kind = lambda li: "%dx%d_%d" % (net.layers[li]["size"], net.layers[li]["size"], net.layers[li]["out_w"])
gaps = {}
idle_total = 0.0; last_end = []
for p in range(74):
    idx = np.nonzero(pair == p)[0]
    idx = idx[np.argsort(pos[idx])]
    for a, b in zip(idx[:-1], idx[1:]):
        g = tr[b, 1] - tr[a, 2]
        key = (kind(layer_of[a]), kind(layer_of[b]))
        gaps.setdefault(key, []).append(g)
        idle_total += max(g, 0)
    last_end.append(tr[idx[-1], 3])
print("tensor-issue idle between items: %.1f us per pair; pairs finish between %.1f and %.1f us" % (idle_total / 74, min(last_end), max(last_end)))
for key, v in sorted(gaps.items(), key=lambda kv: -sum(kv[1])):
    v = np.array(v)
    print("  %-10s -> %-10s  n %6d  mean gap %6.2f us  p90 %6.2f  total/pair %7.1f us" % (key[0], key[1], len(v), v.mean(), np.percentile(v, 90), v.clip(0).sum() / 74))
