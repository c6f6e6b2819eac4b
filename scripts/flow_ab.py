#!/usr/bin/env python3
"""A/B of the flow kernels against one launch per layer in ONE process, alternating segments of back-to-back forward passes
(each ~1 s, so the power-capped steady state is what is compared) with nvidia-smi clocks and power sampled per segment.
Usage: flow_ab.py [model] [batch] [size] [rounds] [passes per segment]"""
import os, sys, time, subprocess, threading
os.environ.setdefault("B200_FLOW", "1")
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tensorflow_b200 import synth, darknet as dn
model = sys.argv[1] if len(sys.argv) > 1 else "yolov3"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
size = int(sys.argv[3]) if len(sys.argv) > 3 else 416
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 3
passes = int(sys.argv[5]) if len(sys.argv) > 5 else 200
rest = float(sys.argv[6]) if len(sys.argv) > 6 else 0.     # idle seconds before every segment (> 0: the burst regime of a short bench run)
warm = int(sys.argv[7]) if len(sys.argv) > 7 else 20
samples = []
def sampler():
    p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "25"],
                         stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        try:
            a, b = line.split(",")
            samples.append((time.time(), float(a), float(b)))
        except Exception:
            pass
threading.Thread(target=sampler, daemon=True).start()
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
net.predict(x); net.predict(x)
res = {0: [], 1: []}
for r in range(rounds):
    for on in (0, 1):
        net.set_flow(on)
        if rest > 0: time.sleep(rest)
        net.profile_forward(warm)
        t0 = time.time()
        ms, first = net.profile_forward(passes)
        t1 = time.time()
        seg = [(c, w) for (t, c, w) in samples if t0 + (0.1 if passes > 50 else 0.) <= t <= t1]
        clk = np.median([c for c, w in seg]) if seg else float("nan")
        pw = np.median([w for c, w in seg]) if seg else float("nan")
        res[on].append(ms)
        print("round %d flow %d: forward %.4f ms  (SM %.0f MHz, %.0f W, %d samples)" % (r, on, ms, clk, pw, len(seg)), flush=True)
print("per-layer launches: median %.4f ms; flows: median %.4f ms; ratio %.3f" % (np.median(res[0]), np.median(res[1]), np.median(res[1]) / np.median(res[0])))
