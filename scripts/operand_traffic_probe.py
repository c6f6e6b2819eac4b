#!/usr/bin/env python3
"""What does operand traffic (L2 -> SM) cost a tensor-bound layer?  One 26x26 3x3 256->512 layer at batch 256 (CTA-pair kernel,
268 us) timed cold (20 passes after a rest) and sustained (~1 s back to back), with the B200_EXP knob of the pair kernel:
0 = normal, 1 = weight stages re-used stale (no B loads), 2 = no A loads, 3 = neither (pure tensor pipe + epilogue).
Results of the experiment variants are garbage by design; only their time matters.  The knobs are compiled in only by
`make -C yolo_tensorflow_b200/csrc clean all EXPERIMENTS=1`.  Usage: B200_EXP=n [B200_EXP_STAGES=k] operand_traffic_probe.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tensorflow_b200 import synth, darknet as dn
WORK = "/tmp/b200_anatomy"; os.makedirs(WORK, exist_ok=True)
hw, c, filters, size, batch = 26, 256, 512, 3, 256
if len(sys.argv) > 1: hw, c, filters, size, batch = [int(v) for v in sys.argv[1:6]]
cfg = os.path.join(WORK, "probe.cfg"); w = os.path.join(WORK, "probe.weights")
open(cfg, "w").write(f"[net]\nbatch={batch}\nsubdivisions=1\nheight={hw}\nwidth={hw}\nchannels=3\nmomentum=0.9\ndecay=0.0005\nlearning_rate=0.001\nmax_batches=1\npolicy=constant\n"
                     f"[convolutional]\nbatch_normalize=1\nfilters={c}\nsize=3\nstride=1\npad=1\nactivation=leaky\n"
                     f"[convolutional]\nbatch_normalize=1\nfilters={filters}\nsize={size}\nstride=1\npad=1\nactivation=leaky\n")
synth.write_weights(cfg, w, seed=1, damp_heads=False)
fd = os.dup(2); dv = os.open(os.devnull, os.O_WRONLY); os.dup2(dv, 2)
net = dn.Network(cfg, w, precision=dn.PREC_BF16)
os.dup2(fd, 2)
img = synth.make_images(batch, 3, hw, hw, 3)
if os.environ.get("PROBE_ZERO_INPUT"): img = np.zeros_like(img)       # every pixel row of A identical: does DATA (bit toggling) change the time?
net.predict(img)
flops = 2.0 * filters * size * size * c * hw * hw * batch
import torch
stream = torch.cuda.ExternalStream(net.stream_ptr())
def run(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(n): dn.lib.b200_run_layers(net.ptr, 1, 2)
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
run(5); time.sleep(3)
cold = run(20)
sus = run(4000)
sus2 = run(1000)
print("B200_EXP=%s zero_input=%s  %s" % (os.environ.get("B200_EXP", "0"), os.environ.get("PROBE_ZERO_INPUT", "0"), dn.lib.b200_layer_plan(net.ptr, 1).decode()[:100]))
print("  cold %.1f us (%.0f TFLOP/s)   sustained %.1f us (%.0f TFLOP/s), then %.1f us" % (cold, flops / cold / 1e6, sus, flops / sus / 1e6, sus2))
