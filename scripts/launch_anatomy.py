#!/usr/bin/env python3
"""Fixed versus per-tile cost of one tcgen05 convolution launch: a single layer (stem + the layer under test) timed at
growing batch sizes, i.e. a growing number of tile waves, with CUDA events around the layer (b200_profile_layers, 20 passes).
T(batch) = fixed + per_wave * waves: the intercept is what every one of the 74 launches of a YOLOv3 step pays for pipeline
fill, drain and the kernel boundary.  Usage: launch_anatomy.py"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tensorflow_b200 import synth, darknet as dn

WORK = "/tmp/b200_anatomy"
os.makedirs(WORK, exist_ok=True)
CASES = [  # name, h=w, cin, filters, size
    ("13x13 1x1 1024->512", 13, 1024, 512, 1),
    ("13x13 3x3 512->1024", 13, 512, 1024, 3),
    ("26x26 1x1 512->256", 26, 512, 256, 1),
    ("26x26 3x3 256->512", 26, 256, 512, 3),
    ("52x52 1x1 256->128", 52, 256, 128, 1),
    ("52x52 3x3 128->256", 52, 128, 256, 3),
]


def cfg_for(path, hw, c, filters, size, batch):
    text = (f"[net]\nbatch={batch}\nsubdivisions=1\nheight={hw}\nwidth={hw}\nchannels=3\nmomentum=0.9\ndecay=0.0005\nlearning_rate=0.001\n"
            "max_batches=1\npolicy=constant\n"
            f"[convolutional]\nbatch_normalize=1\nfilters={c}\nsize=3\nstride=1\npad=1\nactivation=leaky\n"
            f"[convolutional]\nbatch_normalize=1\nfilters={filters}\nsize={size}\nstride=1\npad=1\nactivation=leaky\n")
    open(path, "w").write(text)


for name, hw, c, filters, size in CASES:
    rows = []
    for batch in (16, 32, 64, 128, 256):
        cfg = os.path.join(WORK, "one.cfg"); w = os.path.join(WORK, "one.weights")
        cfg_for(cfg, hw, c, filters, size, batch)
        synth.write_weights(cfg, w, seed=1, damp_heads=False)
        fd = os.dup(2); dv = os.open(os.devnull, os.O_WRONLY); os.dup2(dv, 2)
        try:
            net = dn.Network(cfg, w, precision=dn.PREC_BF16)
        finally:
            os.dup2(fd, 2); os.close(fd); os.close(dv)
        net.predict(synth.make_images(batch, 3, hw, hw, 3))
        ms = net.profile_layers(20)
        plan = dn.lib.b200_layer_plan(net.ptr, 1).decode()
        flops = 2.0 * filters * size * size * c * hw * hw * batch
        rows.append((batch, float(ms[1]) * 1e3, flops / (float(ms[1]) * 1e-3) / 1e12, plan.split("smem")[0][-70:]))
        net.close()
    print("==", name)
    for b, us, tf, plan in rows:
        print(f"  batch {b:4d}: {us:8.2f} us  {tf:7.1f} TFLOP/s   {plan}")
    # least-squares line through (batch, us): intercept = fixed cost of a launch
    x = np.array([r[0] for r in rows], float); y = np.array([r[1] for r in rows], float)
    k, b0 = np.polyfit(x, y, 1)
    print(f"  fit: {b0:.2f} us fixed + {k * 64:.2f} us per 64 images")
