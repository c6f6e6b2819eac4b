#!/usr/bin/env python3
"""Run the fused path many times on the same batch and require bit-identical records every time (catches races between the
asynchronous roles of the tcgen05 kernels that a single parity run can miss). Usage: determinism_check.py [iters]"""
import hashlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tensorflow_b200 import synth, darknet as dn
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
work = "/tmp/b200_bench"
cfg = synth.make_cfg("yolov3", work, batch=64, width=416, height=416)
wpath = os.path.join(work, "yolov3_seed0_damped.weights")
if not os.path.exists(wpath): synth.write_weights(cfg, wpath, seed=0, damp_heads=True)
fd = os.dup(2); dv = os.open(os.devnull, os.O_WRONLY); os.dup2(dv, 2)
net = dn.Network(cfg, wpath, precision=dn.PREC_BF16)
os.dup2(fd, 2)
net.set_head_sync(0)
x = synth.make_images(64, 3, 416, 416, 1002)
seen = set()
for it in range(iters):
    rec, counts = net.detect_batch(x, 416, 416, .5, .45)
    rec = rec[np.lexsort((rec["cls"], rec["box_id"], rec["image"]))]
    seen.add(hashlib.sha256(rec.tobytes() + counts.tobytes()).hexdigest())
# the same batch through the pipelined serving loop (forward k+1 enqueued while the records of batch k are read back)
import ctypes
RES = ctypes.c_void_p(1)
out = (dn.B200_DET * (1 << 20))(); cnt = (ctypes.c_int * 64)()
dn.lib.b200_submit_batch(net.ptr, RES)
for it in range(iters):
    n = dn.lib.b200_detect_submitted(net.ptr, RES if it + 1 < iters else None, 416, 416, .5, .45, 1, out, 1 << 20, cnt)
    rec = np.ctypeslib.as_array(out)[:n].copy()
    rec = rec[np.lexsort((rec["cls"], rec["box_id"], rec["image"]))]
    seen.add(hashlib.sha256(rec.tobytes() + np.array(list(cnt)).tobytes()).hexdigest())
print("iterations", iters, "x 2 (synchronous + pipelined resident loop), records", len(rec), "distinct results", len(seen))
sys.exit(0 if len(seen) == 1 else 1)
