#!/usr/bin/env python3
"""What bf16 activation storage alone costs, measured on the CPU: the oracle's numpy port run twice on the same image — once as
the reference computes (fp32 everywhere) and once with the engine's rounding points emulated (weights and every stored
activation rounded to bf16, fp32 accumulation, fp32 folded batch-norm, head logits kept fp32) — and the decoded boxes of the
two compared by identity.  This is the derivation behind the free-running bf16 tolerances in tests/test_gpu_parity.py and
DESIGN.md §2: they are a property of the number format over 75 stacked convolutions, not of the kernels.
  python scripts/bf16_error_model.py [model] [size] [seed]            (CPU only)
  python scripts/bf16_error_model.py yolov3 416 1002 --engine         (GPU box: adds the engine's own free-running numbers)
TEST INFRASTRUCTURE: imports oracle/."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import np_darknet as P  # noqa: E402
from yolo_tensorflow_b200 import synth  # noqa: E402


def r16(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float32).numpy()


def forward_bf16(net, x):
    """the port with the engine's rounding points: bf16 weights, bf16 stored activations, fp32 head logits"""
    outs, cur = [], r16(x)
    for i, L in enumerate(net.layers):
        nxt = net.layers[i + 1].type if i + 1 < len(net.layers) else ""
        if L.type in ("convolutional", "conv", "local", "connected", "conn"):
            L2 = P.Layer(L); L2.weights = r16(L.weights)
            cur = P.FORWARD[L.type](net, L2, cur, outs)
            if nxt not in ("yolo", "region", "detection"):
                cur = r16(cur)
        else:
            cur = P.FORWARD[L.type](net, L, cur, outs)
            if L.type not in ("yolo", "region", "detection"):
                cur = r16(cur)
        outs.append(cur)
    return outs


def compare(tag, net, ref_outs, got_heads, size, thresh):
    """got_heads: {layer index: [1, outputs] head activations}; prints logit/activation and matched-box error statistics"""
    outs = list(ref_outs)
    for i, a in got_heads.items():
        r = ref_outs[i].ravel(); a = a.ravel()
        print(f"{tag} head {i}: max|err|/max|ref| {np.abs(a - r).max() / np.abs(r).max():.4f}  rms(err)/rms(ref) "
              f"{np.sqrt(((a - r) ** 2).mean()) / np.sqrt((r ** 2).mean()):.4f}")
        outs[i] = a.reshape(ref_outs[i].shape)
    rb, ro, rp, rid = P.get_network_boxes(net, ref_outs, 0, size, size, thresh)
    gb, go, gp, gid = P.get_network_boxes(net, outs, 0, size, size, thresh)
    ri = {int(v): k for k, v in enumerate(rid)}; gi = {int(v): k for k, v in enumerate(gid)}
    common = sorted(set(ri) & set(gi))
    print(f"{tag} candidates ref {len(ri)} got {len(gi)} common {len(common)}; unmatched objectness: ref-only "
          f"{[round(float(ro[ri[v]]), 4) for v in set(ri) - set(gi)]} got-only {[round(float(go[gi[v]]), 4) for v in set(gi) - set(ri)]}")
    A = np.array([rb[ri[v]] for v in common]); B = np.array([gb[gi[v]] for v in common])
    rel = np.abs(A - B) / np.abs(A)
    print(f"{tag} x,y abs err max {np.abs(A - B)[:, :2].max():.5f}; w,h rel err median {np.median(rel[:, 2:]):.4f} max {rel[:, 2:].max():.4f}")
    pa = np.array([rp[ri[v]] for v in common]); pb = np.array([gp[gi[v]] for v in common])
    both = (pa > 0) & (pb > 0); mm = (pa > 0) != (pb > 0)
    print(f"{tag} prob abs err max {np.abs(pa - pb)[both].max():.4f} median {np.median(np.abs(pa - pb)[both]):.5f}; "
          f"{int(mm.sum())} of {int(both.sum()) + int(mm.sum())} (box, class) pairs on different sides of the threshold, "
          f"their scores within {np.abs(np.maximum(pa, pb)[mm] - thresh).max() if mm.any() else 0:.4f} of it")


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    model = args[0] if args else "yolov3"
    size = int(args[1]) if len(args) > 1 else 416
    seed = int(args[2]) if len(args) > 2 else 1002
    thresh = .2 if model == "yolov1" else .5
    work = "/tmp/b200_bench"
    cfg = synth.make_cfg(model, work, batch=1, width=size, height=size)
    wpath = os.path.join(work, f"{model}_seed0_damped.weights")
    if not os.path.exists(wpath):
        synth.write_weights(cfg, wpath, seed=0, damp_heads=True)
    net = P.Net(cfg, wpath)
    x = synth.make_images(1, 3, size, size, seed)
    ref = net.forward(x)
    heads = [i for i, L in enumerate(net.layers) if L.type in ("yolo", "region", "detection")]
    emu = forward_bf16(net, x)
    compare("[bf16 model]", net, ref, {i: emu[i] for i in heads}, size, thresh)
    if "--engine" in sys.argv:
        from yolo_tensorflow_b200 import darknet as dn
        fd = os.dup(2); dv = os.open(os.devnull, os.O_WRONLY); os.dup2(dv, 2)
        eng = dn.Network(cfg, wpath, precision=dn.PREC_BF16)
        os.dup2(fd, 2)
        eng.predict(x)
        compare("[engine]    ", net, ref, {i: eng.layer_output(i) for i in heads}, size, thresh)
        compare("[engine vs bf16 model]", net, emu, {i: eng.layer_output(i) for i in heads}, size, thresh)
