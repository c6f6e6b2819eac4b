#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the UNMODIFIED reference CPU build (oracle/_ref, built from
/root/reference by oracle/Makefile) on seeded synthetic cfg/weights/images.  Run in the authoring container
(where /root/reference exists); the fixtures are committed, /root/reference never travels to the GPU box.

  python tests/golden/make_golden.py            # writes the fixtures
  python tests/golden/make_golden.py --calibrate  # prints the undamped head-logit std used by synth.HEAD_RAW_STD

Each fixture holds, for one (model, size, batch): a strided sample of EVERY layer's output, the full head
outputs, the reference's get_network_boxes result per image and its do_nms_sort result (identity-indexed).
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from yolo_tensorflow_b200 import synth
from oracle import ref_darknet as R

WORK = "/tmp/b200_golden"
CASES = [  # name, model, size, batch, image seed, thresh, nms
    ("yolov3-tiny_96_b2", "yolov3-tiny", 96, 2, 1000, .5, .45),
    ("yolov3-tiny_416_b1", "yolov3-tiny", 416, 1, 1000, .5, .45),       # BASELINE configs[0]
    ("yolov3_96_b1", "yolov3", 96, 1, 1002, .05, .45),
    ("yolov2_96_b2", "yolov2", 96, 2, 1001, .05, .45),
    ("yolov1_448_b1", "yolov1", 448, 1, 1004, .2, .4),
]
SAMPLE = 512


def sample_idx(n):
    return np.unique(np.linspace(0, n - 1, min(n, SAMPLE)).astype(np.int64))


def build(name, model, size, batch, seed, thresh, nms):
    cfg = synth.make_cfg(model, WORK, batch=batch, width=size, height=size)
    wpath = os.path.join(WORK, f"{model}.weights")
    if not os.path.exists(wpath):
        synth.write_weights(cfg, wpath, seed=0, damp_heads=True)
    net = R.RefNet(cfg, wpath)
    x = synth.make_images(batch, 3, size, size, seed)
    net.predict(x)
    d = dict(model=model, size=size, batch=batch, seed=seed, thresh=thresh, nms=nms, nlayers=net.n)
    heads = net.head_layers()
    for i in range(net.n):
        o = net.layer_output(i)
        idx = sample_idx(o.shape[1])
        d[f"layer{i}_idx"] = idx
        d[f"layer{i}_val"] = o[:, idx]
        d[f"layer{i}_absmax"] = np.abs(o).max()
        if i in heads and (size <= 96 or model == "yolov1"):
            d[f"head{i}"] = o
    d["heads"] = np.array(heads)
    w_, h_ = (1, 1) if model == "yolov1" else (size, size)
    for b in range(batch):
        dets, n = net.boxes(b, w_, h_, thresh)
        boxes, obj, probs = net.dets_arrays(dets, n)
        after = R.ref_nms_sort_arrays(boxes, probs, nms, obj)
        d[f"img{b}_boxes"], d[f"img{b}_obj"] = boxes, obj
        nz = np.nonzero(probs)
        d[f"img{b}_prob_rc"] = np.stack(nz).astype(np.int32)
        d[f"img{b}_prob_v"] = probs[nz]
        d[f"img{b}_kept_rc"] = np.stack(np.nonzero(after)).astype(np.int32)
        net.free_dets(dets, n)
    net.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, "layers", d["nlayers"], "dets/img", [len(d[f"img{b}_obj"]) for b in range(batch)],
          "kept", [d[f"img{b}_kept_rc"].shape[1] for b in range(batch)])


FLIP_CASES = [  # name, model, size, image seed, thresh  -- cfg batch=2: item 1 is the mirrored item 0 (`detector valid2`)
    ("yolov3-tiny_96_flip", "yolov3-tiny", 96, 1000, .3),
    ("yolov2_96_flip", "yolov2", 96, 1001, .05),
    ("yolov3-tiny_105_flip", "yolov3-tiny", 105, 1002, .3),             # odd grid width: the middle column keeps its sign
]


def build_flip(name, model, size, seed, thresh):
    """get_network_boxes with l.batch == 2: avg_flipped_yolo / the region twin rewrite l.output in place, then item 0 is decoded"""
    from oracle import np_darknet as P
    cfg = synth.make_cfg(model, WORK, batch=2, width=size, height=size)
    wpath = os.path.join(WORK, f"{model}.weights")
    if not os.path.exists(wpath):
        synth.write_weights(cfg, wpath, seed=0, damp_heads=True)
    x0 = synth.make_images(1, 3, size, size, seed)
    x = np.ascontiguousarray(np.concatenate([x0, x0[..., ::-1]]))
    # the reference overruns its allocation when the averaged output has more boxes over the threshold than the un-averaged
    # item 0 (count first, average inside the fill): check with the port that this input does not do that
    port = P.Net(cfg, wpath)
    outs = port.forward(x)
    before = len(P.get_network_boxes(port, outs, 0, size, size, thresh)[1])
    after = len(P.get_network_boxes(port, P.avg_flipped(port, outs), 0, size, size, thresh)[1])
    assert after <= before, (name, before, after)
    net = R.RefNet(cfg, wpath)
    net.predict(x)
    heads = net.head_layers()
    d = dict(model=model, size=size, seed=seed, thresh=thresh, heads=np.array(heads))
    for i in heads:
        d[f"head{i}_before"] = net.layer_output(i)
    dets, n = net.boxes_as_is(size, size, thresh)
    boxes, obj, probs = net.dets_arrays(dets, n)
    for i in heads:
        d[f"head{i}_after"] = net.layer_output(i)
    d["boxes"], d["obj"], d["num"] = boxes, obj, n
    nz = np.nonzero(probs)
    d["prob_rc"] = np.stack(nz).astype(np.int32); d["prob_v"] = probs[nz]
    net.free_dets(dets, n); net.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, "num", n, "filled", int((obj != 0).sum()) if model != "yolov2" else n, "port before/after", before, after)


THRESH_TREE = .05


def build_tree(name="yolo9000-small_tree"):
    """YOLO9000 box extraction pinned on the reference: its CPU forward cannot produce class probabilities for a
    [region] layer with tree= (softmax at l.temperature = 0, region_layer.c:179), so the head activations come from the oracle's
    port (GPU-build semantics), are written into the reference network's l.output, and the REFERENCE's get_network_boxes —
    hierarchy_predictions + hierarchy_top_prediction / the map branch, region_layer.c:412-424 — turns them into detections.
    Also holds the reference's own softmax() (blas.c:305) on every sibling group of one box, which pins the port's group softmax."""
    import ctypes
    from oracle import np_darknet as P
    tree_path = os.path.join(HERE, "wordtree_240.tree")
    cfg = synth.make_tree_cfg(WORK, tree_path, batch=1, size=32)
    wpath = os.path.join(WORK, "yolo9000-small.weights")
    synth.write_weights(cfg, wpath, seed=0, damp_heads=True)
    port = P.Net(cfg, wpath)
    x = synth.make_images(1, 3, 32, 32, 1009)
    outs = port.forward(x)
    L = port.layers[-1]
    head = outs[-1].reshape(1, -1).astype(np.float32)
    d = dict(seed=1009, size=32, thresh=THRESH_TREE, head=head, logits=outs[-2].reshape(1, -1))
    print("objectness quantiles", np.quantile(head.reshape(L.n, -1, L.h * L.w)[:, 4], [.5, .9, .99, 1.]))
    # the reference's softmax() on the sibling groups of box (anchor 1, cell 37)
    ref = R.RefNet(cfg, wpath)
    ref.lib.softmax.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
    hw = L.h * L.w
    logit = outs[-2].reshape(L.n, L.coords + L.classes + 1, hw)[1, L.coords + 1:, 37].astype(np.float32).copy()
    sm = np.zeros_like(logit)
    for g in range(L.tree.groups):
        a, n = L.tree.group_offset[g], L.tree.group_size[g]
        if n:
            ref.lib.softmax(logit[a:].ctypes.data_as(ctypes.POINTER(ctypes.c_float)), n, 1.0, 1, sm[a:].ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    d["box_logits"], d["box_softmax"] = logit, sm
    rng = np.random.default_rng(8)
    cmap = rng.integers(0, L.classes, 200).astype(np.int32)
    d["map"] = cmap
    ref.predict(x)
    i = ref.n - 1
    for tag, hier, mp in (("top", .5, None), ("top_lo", .1, None), ("map", .5, cmap)):
        buf = np.ctypeslib.as_array(ctypes.cast(ref.layer_output_ptr(i), ctypes.POINTER(ctypes.c_float)), shape=(head.size,))
        buf[:] = head.ravel()                                        # hierarchy_predictions works in place: start from fresh activations
        num = ctypes.c_int(0)
        mptr = mp.ctypes.data_as(ctypes.POINTER(ctypes.c_int)) if mp is not None else None
        dets = ref.lib.get_network_boxes(ref.ptr, 32, 32, ctypes.c_float(THRESH_TREE), ctypes.c_float(hier), mptr, 1, ctypes.byref(num))
        boxes, obj, probs = ref.dets_arrays(dets, num.value)
        nz = np.nonzero(probs)
        d[f"{tag}_boxes"], d[f"{tag}_obj"], d[f"{tag}_num"] = boxes, obj, num.value
        d[f"{tag}_prob_rc"] = np.stack(nz).astype(np.int32); d[f"{tag}_prob_v"] = probs[nz]
        d[f"{tag}_head_after"] = buf.copy().reshape(1, -1)
        ref.free_dets(dets, num.value)
        print(name, tag, "boxes", num.value, "with a class", len(set(nz[0].tolist())), "distinct classes", len(set(nz[1].tolist())))
    ref.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)


def calibrate():
    for model, seed in (("yolov3-tiny", 1000), ("yolov3", 1002), ("yolov2", 1001), ("yolov1", 1004)):
        cfg = synth.make_cfg(model, WORK, batch=1)
        wpath = os.path.join(WORK, f"{model}_raw.weights")
        synth.write_weights(cfg, wpath, seed=0, damp_heads=False)
        net = R.RefNet(cfg, wpath)
        net.predict(synth.make_images(1, 3, net.h, net.w, seed))
        print(model, [float(net.layer_output(i - 1).std()) for i in net.head_layers()])
        net.close()


if __name__ == "__main__":
    os.makedirs(WORK, exist_ok=True)
    if "--calibrate" in sys.argv:
        calibrate()
    else:
        only = [a for a in sys.argv[1:] if not a.startswith("-")]
        for c in CASES:
            if not only or c[0] in only:
                build(*c)
        for c in FLIP_CASES:
            if not only or c[0] in only:
                build_flip(*c)
        if not only or "yolo9000-small_tree" in only:
            build_tree()
