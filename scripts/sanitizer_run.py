import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_tensorflow_b200 import synth, darknet as dn
work = "/tmp/b200_san"; os.makedirs(work, exist_ok=True)
for model, size, batch in (("yolov3", 160, 2), ("yolov3-tiny", 96, 3), ("yolov2", 160, 2)):
    cfg = synth.make_cfg(model, work, batch=batch, width=size, height=size)
    wpath = os.path.join(work, model + ".weights")
    if not os.path.exists(wpath): synth.write_weights(cfg, wpath, seed=0, damp_heads=True)
    net = dn.Network(cfg, wpath, precision=dn.PREC_BF16)
    x = synth.make_images(batch, 3, size, size, 3)
    rec, counts = net.detect_batch(x, size, size, .3, .45)
    print(model, len(rec), counts.tolist())
    u8 = [np.random.default_rng(i).integers(0, 256, (120 + 7 * i, 200 - 11 * i, 3), dtype=np.uint8) for i in range(batch)]
    net.letterbox_batch_u8(u8)
    rec, counts = net.detect_batch(None, 0, 0, .3, .45, relative=0)
    print(" letterboxed", len(rec))
    import tempfile
    imgs = [u8[i % batch] for i in range(2 * batch + 1)]
    n = net.validate_images(imgs, ["im_%d.jpg" % i for i in range(len(imgs))], "coco", tempfile.mkdtemp(), thresh=.3)
    print(" validate driver (pipelined, short last batch)", n)
    net.close()

# ---- round 2: the reference-API extras and the restructured NMS -------------------------------------------------------------
import ctypes
# batch == 2 flip-average through get_network_boxes; get_network_boxes after the host rewrote l.output
cfg = synth.make_cfg("yolov3-tiny", work, batch=2, width=105, height=105)
net = dn.Network(cfg, os.path.join(work, "yolov3-tiny.weights"), precision=dn.PREC_FP32)
x0 = synth.make_images(1, 3, 105, 105, 5)
net.predict(np.ascontiguousarray(np.concatenate([x0, x0[..., ::-1]])))
num = ctypes.c_int(0)
dets = dn.get_network_boxes(net.ptr, 105, 105, .2, .5, None, 1, ctypes.byref(num))
dn.do_nms_sort(dets, num.value, 80, .45)
dn.free_detections(dets, num.value)
print("flip-average", num.value)
net.close()
# YOLO9000-style head: softmax per sibling group, hierarchy_predictions, top prediction and map
tree = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "wordtree_240.tree")
cfg = synth.make_tree_cfg(work, tree, batch=2, size=32)
w9 = os.path.join(work, "y9k.weights"); synth.write_weights(cfg, w9, seed=0, damp_heads=True)
net = dn.Network(cfg, w9, precision=dn.PREC_BF16)
net.predict(synth.make_images(2, 3, 32, 32, 6))
cmap = (np.arange(200) % 240).astype(np.int32)
for mp in (None, cmap.ctypes.data_as(ctypes.POINTER(ctypes.c_int))):
    dets = dn.get_network_boxes(net.ptr, 32, 32, .05, .5, mp, 1, ctypes.byref(num))
    dn.do_nms_sort(dets, num.value, 240, .45)
    dn.free_detections(dets, num.value)
print("wordtree", num.value)
net.close()
# NMS: every path of the general kernel — matrix in shared memory (<= 512 survivors), lists in shared memory + chunked rows
# (<= 1024), lists in the HBM slab, more than 32 words per row, more than 32768 survivors (removed set in the slab)
rng = np.random.default_rng(1)
for n, classes in ((300, 3), (513, 1), (1025, 1), (2049, 2), (5000, 1), (33000, 1)):
    boxes = np.concatenate([rng.random((n, 2)), rng.random((n, 2)) * (.05 if n > 4000 else .3) + .01], axis=1).astype(np.float32)
    probs = (rng.random((n, classes)) + .01).astype(np.float32)
    out = dn.nms_sort_arrays(boxes, probs, .45)
    print("nms", n, classes, int((out > 0).sum()))
# resize_network + reorg table + tiled region head
cfg = synth.make_cfg("yolov2", work, batch=2, width=160, height=160)
net = dn.Network(cfg, os.path.join(work, "yolov2.weights"), precision=dn.PREC_BF16)
net.resize(224, 192)
rec, counts = net.detect_batch(synth.make_images(2, 3, 192, 224, 7), 224, 192, .3, .45)
print("resized yolov2", len(rec))
net.close()
# round 2, second part: the flow kernel (B200_FLOW=1) and the im2col-mode loads at odd sizes / strides
os.environ["B200_FLOW"] = "1"
cfg = synth.make_cfg("yolov3", work, batch=3, width=160, height=96)
net = dn.Network(cfg, os.path.join(work, "yolov3.weights"), precision=dn.PREC_BF16)
os.environ.pop("B200_FLOW")
print("flows", [(a, b) for a, b, _ in net.flows()])
x = synth.make_images(3, 3, 96, 160, 9)
for _ in range(2):
    rec, counts = net.detect_batch(x, 160, 96, .3, .45)
print("flow detect", len(rec), net.flow_stats()[:3])
net.close()
