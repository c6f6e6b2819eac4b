import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
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
