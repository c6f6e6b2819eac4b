import os, sys, ctypes, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import model_files
from yolo_tensorflow_b200 import synth, darknet as dn
from oracle import np_darknet as P
cfg, wpath = model_files("yolov3-tiny", 1, 416, "/tmp/b200_tests")
net = dn.Network(cfg, wpath, precision=dn.PREC_FP32)
rng = np.random.default_rng(4)
img = (rng.random((120, 200, 3)) * 255).astype(np.uint8)
chw = img.astype(np.float32).transpose(2, 0, 1) / 255.
chw = np.ascontiguousarray(chw)
im = dn.make_image(200, 120, 3)
ctypes.memmove(im.data, chw.ctypes.data, chw.nbytes)
boxed = dn.letterbox_image(im, 416, 416)
lb = np.ctypeslib.as_array(boxed.data, shape=(1, 3, 416, 416)).copy()
dn.predict_image(net.ptr, im)
num = ctypes.c_int(0)
dets = dn.get_network_boxes(net.ptr, 200, 120, .3, .5, None, 0, ctypes.byref(num))
eb, eo, ep = dn.dets_to_arrays(dets, num.value, 80)
port = P.Net(cfg, wpath)
outs = port.forward(lb)
pb, po, pp, _ = P.get_network_boxes(port, outs, 0, 200, 120, .3, relative=0)
print("n", num.value, len(po), "pairs", (ep>0).sum(), (pp>0).sum())
if num.value == len(po):
    print("box err", np.abs(eb-pb).max(), "obj err", np.abs(eo-po).max())
o2, p2 = P.do_nms_obj(pb, po, pp, .45)
o3, p3 = P.do_nms_obj(eb, eo, ep, .45)
dn.do_nms_obj(dets, num.value, 80, .45)
_, eo2, ep2 = dn.dets_to_arrays(dets, num.value, 80)
print("after: engine", (ep2>0).sum(), "port-on-port", (p2>0).sum(), "port-on-engine-dets", (p3>0).sum(), "objs", (eo2>0).sum(), (o2>0).sum(), (o3>0).sum())
