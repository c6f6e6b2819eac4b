"""TEST INFRASTRUCTURE — ctypes harness around the UNMODIFIED reference darknet CPU library
(oracle/_ref/libdarknet_ref.so, built by oracle/Makefile from /root/reference with GPU=0 OPENMP=1).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product never does.  It is the strongest oracle available: the reference's own
parse_network_cfg / load_weights / network_predict / get_network_boxes / do_nms_sort run in-process.

Struct access uses the field offsets measured on the reference header (tests/golden/abi_layout.txt,
reference include/darknet.h:118-495 with GPU undefined).  The fork prints weight dumps to stdout inside
load_weights (parser.c:1087-1100,1176-1227,1317-1335); stdout is redirected around those calls.
"""
import contextlib
import ctypes
import os
import sys
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libdarknet_ref.so")
REF_SO_O2 = os.path.join(HERE, "_ref", "libdarknet_ref_O2.so")

# offsets on x86-64, GPU undefined (tests/golden/abi_layout.txt)
SIZEOF_LAYER = 1160
L_TYPE, L_BATCH, L_INPUTS, L_OUTPUTS = 0, 72, 84, 88
L_H, L_W, L_C, L_OUT_H, L_OUT_W, L_OUT_C, L_N = 108, 112, 116, 120, 124, 128, 132
L_CLASSES, L_OUTPUT = 248, 592
N_N, N_BATCH, N_LAYERS, N_OUTPUT, N_INPUTS, N_OUTPUTS, N_H, N_W, N_C = 0, 4, 32, 40, 128, 132, 144, 148, 152


class BOX(Structure):
    _fields_ = [("x", c_float), ("y", c_float), ("w", c_float), ("h", c_float)]


class DETECTION(Structure):
    _fields_ = [("bbox", BOX), ("classes", c_int), ("prob", POINTER(c_float)), ("mask", POINTER(c_float)),
                ("objectness", c_float), ("sort_class", c_int)]


def available(o2=False):
    return os.path.exists(REF_SO_O2 if o2 else REF_SO)


@contextlib.contextmanager
def _quiet(stdout=True, stderr=True):
    """redirect the C-level stdout/stderr to /dev/null (the fork dumps megabytes of weights)"""
    sys.stdout.flush(); sys.stderr.flush()
    saved = []
    devnull = os.open(os.devnull, os.O_WRONLY)
    for fd, on in ((1, stdout), (2, stderr)):
        if on:
            saved.append((fd, os.dup(fd)))
            os.dup2(devnull, fd)
    try:
        yield
    finally:
        for fd, old in saved:
            os.dup2(old, fd); os.close(old)
        os.close(devnull)


class RefNet:
    def __init__(self, cfg, weights=None, o2=False, threads=None):
        if threads:
            os.environ["OMP_NUM_THREADS"] = str(threads)
        path = REF_SO_O2 if o2 else REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` where /root/reference exists")
        # RTLD_LOCAL: the reference exports the same symbol names as the product library
        self.lib = lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        lib.parse_network_cfg.restype = c_void_p; lib.parse_network_cfg.argtypes = [c_char_p]
        lib.load_weights.argtypes = [c_void_p, c_char_p]
        lib.network_predict.restype = POINTER(c_float); lib.network_predict.argtypes = [c_void_p, POINTER(c_float)]
        lib.get_network_boxes.restype = POINTER(DETECTION)
        lib.get_network_boxes.argtypes = [c_void_p, c_int, c_int, c_float, c_float, POINTER(c_int), c_int, POINTER(c_int)]
        lib.do_nms_sort.argtypes = [POINTER(DETECTION), c_int, c_int, c_float]
        lib.do_nms_obj.argtypes = [POINTER(DETECTION), c_int, c_int, c_float]
        lib.free_detections.argtypes = [POINTER(DETECTION), c_int]
        lib.free_network.argtypes = [c_void_p]
        with _quiet():
            self.ptr = lib.parse_network_cfg(str(cfg).encode())
            if weights:
                lib.load_weights(self.ptr, str(weights).encode())
        self.n = self._net_int(N_N)
        self.batch = self._net_int(N_BATCH)
        self.w, self.h, self.c = self._net_int(N_W), self._net_int(N_H), self._net_int(N_C)
        self.inputs = self._net_int(N_INPUTS)
        self.layers_ptr = ctypes.cast(self.ptr + N_LAYERS, POINTER(c_void_p))[0]

    # -- raw struct access ---------------------------------------------------------------------------
    def _net_int(self, off):
        return ctypes.cast(self.ptr + off, POINTER(c_int))[0]

    def _layer_addr(self, i):
        return self.layers_ptr + i * SIZEOF_LAYER

    def layer_int(self, i, off):
        return ctypes.cast(self._layer_addr(i) + off, POINTER(c_int))[0]

    def set_layer_int(self, i, off, v):
        ctypes.cast(self._layer_addr(i) + off, POINTER(c_int))[0] = v

    def layer_output_ptr(self, i):
        return ctypes.cast(self._layer_addr(i) + L_OUTPUT, POINTER(c_void_p))[0]

    def set_layer_output_ptr(self, i, p):
        ctypes.cast(self._layer_addr(i) + L_OUTPUT, POINTER(c_void_p))[0] = p

    def layer_shape(self, i):
        return dict(type=self.layer_int(i, L_TYPE), batch=self.layer_int(i, L_BATCH), outputs=self.layer_int(i, L_OUTPUTS),
                    out_c=self.layer_int(i, L_OUT_C), out_h=self.layer_int(i, L_OUT_H), out_w=self.layer_int(i, L_OUT_W),
                    n=self.layer_int(i, L_N), classes=self.layer_int(i, L_CLASSES))

    # -- API -------------------------------------------------------------------------------------------
    def predict(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.size == self.batch * self.inputs, (x.size, self.batch, self.inputs)
        with _quiet(stderr=False):
            out = self.lib.network_predict(self.ptr, x.ctypes.data_as(POINTER(c_float)))
        return out

    def layer_output(self, i):
        """copy of layer i's host output, [batch, outputs] (fp32 NCHW flattened)"""
        outputs = self.layer_int(i, L_OUTPUTS)
        p = self.layer_output_ptr(i)
        a = np.ctypeslib.as_array(ctypes.cast(p, POINTER(c_float)), shape=(self.batch * outputs,))
        return a.reshape(self.batch, outputs).copy()

    def head_layers(self):
        YOLO, REGION, DETECTION_T = 23, 22, 5
        return [i for i in range(self.n) if self.layer_int(i, L_TYPE) in (YOLO, REGION, DETECTION_T)]

    def boxes(self, b, w, h, thresh, relative=1):
        """per-image get_network_boxes via the SURVEY §8b trick: offset every head's output pointer to image b and
        set its batch to 1 (the struct is public), call the reference, restore."""
        heads = self.head_layers()
        saved = []
        for i in heads:
            p, bt = self.layer_output_ptr(i), self.layer_int(i, L_BATCH)
            saved.append((i, p, bt))
            self.set_layer_output_ptr(i, p + 4 * b * self.layer_int(i, L_OUTPUTS))
            self.set_layer_int(i, L_BATCH, 1)
        num = c_int(0)
        dets = self.lib.get_network_boxes(self.ptr, w, h, thresh, .5, None, relative, ctypes.byref(num))
        for i, p, bt in saved:
            self.set_layer_output_ptr(i, p); self.set_layer_int(i, L_BATCH, bt)
        return dets, num.value

    def boxes_as_is(self, w, h, thresh, relative=1):
        """the reference's get_network_boxes exactly as a driver calls it (batch item 0; with cfg batch=2 the heads are
        flip-averaged in place first, yolo_layer.c:320 / region_layer.c:368-390)"""
        num = c_int(0)
        dets = self.lib.get_network_boxes(self.ptr, w, h, thresh, .5, None, relative, ctypes.byref(num))
        return dets, num.value

    def resize(self, w, h):
        """the reference's resize_network (network.c:358-438)"""
        self.lib.resize_network.argtypes = [c_void_p, c_int, c_int]; self.lib.resize_network.restype = c_int
        with _quiet():
            rc = self.lib.resize_network(self.ptr, w, h)
        self.w, self.h = self._net_int(N_W), self._net_int(N_H)
        self.inputs = self._net_int(N_INPUTS)
        return rc

    def classes(self):
        return self.layer_int(self.n - 1, L_CLASSES)

    def dets_arrays(self, dets, n):
        classes = self.classes()
        boxes = np.zeros((n, 4), np.float32); obj = np.zeros(n, np.float32); probs = np.zeros((n, classes), np.float32)
        for i in range(n):
            d = dets[i]
            boxes[i] = (d.bbox.x, d.bbox.y, d.bbox.w, d.bbox.h)
            obj[i] = d.objectness
            probs[i] = np.ctypeslib.as_array(d.prob, shape=(classes,))
        return boxes, obj, probs

    def nms_sort(self, dets, n, thresh):
        self.lib.do_nms_sort(dets, n, self.classes(), thresh)

    def free_dets(self, dets, n):
        self.lib.free_detections(dets, n)

    def close(self):
        if self.ptr:
            self.lib.free_network(self.ptr); self.ptr = None


def ref_nms_sort_arrays(boxes, probs, thresh, objectness=None, o2=False, want_order=False):
    """run the reference do_nms_sort on plain arrays -> probs after suppression, rows in INPUT order
    (want_order: also the input row that ended up at each position of the reference's re-sorted array)"""
    path = REF_SO_O2 if o2 else REF_SO
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    lib.do_nms_sort.argtypes = [POINTER(DETECTION), c_int, c_int, c_float]
    n, classes = probs.shape
    arr = (DETECTION * n)()
    keep_alive = []
    for i in range(n):
        p = (c_float * classes)(*probs[i].tolist())
        keep_alive.append(p)
        arr[i].bbox = BOX(*[float(v) for v in boxes[i]])
        arr[i].classes = classes
        arr[i].prob = ctypes.cast(p, POINTER(c_float))
        arr[i].objectness = float(objectness[i]) if objectness is not None else 1.0
        arr[i].sort_class = i                      # scratch: overwritten by the call; identity is recovered via prob pointer
    addr_to_row = {ctypes.addressof(p): i for i, p in enumerate(keep_alive)}
    lib.do_nms_sort(arr, n, classes, thresh)
    out = np.zeros_like(probs)
    order = []
    for j in range(n):
        row = addr_to_row[ctypes.cast(arr[j].prob, c_void_p).value]
        out[row] = np.ctypeslib.as_array(arr[j].prob, shape=(classes,))
        order.append(row)
    return (out, order) if want_order else out
