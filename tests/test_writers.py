"""Result writers: the library's COCO / VOC / ImageNet result files against the oracle's restatement of
examples/detector.c:157-232.  Host C only: no GPU needed."""
import ctypes
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import np_darknet as P  # noqa: E402


def make_records(dn, rng, n, images, classes, sizes):
    rec = np.zeros(n, dtype=np.dtype(dn.B200_DET))
    rec["image"] = rng.integers(0, images, n)
    rec["cls"] = rng.integers(0, classes, n)
    rec["box_id"] = rng.permutation(n * 3)[:n]
    rec["prob"] = rng.random(n).astype(np.float32)
    rec["prob"][::17] = 0.0                                             # suppressed entries are skipped
    for i in range(n):
        w, h = sizes[rec["image"][i]]
        # boxes around and across the image borders: exercises every clamp
        rec["bbox"]["x"][i] = rng.uniform(-0.1, 1.1) * w; rec["bbox"]["y"][i] = rng.uniform(-0.1, 1.1) * h
        rec["bbox"]["w"][i] = rng.uniform(0.01, 0.8) * w; rec["bbox"]["h"][i] = rng.uniform(0.01, 0.8) * h
    return rec


def as_tuples(rec):
    return [(int(r["image"]), int(r["box_id"]), int(r["cls"]), np.float32(r["prob"]),
             (r["bbox"]["x"], r["bbox"]["y"], r["bbox"]["w"], r["bbox"]["h"])) for r in rec]


def c_strings(items):
    arr = (ctypes.c_char_p * len(items))(*[s.encode() for s in items])
    return arr


def test_writers_match_the_reference_formats(built_library, tmp_path):
    from yolo_tensorflow_b200 import darknet as dn
    lib = dn.lib
    rng = np.random.default_rng(3)
    sizes = [(500, 375), (640, 480), (333, 500), (97, 61)]
    widths = (ctypes.c_int * 4)(*[s[0] for s in sizes]); heights = (ctypes.c_int * 4)(*[s[1] for s in sizes])
    paths = ["/data/coco/images/val2014/COCO_val2014_000000000042.jpg", "val/COCO_val2014_000000581929.jpg", "x_7.png", "/a/b/000123.jpg"]
    assert [lib.b200_coco_image_id(p.encode()) for p in paths] == [42, 581929, 7, 123] == [P.coco_image_id(p) for p in paths]
    rec = make_records(dn, rng, 400, 4, 80, sizes)
    want = as_tuples(rec)
    ptr = rec.ctypes.data_as(ctypes.POINTER(dn.B200_DET))

    coco = tmp_path / "coco_results.json"
    assert lib.b200_append_coco(str(coco).encode(), ptr, len(rec), c_strings(paths), widths, heights) == 0
    assert coco.read_text() == P.print_cocos(want, paths, [s[0] for s in sizes], [s[1] for s in sizes])

    ids = ["2008_000001", "2008_000002", "2008_000003", "2008_000004"]
    names = ["c%02d" % j for j in range(20)]
    rec20 = make_records(dn, rng, 300, 4, 20, sizes)
    prefix = str(tmp_path / "comp4_det_test_")
    assert lib.b200_append_voc(prefix.encode(), c_strings(names), 20, rec20.ctypes.data_as(ctypes.POINTER(dn.B200_DET)), len(rec20),
                               c_strings(ids), widths, heights) == 0
    voc = P.print_detector_detections(as_tuples(rec20), ids, [s[0] for s in sizes], [s[1] for s in sizes], 20)
    for j, name in enumerate(names):
        assert open(prefix + name + ".txt").read() == voc[j]

    imnet = tmp_path / "imagenet-detection.txt"
    image_ids = (ctypes.c_int * 4)(11, 12, 13, 14)
    assert lib.b200_append_imagenet(str(imnet).encode(), ptr, len(rec), image_ids, widths, heights) == 0
    assert imnet.read_text() == P.print_imagenet_detections(want, [11, 12, 13, 14], [s[0] for s in sizes], [s[1] for s in sizes])


def test_oracle_writers_pinned_on_the_reference_build(tmp_path):
    """print_detector_detections / print_imagenet_detections of the UNMODIFIED examples/detector.c (oracle/_ref/libdetector_ref.so)
    against the oracle's restatement, on detection arrays with the reference's own struct layout"""
    import pytest
    from oracle import ref_darknet as R
    so = os.path.join(os.path.dirname(R.REF_SO), "libdetector_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libdetector_ref.so not built on this box")
    ref = ctypes.CDLL(so)
    libc = ctypes.CDLL(None)
    libc.fopen.restype = ctypes.c_void_p; libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    libc.fclose.argtypes = [ctypes.c_void_p]
    rng = np.random.default_rng(5)
    classes, total, w, h = 20, 60, 500, 375
    dets = (R.DETECTION * total)()
    keep = []
    probs = []
    records = []
    for i in range(total):
        box = (rng.uniform(-0.1, 1.1) * w, rng.uniform(-0.1, 1.1) * h, rng.uniform(0.01, 0.8) * w, rng.uniform(0.01, 0.8) * h)
        dets[i].bbox.x, dets[i].bbox.y, dets[i].bbox.w, dets[i].bbox.h = box
        pr = (rng.random(classes) * (rng.random(classes) < .2)).astype(np.float32)
        probs.append(pr)                                                    # keep the buffers alive
        dets[i].prob = pr.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        dets[i].classes = classes
        b32 = tuple(np.float32(v) for v in box)
        records += [(0, i, j, pr[j], b32) for j in range(classes) if pr[j]]
    # VOC: one file per class
    files = [str(tmp_path / ("voc_%02d.txt" % j)) for j in range(classes)]
    fps = (ctypes.c_void_p * classes)(*[libc.fopen(f.encode(), b"w") for f in files])
    ref.print_detector_detections.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p, ctypes.POINTER(R.DETECTION), ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int]
    ref.print_detector_detections(fps, b"2008_000001", dets, total, classes, w, h)
    for fp in fps:
        libc.fclose(fp)
    want = P.print_detector_detections(records, ["2008_000001"], [w], [h], classes)
    for j in range(classes):
        assert open(files[j]).read() == want[j]
    # ImageNet
    path = str(tmp_path / "imagenet.txt")
    fp = libc.fopen(path.encode(), b"w")
    ref.print_imagenet_detections.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(R.DETECTION), ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int]
    ref.print_imagenet_detections(fp, 77, dets, total, classes, w, h)
    libc.fclose(fp)
    assert open(path).read() == P.print_imagenet_detections(records, [77], [w], [h])
