"""Pins the oracle.  oracle/np_darknet.py (the numpy restatement of the reference's CPU path) is checked
against (1) the committed golden fixtures that tests/golden/make_golden.py produced by RUNNING the unmodified
reference CPU build, and (2) the live reference library oracle/_ref when it is present on the box.
Tolerance: 2e-5 of each layer's abs-max (the port's only intended difference is the fp32 summation order of
the conv GEMM); detections identical in count and identity; NMS keep-lists bit-exact."""
import os
import sys

import numpy as np
import pytest

from conftest import golden_probs, load_golden, model_files

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import np_darknet as P  # noqa: E402
from oracle import ref_darknet as R  # noqa: E402
from yolo_tensorflow_b200 import synth  # noqa: E402

CASES = ["yolov3-tiny_96_b2", "yolov3-tiny_416_b1", "yolov3_96_b1", "yolov2_96_b2", "yolov1_448_b1"]


@pytest.mark.parametrize("name", CASES)
def test_port_matches_reference_golden(name, workdir):
    g = load_golden(name)
    model, size, batch = str(g["model"]), int(g["size"]), int(g["batch"])
    cfg, wpath = model_files(model, batch, size, workdir)
    net = P.Net(cfg, wpath)
    outs = net.forward(synth.make_images(batch, 3, size, size, int(g["seed"])))
    assert len(outs) == int(g["nlayers"])
    for i, o in enumerate(outs):
        o = o.reshape(batch, -1)
        ref = g[f"layer{i}_val"]
        got = o[:, g[f"layer{i}_idx"]]
        tol = 2e-5 * float(g[f"layer{i}_absmax"]) + 1e-7
        assert np.abs(got - ref).max() <= tol, (i, net.layers[i].type)
        if f"head{i}" in g.files:
            assert np.abs(o - g[f"head{i}"]).max() <= tol
    classes = net.layers[-1].classes
    w_, h_ = (1, 1) if model == "yolov1" else (size, size)
    for b in range(batch):
        boxes, obj, probs, _ = P.get_network_boxes(net, outs, b, w_, h_, float(g["thresh"]))
        assert len(obj) == len(g[f"img{b}_obj"])
        if len(obj):
            np.testing.assert_allclose(boxes, g[f"img{b}_boxes"], rtol=2e-3, atol=1e-5)
            np.testing.assert_allclose(obj, g[f"img{b}_obj"], rtol=1e-4, atol=1e-6)
            gp = golden_probs(g, b, classes)
            # a probability within fp32 noise of the threshold may flip between 0 and p; everything else agrees
            close = np.isclose(probs, gp, rtol=1e-3, atol=1e-5)
            border = np.abs(np.maximum(probs, gp) - float(g["thresh"])) < 1e-4
            assert (close | border).all()
        # NMS: the port's do_nms_sort on the REFERENCE's boxes reproduces the reference keep-list exactly
        kept = P.do_nms_sort(g[f"img{b}_boxes"], g[f"img{b}_obj"], golden_probs(g, b, classes), float(g["nms"]))
        rc = np.stack(np.nonzero(kept)).astype(np.int32)
        assert np.array_equal(rc, g[f"img{b}_kept_rc"])


FLIP_CASES = ["yolov3-tiny_96_flip", "yolov2_96_flip", "yolov3-tiny_105_flip"]


def flip_golden_arrays(g, classes):
    n = int(g["num"])
    probs = np.zeros((n, classes), np.float32)
    probs[g["prob_rc"][0], g["prob_rc"][1]] = g["prob_v"]
    return g["boxes"], g["obj"], probs


@pytest.mark.parametrize("name", FLIP_CASES)
def test_port_flip_average_matches_reference_golden(name, workdir):
    """cfg batch=2 (`detector valid2`): the reference's get_network_boxes rewrites l.output (item 1 mirrored, item 0 averaged)
    and decodes item 0; the golden holds l.output before/after and the returned array (counted BEFORE the average)"""
    g = load_golden(name)
    model, size, thresh = str(g["model"]), int(g["size"]), float(g["thresh"])
    cfg, wpath = model_files(model, 2, size, workdir)
    net = P.Net(cfg, wpath)
    heads = [int(i) for i in g["heads"]]
    outs = [None] * len(net.layers)
    for i in heads:
        outs[i] = g[f"head{i}_before"]                      # the reference's own activations: only the flip/average is under test
    av = P.avg_flipped(net, outs)
    for i in heads:
        assert np.array_equal(av[i], g[f"head{i}_after"]), i
    boxes, obj, probs, _ = P.get_network_boxes(net, av, 0, size, size, thresh)
    gb, go, gp = flip_golden_arrays(g, net.layers[-1].classes)
    filled = len(obj)
    assert filled <= int(g["num"])
    np.testing.assert_allclose(boxes, gb[:filled], rtol=1e-5, atol=1e-7)
    assert np.array_equal(obj, go[:filled])
    assert np.allclose(probs, gp[:filled], rtol=1e-5, atol=1e-7)
    assert not go[filled:].any() and not gp[filled:].any()  # records the averaged output no longer fills stay calloc'd zeros


def test_reorg_is_the_reference_permutation():
    """blas.c:9-30 with forward=0 (reorg_layer.c:107-109) on c=8,h=w=4,s=2: a permutation that is NOT space_to_depth"""
    x = np.arange(8 * 4 * 4, dtype=np.float32).reshape(1, 8, 4, 4)
    y = P.reorg_cpu(x, 4, 4, 8, 2, 0).reshape(-1)
    assert sorted(y.tolist()) == list(range(128))
    s2d = x.reshape(1, 8, 2, 2, 2, 2).transpose(0, 3, 5, 1, 2, 4).reshape(-1)
    assert not np.array_equal(y, s2d)
    # first outputs, computed by hand from the index formula: out[in_index] = x[out_index]
    assert y[:4].tolist() == [0.0, 2.0, 4.0, 6.0]


def test_iou_matrix_equals_scalar_box_iou():
    rng = np.random.default_rng(3)
    b = rng.random((40, 4)).astype(np.float32)
    m = P.iou_matrix(b)
    for i in range(0, 40, 7):
        for j in range(40):
            assert m[i, j] == P.box_iou(b[i], b[j]) or (np.isnan(m[i, j]) and np.isnan(P.box_iou(b[i], b[j])))


def test_nms_edge_cases():
    boxes = np.array([[.5, .5, .2, .2]] * 3 + [[.1, .1, .05, .05]], np.float32)
    probs = np.array([[.9, 0], [.8, .7], [.6, .9], [.5, .5]], np.float32)
    out = P.do_nms_sort(boxes, np.ones(4, np.float32), probs, .45)
    assert out.tolist() == [[np.float32(.9), 0], [0, 0], [0, np.float32(.9)], [.5, .5]]
    # objectness 0 rows are partitioned away (box.c:60-70) and keep their probs untouched
    out = P.do_nms_sort(boxes, np.array([0, 1, 1, 1], np.float32), probs, .45)
    assert out[0].tolist() == [np.float32(.9), 0] and out[1, 0] == np.float32(.8)
    # empty input and zero-area boxes (union 0 -> NaN never suppresses)
    assert P.do_nms_sort(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), np.zeros((0, 3), np.float32), .45).shape == (0, 3)
    z = np.zeros((2, 4), np.float32)
    assert (P.do_nms_sort(z, np.ones(2, np.float32), np.ones((2, 1), np.float32), .45) == 1).all()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built on this box")
@pytest.mark.parametrize("model,size,batch,seed", [("yolov3-tiny", 160, 2, 7), ("yolov2", 128, 1, 8), ("yolov3", 128, 1, 9)])
def test_port_matches_live_reference(model, size, batch, seed, workdir):
    cfg, wpath = model_files(model, batch, size, workdir)
    ref = R.RefNet(cfg, wpath)
    x = synth.make_images(batch, 3, size, size, seed)
    ref.predict(x)
    net = P.Net(cfg, wpath)
    outs = net.forward(x)
    for i, o in enumerate(outs):
        r = ref.layer_output(i)
        assert np.abs(o.reshape(batch, -1) - r).max() <= 2e-5 * np.abs(r).max() + 1e-7, i
    for b in range(batch):
        dets, n = ref.boxes(b, size, size, .1)
        rb, ro, rp = ref.dets_arrays(dets, n)
        pb, po, pp, _ = P.get_network_boxes(net, outs, b, size, size, .1)
        assert n == len(po)
        ref_after = R.ref_nms_sort_arrays(rb, rp, .45, ro)
        assert np.array_equal(ref_after, P.do_nms_sort(rb, ro, rp, .45))
        ref.free_dets(dets, n)
    ref.close()


@pytest.mark.skipif(not (R.available() and R.available(o2=True)), reason="oracle/_ref not built on this box")
def test_reference_is_independent_of_fast_math_and_threads(workdir):
    """SURVEY §8c determinism: -Ofast vs -O2 builds give bit-identical activations and NMS keep-lists"""
    cfg, wpath = model_files("yolov3-tiny", 1, 160, workdir)
    x = synth.make_images(1, 3, 160, 160, 11)
    a, b = R.RefNet(cfg, wpath), R.RefNet(cfg, wpath, o2=True)
    a.predict(x); b.predict(x)
    for i in a.head_layers():
        assert np.array_equal(a.layer_output(i), b.layer_output(i))
    rng = np.random.default_rng(5)
    boxes = np.concatenate([rng.random((200, 2)), rng.random((200, 2)) * .3 + .05], axis=1).astype(np.float32)
    probs = (rng.random((200, 5)) * (rng.random((200, 5)) > .5)).astype(np.float32)
    assert np.array_equal(R.ref_nms_sort_arrays(boxes, probs, .45), R.ref_nms_sort_arrays(boxes, probs, .45, o2=True))
    a.close(); b.close()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built on this box")
def test_reference_nms_leaves_the_array_in_stable_chain_order():
    """do_nms_sort re-sorts the struct array once per class with qsort (box.c:72-77).  On glibc that is a stable merge sort,
    so the order a driver sees afterwards (draw_detections prints in it) is well defined: pinned here on the reference build,
    restated by P.nms_sort_final_order and reproduced by the product's do_nms_sort (tests/test_gpu_parity.py)."""
    rng = np.random.default_rng(12)
    n, classes = 400, 6
    boxes = np.concatenate([rng.random((n, 2)), rng.random((n, 2)) * .3 + .02], axis=1).astype(np.float32)
    probs = (rng.random((n, classes)) * (rng.random((n, classes)) < .3)).astype(np.float32)       # many zero ties
    probs[rng.random((n, classes)) < .1] = np.float32(.5)                                          # and equal non-zero scores
    obj = (rng.random(n) > .15).astype(np.float32)
    _, order = R.ref_nms_sort_arrays(boxes, probs, .45, obj, want_order=True)
    assert order == P.nms_sort_final_order(obj, probs)


def tree_model(workdir, batch=1):
    """the small YOLO9000-style network of the tree golden: cfg (with the repo's WordTree fixture as tree=) and seeded weights"""
    tree_path = os.path.join(REPO, "tests", "golden", "wordtree_240.tree")
    cfg = synth.make_tree_cfg(workdir, tree_path, batch=batch, size=32)
    wpath = os.path.join(workdir, "yolo9000-small.weights")
    if not os.path.exists(wpath):
        tmp = wpath + ".tmp%d" % os.getpid()
        synth.write_weights(cfg, tmp, seed=0, damp_heads=True)
        os.replace(tmp, wpath)
    return cfg, wpath


def tree_golden_arrays(g, tag, classes):
    n = int(g[f"{tag}_num"])
    probs = np.zeros((n, classes), np.float32)
    probs[g[f"{tag}_prob_rc"][0], g[f"{tag}_prob_rc"][1]] = g[f"{tag}_prob_v"]
    return g[f"{tag}_boxes"], g[f"{tag}_obj"], probs


def test_port_wordtree_matches_reference_golden(workdir):
    """YOLO9000: the port's per-group softmax against the reference's own softmax() (blas.c:305) and the port's
    hierarchy_predictions / hierarchy_top_prediction / map branch against the reference's get_network_boxes
    (region_layer.c:412-424, tree.c:37-81) on the same head activations"""
    g = load_golden("yolo9000-small_tree")
    cfg, wpath = tree_model(workdir)
    net = P.Net(cfg, wpath)
    L = net.layers[-1]
    assert L.tree.n == 240 and L.tree.groups == 73
    outs = net.forward(synth.make_images(1, 3, 32, 32, int(g["seed"])))
    assert np.abs(outs[-1].reshape(1, -1) - g["head"]).max() <= 1e-6
    got = outs[-1].reshape(L.n, L.coords + L.classes + 1, L.h * L.w)[1, L.coords + 1:, 37]
    assert np.abs(got - g["box_softmax"]).max() <= 1e-6                      # the reference's softmax() on every sibling group
    thresh = float(g["thresh"])
    for tag, hier, cmap in (("top", .5, None), ("top_lo", .1, None), ("map", .5, g["map"])):
        o = [None] * (len(net.layers) - 1) + [g["head"].copy()]
        o = P.hierarchy_predictions(net, o)
        assert np.array_equal(o[-1], g[f"{tag}_head_after"])                 # in place, bit for bit
        boxes, obj, probs, _ = P.get_network_boxes(net, o, 0, 32, 32, thresh, hier=hier, map=cmap)
        gb, go, gp = tree_golden_arrays(g, tag, L.classes)
        assert len(obj) == int(g[f"{tag}_num"])
        np.testing.assert_allclose(boxes, gb, rtol=1e-5, atol=1e-7)
        assert np.array_equal(obj, go) and np.array_equal(probs, gp)
