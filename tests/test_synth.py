"""synthetic workload generator: file layout (parser.c:1241-1345 order) and determinism"""
import os
import struct
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from yolo_tensorflow_b200 import synth  # noqa: E402


def test_weights_file_layout_and_determinism(tmp_path):
    cfg = synth.make_cfg("yolov3-tiny", str(tmp_path), batch=2, width=96, height=96)
    a, b = str(tmp_path / "a.weights"), str(tmp_path / "b.weights")
    n = synth.write_weights(cfg, a, seed=0)
    synth.write_weights(cfg, b, seed=0)
    raw = open(a, "rb").read()
    assert raw == open(b, "rb").read()
    assert struct.unpack("<iiiQ", raw[:20]) == (0, 2, 0, 0)
    assert len(raw) == 20 + 4 * n
    layers = synth.walk_shapes(cfg)["layers"]
    expect = sum(L["n"] * (1 + 3 * L["bn"]) + L["n"] * L["size"] ** 2 * L["c"] for L in layers if L["type"] == "convolutional")
    assert n == expect == 8_858_734
    first = np.frombuffer(raw[20:20 + 4 * 16], dtype=np.float32)           # layer 0 biases ~ N(0, .1)
    assert np.abs(first).max() < 1.0


def test_make_cfg_rewrites_only_net_section(tmp_path):
    cfg = synth.make_cfg("yolov3", str(tmp_path), batch=64, width=608, height=608)
    secs = synth.read_cfg(cfg)
    assert secs[0][1]["batch"] == "64" and secs[0][1]["width"] == "608" and secs[0][1]["height"] == "608"
    assert len(secs) == 108
    shapes = synth.walk_shapes(cfg)["layers"]
    assert shapes[82]["out"] == (255, 19, 19) and shapes[106]["out"] == (255, 76, 76)


def test_images_are_seeded():
    a, b = synth.make_images(2, 3, 8, 8, 1002), synth.make_images(2, 3, 8, 8, 1002)
    assert np.array_equal(a, b) and a.dtype == np.float32 and 0 <= a.min() and a.max() < 1
