"""The reference's own C driver, unchanged, on top of this library ("existing .cfg/.weights drivers drop in").

oracle/Makefile extracts the text of `test_detector` (examples/detector.c:562-627) at build time and compiles it twice without
touching it: against the reference header + CPU library (oracle/_ref/test_detector_ref) and against include/darknet.h +
libdarknet.so (oracle/_ref/test_detector_b200).  Both binaries are git-ignored build products that travel with the snapshot;
/root/reference is not needed at run time.  The GPU test runs both on the same cfg / weights / names / glyphs / image and
compares the detections they print (draw_detections' "<name>: <pct>%" lines, image.c:255) and the picture they save.
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from conftest import model_files  # noqa: E402

REF_BIN = os.path.join(REPO, "oracle", "_ref", "test_detector_ref")
OUR_BIN = os.path.join(REPO, "oracle", "_ref", "test_detector_b200")
needs_bins = pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(OUR_BIN)),
                                reason="oracle/_ref driver binaries not built (make -C oracle where /root/reference exists)")


def write_ppm(path, rgb):
    h, w, _ = rgb.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h) + np.ascontiguousarray(rgb, np.uint8).tobytes())


def make_driver_dir(root, model, size, workdir):
    """cfg, weights, names, .data file, data/labels glyphs (PNM data under the .png names load_alphabet asks for; the
    reference's stb_image sniffs the format) and one input picture"""
    os.makedirs(os.path.join(root, "data", "labels"), exist_ok=True)
    rng = np.random.default_rng(5)
    for j in range(8):
        gh, gw = 6 * (j + 2), 4 * (j + 2)
        for ch in range(32, 127):
            glyph = np.where(rng.random((gh, gw, 1)) < .4, 0, 255).astype(np.uint8).repeat(3, axis=2)
            write_ppm(os.path.join(root, "data", "labels", "%d_%d.png" % (ch, j)), glyph)
    cfg, wpath = model_files(model, 1, size, workdir)
    names = os.path.join(root, "synthetic.names")
    with open(names, "w") as f:
        f.write("".join("class%02d\n" % i for i in range(80)))
    data = os.path.join(root, "synthetic.data")
    with open(data, "w") as f:
        f.write("classes= 80\nnames = %s\nbackup = /tmp\n" % names)
    img = (np.random.default_rng(9).random((300, 420, 3)) * 255).astype(np.uint8)
    ppm = os.path.join(root, "input.ppm")
    write_ppm(ppm, img)
    return data, cfg, wpath, ppm


def run_driver(binary, root, data, cfg, wpath, ppm, thresh, out, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([binary, data, cfg, wpath, ppm, str(thresh), out], cwd=root, capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    # the fork dumps the weights to stdout inside load_weights: keep only draw_detections' lines
    dets = [m.groups() for m in (re.fullmatch(r"(class\d\d): (\d+)%", ln.strip()) for ln in r.stdout.splitlines()) if m]
    assert any("Predicted in" in ln for ln in r.stdout.splitlines())
    return [(n, int(p)) for n, p in dets]


@needs_bins
def test_driver_binary_links_against_the_product_library():
    """no GPU needed: the unchanged reference function compiled against include/darknet.h resolves every symbol in libdarknet.so"""
    r = subprocess.run([OUR_BIN], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    ldd = subprocess.run(["ldd", OUR_BIN], capture_output=True, text=True).stdout
    assert "libdarknet.so" in ldd and "not found" not in ldd


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("model,size,thresh", [("yolov3-tiny", 416, .3), ("yolov2", 416, .3)])
def test_reference_test_detector_runs_unchanged(model, size, thresh, workdir, tmp_path):
    root = str(tmp_path)
    data, cfg, wpath, ppm = make_driver_dir(root, model, size, workdir)
    theirs = run_driver(REF_BIN, root, data, cfg, wpath, ppm, thresh, os.path.join(root, "ref_pred"))
    ours = run_driver(OUR_BIN, root, data, cfg, wpath, ppm, thresh, os.path.join(root, "our_pred"), env={"B200_PRECISION": "fp32"})
    assert len(theirs) > 3
    # same detections in the same order (do_nms_sort leaves the array ordered by the last class's score, box.c:72-77, and
    # draw_detections walks it in that order); a percentage may differ by one where fp32 noise crosses a rounding boundary
    assert [n for n, _ in ours] == [n for n, _ in theirs]
    assert all(abs(a - b) <= 1 for (_, a), (_, b) in zip(ours, theirs))
    from PIL import Image
    a = np.asarray(Image.open(os.path.join(root, "our_pred.png")).convert("RGB"), np.int32)
    b = np.asarray(Image.open(os.path.join(root, "ref_pred.png")).convert("RGB"), np.int32)
    assert a.shape == b.shape
    assert (np.abs(a - b) > 1).mean() < 2e-3              # boxes and labels drawn in the same places with the same colours
    # the tcgen05 path through the same unchanged driver: same classes detected
    bf = run_driver(OUR_BIN, root, data, cfg, wpath, ppm, thresh, os.path.join(root, "bf16_pred"))
    assert len(set(n for n, _ in bf) & set(n for n, _ in theirs)) >= 0.7 * len(set(n for n, _ in theirs))
