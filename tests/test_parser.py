"""cfg parser: the stderr layer table is the reference's structural golden (V*/yolov*.txt);
shapes, batch semantics and error behaviour of the host C code.  No GPU needed."""
import os
import subprocess
import sys
import textwrap

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLES = os.path.join(REPO, "tests", "golden", "layer_tables")

PARSE = textwrap.dedent("""
    import sys
    sys.path.insert(0, %r)
    from yolo_tensorflow_b200 import darknet as dn
    net = dn.Network(sys.argv[1])
    print(net.n, net.batch, net.w, net.h)
    for L in net.layers:
        print(L["type_name"], L["out_w"], L["out_h"], L["out_c"], L["outputs"], L["nweights"])
    net.close()
""" % REPO)


def run_parse(cfg):
    return subprocess.run([sys.executable, "-c", PARSE, cfg], capture_output=True, text=True)


@pytest.mark.parametrize("model", ["yolov3", "yolov2", "yolov1", "yolov3-tiny"])
def test_layer_table_matches_reference_printout(model):
    r = run_parse(os.path.join(REPO, "cfg", model + ".cfg"))
    assert r.returncode == 0, r.stderr
    golden = open(os.path.join(TABLES, model + ".txt")).read().splitlines()
    golden[0] = golden[0].lstrip()          # the txt files carry a copy-paste indent on the header row only
    assert r.stderr.splitlines() == golden


@pytest.mark.parametrize("model", ["yolov3", "yolov2", "yolov1", "yolov3-tiny"])
def test_reference_parser_agrees_when_available(model, tmp_path):
    sys.path.insert(0, REPO)
    from oracle import ref_darknet as R
    if not R.available():
        pytest.skip("oracle/_ref not built on this box")
    code = "import sys; sys.path.insert(0, %r); from oracle import ref_darknet as R; import os; n = R.RefNet(sys.argv[1]); " % REPO
    # RefNet silences the C stderr; call the raw parser instead
    code = ("import ctypes, sys; lib = ctypes.CDLL(%r); lib.parse_network_cfg.restype = ctypes.c_void_p; "
            "lib.parse_network_cfg.argtypes = [ctypes.c_char_p]; lib.parse_network_cfg(sys.argv[1].encode())" % R.REF_SO)
    ref = subprocess.run([sys.executable, "-c", code, os.path.join(REPO, "cfg", model + ".cfg")], capture_output=True, text=True)
    ours = run_parse(os.path.join(REPO, "cfg", model + ".cfg"))
    assert ref.stderr == ours.stderr


def test_shapes_and_weight_counts_yolov3():
    r = run_parse(os.path.join(REPO, "cfg", "yolov3.cfg"))
    rows = r.stdout.splitlines()
    assert rows[0] == "107 1 416 416"
    layers = [x.split() for x in rows[1:]]
    assert sum(1 for x in layers if x[0] == "CONVOLUTIONAL") == 75
    assert sum(int(x[5]) for x in layers if x[0] == "CONVOLUTIONAL") == 61_895_776     # 61.9 M conv weights
    assert layers[82] == ["YOLO", "13", "13", "255", str(13 * 13 * 255), "0"]
    assert layers[106][:4] == ["YOLO", "52", "52", "255"]
    assert layers[86][:4] == ["ROUTE", "26", "26", "768"]


RESIZE = textwrap.dedent("""
    import sys
    sys.path.insert(0, %r)
    from yolo_tensorflow_b200 import darknet as dn
    net = dn.Network(sys.argv[1])
    rc = net.resize(int(sys.argv[2]), int(sys.argv[2]))
    print(rc, net.w, net.h)
    for L in net.layers:
        print(L["type_name"], L["w"], L["h"], L["out_w"], L["out_h"], L["out_c"], L["inputs"], L["outputs"])
    net.close()
""" % REPO)


@pytest.mark.parametrize("model,size", [("yolov3", 608), ("yolov3", 320), ("yolov2", 608), ("yolov3-tiny", 352)])
def test_resize_network_geometry(model, size, tmp_path):
    """resize_network (network.c:358-438) must leave the layer table as a fresh parse at that size would, and as the
    reference's own resize leaves it."""
    sys.path.insert(0, REPO)
    from yolo_tensorflow_b200 import synth
    cfg = os.path.join(REPO, "cfg", model + ".cfg")
    r = subprocess.run([sys.executable, "-c", RESIZE, cfg, str(size)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = r.stdout.splitlines()
    assert rows[0] == "0 %d %d" % (size, size)
    fresh = synth.make_cfg(model, str(tmp_path), batch=1, width=size, height=size)
    f = subprocess.run([sys.executable, "-c", RESIZE, fresh, str(size)], capture_output=True, text=True)
    assert f.stdout == r.stdout
    from oracle import ref_darknet as R
    if not R.available():
        return
    code = textwrap.dedent("""
        import ctypes, sys
        sys.path.insert(0, %r)
        from oracle import ref_darknet as R
        net = R.RefNet(sys.argv[1])
        net.lib.resize_network.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        with R._quiet():
            rc = net.lib.resize_network(net.ptr, int(sys.argv[2]), int(sys.argv[2]))
        print(rc)
        for i in range(net.n):
            s = net.layer_shape(i)
            print(s["out_w"], s["out_h"], s["out_c"], s["outputs"])
    """ % REPO)
    ref = subprocess.run([sys.executable, "-c", code, cfg, str(size)], capture_output=True, text=True)
    assert ref.returncode == 0, ref.stderr
    ours = ["0"] + [" ".join(x.split()[3:6] + x.split()[7:8]) for x in rows[1:]]
    theirs = ref.stdout.splitlines()
    for i, row in enumerate(rows[1:], 1):
        if row.split()[0] in ("YOLO", "REGION"):
            # resize_yolo_layer/resize_region_layer (yolo_layer.c:63, region_layer.c:56) update w/h/outputs but leave the
            # stale out_w/out_h behind; here they follow the new size like a fresh parse.  Compare what both define.
            ours[i] = ours[i].split()[-1]
            theirs[i] = theirs[i].split()[-1]
    assert theirs == ours


def test_resize_network_refuses_fixed_size_layers():
    r = subprocess.run([sys.executable, "-c", RESIZE, os.path.join(REPO, "cfg", "yolov1.cfg"), "224"], capture_output=True, text=True)
    assert r.stdout.splitlines()[0].split()[0] == "-1"       # local/connected/detection: network.c:428
    assert "Cannot resize this type of layer" in r.stderr


def test_batch_and_subdivisions(tmp_path):
    cfg = tmp_path / "b.cfg"
    text = open(os.path.join(REPO, "cfg", "yolov3-tiny.cfg")).read().replace("batch=1", "batch=8").replace("subdivisions=1", "subdivisions=4")
    cfg.write_text(text)
    r = run_parse(str(cfg))
    assert r.stdout.splitlines()[0].split()[1] == "2"       # parser.c:652 net->batch /= subdivisions


def test_whitespace_comments_and_defaults(tmp_path):
    cfg = tmp_path / "w.cfg"
    cfg.write_text("# comment\n[net]\n batch = 1 \nheight=32\nwidth = 32\nchannels=3\n; other comment\n\n[convolutional]\nfilters = 8\nsize=3\n stride=1\npad=1\nactivation=leaky\nbogus=1\n[maxpool]\nsize=2\nstride=2\n")
    r = run_parse(str(cfg))
    assert r.returncode == 0
    assert "Unused field: 'bogus = 1'" in r.stderr                 # option_list.c:85
    assert "learning_rate: Using default '0.001000'" in r.stderr   # option_list.c:138
    assert r.stdout.splitlines()[1].split()[:4] == ["CONVOLUTIONAL", "32", "32", "8"]
    assert r.stdout.splitlines()[2].split()[:4] == ["MAXPOOL", "16", "16", "8"]


def test_missing_file_behaves_like_file_error():
    r = run_parse("/nonexistent/x.cfg")
    assert r.returncode == 0                                       # utils.c:281-285: message + exit(0)
    assert "Couldn't open file: /nonexistent/x.cfg" in r.stderr


def test_unknown_section_is_reported(tmp_path):
    cfg = tmp_path / "u.cfg"
    cfg.write_text("[net]\nbatch=1\nheight=8\nwidth=8\nchannels=3\n[frobnicate]\nx=1\n")
    r = run_parse(str(cfg))
    assert "Type not recognized: [frobnicate]" in r.stderr         # parser.c:823-825
    assert r.returncode != 0


def test_compute_without_gpu_fails_loudly(tmp_path):
    """no CPU fallback: predicting on a box without a CUDA device aborts with a clear message"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    code = PARSE.replace("net.close()", "import numpy as np; net.predict(np.zeros((1,3,416,416), np.float32))")
    r = subprocess.run([sys.executable, "-c", code, os.path.join(REPO, "cfg", "yolov3-tiny.cfg")], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr.lower() or "There is no CPU fallback" in r.stderr
