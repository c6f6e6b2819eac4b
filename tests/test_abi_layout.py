"""The public structs must be layout-identical to the reference header built with GPU undefined
(SURVEY.md §8b): drivers read layer/network fields directly and copy `layer` by value.
tests/golden/abi_layout.txt was printed by tests/abi_probe.c compiled against the REFERENCE
include/darknet.h (Darknet2Tensorflow/darknet-master/include/darknet.h:118-525)."""
import os
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_struct_layout_matches_reference(tmp_path):
    exe = tmp_path / "abi_probe"
    subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), os.path.join(REPO, "tests", "abi_probe.c"), "-o", str(exe)], check=True)
    ours = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    golden = open(os.path.join(REPO, "tests", "golden", "abi_layout.txt")).read()
    assert ours == golden
    assert "sizeof layer 1160 network 272 detection 48 image 24 box 16" in ours


def test_layout_against_live_reference_header(tmp_path):
    ref = "/root/reference/Darknet2Tensorflow/darknet-master/include"
    if not os.path.isdir(ref):
        import pytest
        pytest.skip("reference tree not present on this box")
    exe = tmp_path / "abi_probe_ref"
    subprocess.run(["gcc", "-I", ref, os.path.join(REPO, "tests", "abi_probe.c"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    assert out == open(os.path.join(REPO, "tests", "golden", "abi_layout.txt")).read()
