import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")
WORK = os.environ.get("B200_TEST_WORKDIR", "/tmp/b200_tests")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """the product is a native library; build it in-tree if it is not there yet (nvcc cross-compiles without a GPU)"""
    so = os.path.join(REPO, "yolo_tensorflow_b200", "lib", "libdarknet.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(REPO, "yolo_tensorflow_b200", "csrc"), "-j8"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


@pytest.fixture(scope="session")
def workdir():
    os.makedirs(WORK, exist_ok=True)
    return WORK


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def model_files(model, batch, size, workdir, damp=True):
    """cfg + seeded synthetic weights shared by oracle and engine"""
    from yolo_tensorflow_b200 import synth
    cfg = synth.make_cfg(model, workdir, batch=batch, width=size, height=size)
    wpath = os.path.join(workdir, f"{model}{'' if damp else '_raw'}.weights")
    if not os.path.exists(wpath):
        tmp = wpath + ".tmp%d" % os.getpid()
        synth.write_weights(cfg, tmp, seed=0, damp_heads=damp)
        os.replace(tmp, wpath)
    return cfg, wpath


def golden_probs(g, b, classes):
    n = len(g[f"img{b}_obj"])
    p = np.zeros((n, classes), np.float32)
    rc = g[f"img{b}_prob_rc"]
    p[rc[0], rc[1]] = g[f"img{b}_prob_v"]
    return p
