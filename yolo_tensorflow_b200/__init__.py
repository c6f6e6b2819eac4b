"""B200-native YOLOv1/v2/v3 inference engine behind darknet's C API.

The product is `lib/libdarknet.so` (host C + hand-written sm_100a CUDA, built by `csrc/Makefile`);
this package is the thin ctypes mirror of the reference's `python/darknet.py` (`yolo_tensorflow_b200.darknet`,
which maps the native library when it is imported) plus synthetic workload generators (`synth`) and the
image-sharding helpers (`shard`) used by the tests and `bench.py`.

Submodules are imported on first use: `from yolo_tensorflow_b200 import synth` does not map the native library
(bench.py's reference arm relies on that), `yolo_tensorflow_b200.darknet` does and fails loudly when it is missing.
"""
import importlib

__all__ = ["darknet", "synth", "shard"]


def __getattr__(name):
    if name in __all__:
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
