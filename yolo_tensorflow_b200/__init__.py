"""B200-native YOLOv1/v2/v3 inference engine behind darknet's C API.

The product is `lib/libdarknet.so` (host C + hand-written sm_100a CUDA, built by `csrc/Makefile`);
this package is the thin ctypes mirror of the reference's `python/darknet.py` plus synthetic
workload generators used by the tests and `bench.py`.
"""
from . import darknet  # noqa: F401
