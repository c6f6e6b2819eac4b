"""The C-ABI shared library must load without a GPU and export every entry point include/*.h declares."""
import ctypes
import os
import re

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = []
    for m in re.finditer(r"^[A-Za-z_][\w \*]*?[\s\*]+(\w+)\s*\([^;{]*\)\s*;", text, flags=re.M):
        name = m.group(1)
        if name not in ("forward", "backward", "update", "forward_gpu", "backward_gpu", "update_gpu"):
            names.append(name)
    return names


def test_every_declared_symbol_is_exported(built_library):
    lib = ctypes.CDLL(built_library)
    missing = []
    total = 0
    for h in ("darknet.h", "b200_engine.h"):
        for name in declared_functions(os.path.join(REPO, "include", h)):
            total += 1
            if not hasattr(lib, name):
                missing.append(f"{h}:{name}")
    assert total > 50
    assert not missing, missing


def test_hot_path_symbols_and_global(built_library):
    lib = ctypes.CDLL(built_library)
    for name in ("parse_network_cfg", "load_weights", "network_predict", "get_network_boxes", "do_nms_sort",
                 "free_detections", "load_network", "set_batch_network", "free_network", "make_network_boxes",
                 "do_nms_obj", "network_predict_image", "letterbox_image", "cuda_set_device"):
        assert hasattr(lib, name), name
    assert ctypes.c_int.in_dll(lib, "gpu_index").value == 0


def test_reference_python_wrapper_surface(built_library):
    """every attribute the reference's python/darknet.py binds at import time (python/darknet.py:48-115) resolves"""
    lib = ctypes.CDLL(built_library)
    for name in ("network_width", "network_height", "network_predict", "cuda_set_device", "make_image", "get_network_boxes",
                 "make_network_boxes", "free_detections", "free_ptrs", "reset_rnn", "load_network", "do_nms_obj",
                 "do_nms_sort", "free_image", "letterbox_image", "get_metadata", "load_image_color", "rgbgr_image",
                 "network_predict_image"):
        assert hasattr(lib, name), name
