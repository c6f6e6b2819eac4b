"""CPU checks of the flow scheduler (conv_tc_plan.cu: flow_schedule): the per-pair item lists it produces must contain every
tile of every layer exactly once and must be executable — replayed WITHOUT the cost model, a pair running its next item only when
the completion counters that item waits on are complete, everything runs (no deadlock).  The kernel that walks these lists is
tested on the GPU (tests/test_gpu_parity.py::test_flow_kernel_is_bit_identical_to_one_launch_per_layer)."""
import ctypes
import pytest

from yolo_tensorflow_b200 import darknet as dn

lib = dn.lib
lib.b200_flow_schedule_selftest.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
lib.b200_flow_schedule_selftest.restype = ctypes.c_int


def darknet53_body(c0=128):
    """layers 12..80 of YOLOv3 as (size, stride, cin, cout): stride-2 3x3, then residual blocks of 1x1 + 3x3"""
    spec, c = [], c0
    for blocks in (8, 8, 4):
        spec.append((3, 2, c, 2 * c)); c *= 2
        for _ in range(blocks):
            spec.append((1, 1, c, c // 2)); spec.append((3, 1, c // 2, c))
    for _ in range(3):
        spec.append((1, 1, c, c // 2)); spec.append((3, 1, c // 2, c))
    return spec


def run(spec, batch, hw, pairs=74):
    arr = (ctypes.c_int * (4 * len(spec)))(*[v for layer in spec for v in layer])
    mk, wk = ctypes.c_double(0), ctypes.c_double(0)
    rc = lib.b200_flow_schedule_selftest(len(spec), arr, batch, hw, pairs, ctypes.byref(mk), ctypes.byref(wk))
    return rc, mk.value, wk.value


@pytest.mark.parametrize("batch,hw", [(64, 104), (32, 152), (3, 40), (1, 26), (5, 13), (64, 26)])
def test_darknet53_body_schedule_is_complete_and_deadlock_free(batch, hw):
    rc, makespan, work = run(darknet53_body(), batch, hw)
    assert rc == -1, ("stuck item", rc)
    assert makespan >= work * 0.999                      # 74 pairs cannot finish before the work divided by 74


def test_schedule_quality_at_the_headline_size():
    """list scheduling across layer borders keeps the pairs busy: at 416x416 batch 64 the simulated makespan of layers 12..80 is
    within 2 % of the work per pair (one launch per layer pays 5-25 % in tail waves alone)"""
    rc, makespan, work = run(darknet53_body(), 64, 104)
    assert rc == -1 and makespan <= 1.02 * work, (makespan, work)


@pytest.mark.parametrize("pairs", [1, 2, 7, 74])
def test_any_number_of_pairs(pairs):
    spec = [(3, 1, 64, 128), (1, 1, 128, 64), (3, 1, 64, 128), (1, 1, 128, 256), (3, 2, 256, 512), (1, 1, 512, 256), (3, 1, 256, 512)]
    rc, _, _ = run(spec, 4, 19, pairs)
    assert rc == -1


def test_odd_sizes_and_single_tile_layers():
    for batch, hw in ((1, 7), (2, 9), (7, 11), (1, 1)):
        rc, _, _ = run([(3, 1, 64, 64), (1, 1, 64, 128), (3, 1, 128, 64), (1, 1, 64, 64), (3, 1, 64, 64)], batch, hw)
        assert rc == -1, (batch, hw, rc)


lib.b200_flow_dep_selftest.argtypes = [ctypes.c_int] * 5
lib.b200_flow_dep_selftest.restype = ctypes.c_int


@pytest.mark.parametrize("size,stride", [(1, 1), (3, 1), (3, 2), (5, 1)])
def test_dependency_ranges_cover_the_taps(size, stride):
    """the counters a tile waits on before loading its operands cover every input pixel any of its taps reads (brute force over
    all tiles), for maps that are multiples of the tile, smaller than a tile, odd and non-square"""
    for batch, h, w in ((64, 13, 13), (8, 26, 26), (3, 52, 52), (2, 19, 19), (1, 7, 5), (5, 1, 1), (2, 40, 56), (1, 104, 104), (3, 76, 38)):
        if size == 1 and stride != 1:
            continue
        assert lib.b200_flow_dep_selftest(batch, h, w, size, stride) == -1, (batch, h, w, size, stride)
