"""Host half of vr_scene_commit that needs no GPU: the reference-order leaf sequence behind the tie ranks
(csrc/scene_build.cpp `reference_leaf_order`, core/bvh.rs:48-130) against a direct numpy restatement, on the
serial path and on the multi-threaded top-of-tree path (forced by VOIDRAY_PAR_MIN)."""
import ctypes as C
import os

import numpy as np
import pytest

from voidray_b200 import _lib

F32 = np.float32


def _total_key(x: np.ndarray) -> np.ndarray:
    i = x.view(np.int32).astype(np.int64)
    return np.where(i < 0, i ^ 0x7FFFFFFF, i)  # f32::total_cmp order


def _naive_order(cen: np.ndarray) -> np.ndarray:
    order = np.arange(len(cen))
    stack = [(0, len(cen))]
    while stack:
        lo, hi = stack.pop()
        if hi - lo < 2:
            continue
        c = cen[order[lo:hi]]
        spread = c.max(0) - c.min(0)
        if spread[0] > spread[1] and spread[0] > spread[2]:
            axis = 0
        elif spread[1] > spread[0] and spread[1] > spread[2]:
            axis = 1
        else:
            axis = 2
        order[lo:hi] = order[lo:hi][np.argsort(_total_key(c[:, axis]), kind="stable")]
        mid = lo + (hi - lo) // 2
        stack += [(lo, mid), (mid, hi)]
    return order


def _native_order(boxes: np.ndarray, par_min=None) -> np.ndarray:
    lib = _lib.load()
    boxes = np.ascontiguousarray(boxes, F32)
    out = np.empty(len(boxes), np.uint32)
    old = os.environ.get("VOIDRAY_PAR_MIN")
    if par_min is not None:
        os.environ["VOIDRAY_PAR_MIN"] = str(par_min)
    try:
        _lib.check(lib.vr_debug_reference_leaf_order(_lib.fptr(boxes), len(boxes), out.ctypes.data_as(C.POINTER(C.c_uint32))))
    finally:
        if par_min is not None:
            if old is None:
                del os.environ["VOIDRAY_PAR_MIN"]
            else:
                os.environ["VOIDRAY_PAR_MIN"] = old
    return out


def _boxes(n, mode, seed):
    rng = np.random.default_rng(seed)
    if mode == "random":
        c = rng.uniform(-5, 5, (n, 3)).astype(F32)
        h = rng.uniform(0, 0.3, (n, 3)).astype(F32)
    elif mode == "ties":  # a handful of distinct coordinates, signed zeros: stability decides almost everything
        c = (rng.integers(0, 7, (n, 3)) - 3).astype(F32)
        c[rng.random((n, 3)) < 0.02] = F32(-0.0)
        h = np.full((n, 3), 0.5, F32)
    else:  # lattice
        c = (rng.integers(0, 1000, (n, 3)) * 0.25 - 100.0).astype(F32)
        h = rng.uniform(0, 0.3, (n, 3)).astype(F32)
    return np.concatenate([c - h, c + h], axis=1).astype(F32)


@pytest.mark.parametrize("mode", ["random", "ties", "lattice"])
@pytest.mark.parametrize("n", [1, 2, 3, 25, 1000, 40000])
def test_leaf_order_matches_direct_restatement(n, mode):
    boxes = _boxes(n, mode, n)
    cen = ((boxes[:, :3] + boxes[:, 3:]) / F32(2.0)).astype(F32)
    want = _naive_order(cen)
    assert np.array_equal(_native_order(boxes), want)
    assert np.array_equal(_native_order(boxes, par_min=64), want)


@pytest.mark.parametrize("mode", ["random", "ties"])
def test_leaf_order_parallel_top_equals_serial(mode):
    boxes = _boxes(400000, mode, 11)
    serial = _native_order(boxes, par_min=10**9)
    assert np.array_equal(_native_order(boxes, par_min=1 << 16), serial)
    assert sorted(serial.tolist()) == list(range(400000))
