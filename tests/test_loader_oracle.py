"""oracle/loader_oracle.py (the cv2.remap arithmetic a device-side input producer must reproduce) pinned
bit-exactly against OpenCV itself and against this repo's / the reference's compute_consistency.  CPU only."""
import importlib.util
import os

import cv2
import numpy as np
import pytest

from oracle import loader_oracle as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed,H,W,spread", [(0, 37, 53, 3.0), (1, 8, 8, 6.0), (2, 64, 96, 0.4), (3, 21, 200, 40.0)])
def test_remap_matches_opencv_bit_for_bit(seed, H, W, spread):
    rng = np.random.RandomState(seed)
    img = rng.randn(H, W, 2).astype(np.float32)
    mx = (np.arange(W)[None, :] + rng.randn(H, W) * spread).astype(np.float32)
    my = (np.arange(H)[:, None] + rng.randn(H, W) * spread).astype(np.float32)
    mx[0, 0], my[0, 0] = -0.5, H - 0.5                       # taps straddling the border
    mx[1, 1], my[1, 1] = 2.015625, 3.484375                  # exact 1/64 offsets: round-half-even of the 1/32 grid
    want = cv2.remap(img, np.stack([mx, my], -1), None, cv2.INTER_LINEAR)
    assert np.array_equal(L.remap_bilinear_zero(img, mx, my), want)


def test_consistency_matches_the_input_producer():
    path = os.path.join(ROOT, "all-in-one-deflicker_b200", "src", "models", "stage_1", "unwrap_utils.py")
    spec = importlib.util.spec_from_file_location("our_unwrap_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                             # itself pinned against the reference (test_loader_golden.py)
    rng = np.random.RandomState(7)
    ys, xs = np.mgrid[0:40, 0:56].astype(np.float32)
    f12 = np.stack([2.0 * np.sin(ys / 6.0), 1.5 * np.cos(xs / 8.0)], -1).astype(np.float32)
    f21 = (-f12 + rng.normal(0, 0.8, f12.shape)).astype(np.float32)
    want = mod.compute_consistency(f12, f21)
    got = L.consistency_error(f12, f21)
    assert got.dtype == want.dtype == np.float32 and np.array_equal(got, want)
    assert 0.05 < float((got < 1.0).mean()) < 0.95


def test_resize_restatement_matches_opencv_bit_for_bit():
    """cv2.resize(INTER_LINEAR) on fp32 flows for every down-scaling / identity geometry class: fractional factors,
    integer factors (4 = the scripts' --down default), exactly 2x in both dimensions (OpenCV's area-fast path), 2x
    in one dimension only, same size."""
    rng = np.random.default_rng(2)
    checked = 0
    for trial in range(120):
        h, w = int(rng.integers(8, 160)), int(rng.integers(8, 220))
        kind = trial % 4
        if kind == 0:
            nh, nw = int(rng.integers(4, h + 1)), int(rng.integers(4, w + 1))
        elif kind == 1:
            k = int(rng.integers(1, 5)); nh, nw = max(2, h // k), max(2, w // k); h, w = nh * k, nw * k
        elif kind == 2:
            nh, nw = h // 2, int(rng.integers(4, w + 1)); h = nh * 2
        else:
            nh, nw = h, w
        src = (rng.standard_normal((h, w, 2)) * 5).astype(np.float32)
        want = cv2.resize(src, (nw, nh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(L.resize_bilinear_down(src, nw, nh), want), (h, w, nh, nw)
        checked += 1
    assert checked == 120


def test_resize_flow_matches_the_input_producer():
    path = os.path.join(ROOT, "all-in-one-deflicker_b200", "src", "models", "stage_1", "unwrap_utils.py")
    spec = importlib.util.spec_from_file_location("our_unwrap_utils2", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                             # pinned against the reference (test_loader_golden.py)
    rng = np.random.default_rng(9)
    for (h, w, nh, nw) in ((36, 52, 30, 44), (360, 640, 90, 160), (90, 160, 45, 80), (40, 40, 40, 40)):
        f = (rng.standard_normal((h, w, 2)) * 4).astype(np.float32)
        assert np.array_equal(L.resize_flow(f, nh, nw), mod.resize_flow(f.copy(), nh, nw)), (h, w, nh, nw)
