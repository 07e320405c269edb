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
