"""RAFTWrapper image loading (decode, long-edge INTER_AREA downsampling, sorted pair, /8 replicate padding;
reference src/models/stage_1/raft_wrapper.py:29-63) replayed against the fixture frozen from the reference by
tests/golden/make_golden_raft_io.py.  Bit-exact, CPU only (no model is constructed)."""
import os
import types

import numpy as np
import torch
from PIL import Image


def test_load_images_bit_exact(golden_dir, tmp_path):
    import src.models.stage_1.raft_wrapper as W
    z = np.load(os.path.join(golden_dir, "raft_io.npz"))
    Image.fromarray(z["a"]).save(str(tmp_path / "a.png"))
    Image.fromarray(z["b"]).save(str(tmp_path / "b.png"))
    saved, W.device = W.device, torch.device("cpu")
    try:
        for edge in (100, 2000):
            w = W.RAFTWrapper.__new__(W.RAFTWrapper)
            w.args = types.SimpleNamespace(max_long_edge=edge)
            im1, im2 = w.load_images(str(tmp_path / "b.png"), str(tmp_path / "a.png"))
            assert torch.equal(im1, torch.from_numpy(z["want_im1_%d" % edge]))
            assert torch.equal(im2, torch.from_numpy(z["want_im2_%d" % edge]))
            assert torch.equal(w.load_image(str(tmp_path / "a.png")), torch.from_numpy(z["want_single_%d" % edge]))
            assert im1.shape[-1] % 8 == 0 and im1.shape[-2] % 8 == 0
            assert torch.equal(w.load_image_list([str(tmp_path / "b.png"), str(tmp_path / "a.png")])[0:1], im1)
    finally:
        W.device = saved
