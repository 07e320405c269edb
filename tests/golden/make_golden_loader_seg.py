"""Pins `load_input_data` (the segmentation variant's loader: `load_input_data_single` + the mattes of `<vid>_seg`,
reference src/models/stage_1/unwrap_utils.py:40-103) against the reference's own function and freezes the matte part
of the fixture (the other seven tensors are covered by loader.npz).

Run ONLY in the build container (needs /root/reference):   python tests/golden/make_golden_loader_seg.py"""
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_loader import NAMES, OUT, REF, ROOT, load_module, synth_inputs, write_inputs  # noqa: E402


def synth_mattes(T, H, W, seed=3):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    out = []
    for t in range(T):
        blob = (np.hypot(yy - H * 0.5, xx - W * (0.3 + 0.1 * t)) < H * 0.3).astype(np.uint8) * 255
        blob[rng.randint(0, H, 20), rng.randint(0, W, 20)] = rng.randint(0, 256, 20)     # grey speckles
        out.append(blob)
    return out


def main():
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    ref = load_module("ref_unwrap_utils", os.path.join(REF, "src/models/stage_1/unwrap_utils.py"))
    sys.path.insert(0, os.path.join(ROOT, "all-in-one-deflicker_b200"))
    mine = load_module("our_unwrap_utils", os.path.join(ROOT, "all-in-one-deflicker_b200/src/models/stage_1/unwrap_utils.py"))
    frames, flows = synth_inputs()
    mattes = synth_mattes(len(frames), frames[0].shape[0], frames[0].shape[1])
    resy, resx = 30, 44
    with tempfile.TemporaryDirectory() as tmp:
        folder = write_inputs(tmp, "vid", frames, flows)
        seg = Path(tmp) / "vid_seg"
        seg.mkdir()
        for i, m in enumerate(mattes):
            Image.fromarray(m).save(str(seg / ("%05d.png" % i)))
        want = ref.load_input_data(resy, resx, 200, folder, True, True, folder.parent, "vid")
        got = mine.load_input_data(resy, resx, 200, folder, True, True, folder.parent, "vid")
    for name, a, b in zip(NAMES, want, got):
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b), name
    vals = np.unique(want[3].numpy())
    assert len(vals) > 8, "the resized mattes should hold intermediate values (bilinear, not nearest)"
    out = {"resy": resy, "resx": resx, "want_mask_frames": want[3].numpy()}
    for i, m in enumerate(mattes):
        out[f"matte{i}"] = m
    np.savez_compressed(os.path.join(OUT, "loader_seg.npz"), **out)
    print("seg loader fixture written;", len(vals), "distinct matte values")


if __name__ == "__main__":
    main()
