"""Pins the stage-2 I/O helpers of this repo (all-in-one-deflicker_b200/src/models/utils.py: load_image, InputPadder,
tensor2img, save_img) against the reference's src/models/utils.py and freezes tests/golden/stage2_io.npz.
Run ONLY in the build container (needs /root/reference):   python tests/golden/make_golden_stage2_io.py"""
import importlib.util
import os
import tempfile

import cv2
import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.dirname(os.path.abspath(__file__))
DIMS = [(1, 3, 70, 101), (1, 6, 1080, 1920), (2, 3, 64, 96), (1, 3, 33, 31)]


def load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_module("ref_models_utils", "/root/reference/src/models/utils.py")
    mine = load_module("our_models_utils", os.path.join(ROOT, "all-in-one-deflicker_b200/src/models/utils.py"))
    rng = np.random.RandomState(9)
    rgb = rng.randint(0, 256, (70, 101, 3)).astype(np.uint8)
    grey = rng.randint(0, 256, (70, 101)).astype(np.uint8)
    out = {"rgb": rgb, "grey": grey}
    with tempfile.TemporaryDirectory() as tmp:
        p_rgb, p_grey = os.path.join(tmp, "a.png"), os.path.join(tmp, "g.png")
        Image.fromarray(rgb).save(p_rgb)
        Image.fromarray(grey).save(p_grey)
        cases = [("rgb_plain", p_rgb, None, False), ("rgb_resize32", p_rgb, None, True), ("rgb_sized", p_rgb, (96, 64), False),
                 ("grey_sized32", p_grey, (120, 90), True)]
        for tag, path, size, resize in cases:
            a, sa = ref.load_image(path, size=size, device="cpu", resize=resize)
            b, sb = mine.load_image(path, size=size, device="cpu", resize=resize)
            assert tuple(sa) == tuple(sb) and a.dtype == b.dtype and torch.equal(a, b), tag
            out["img_" + tag], out["size_" + tag] = a.numpy(), np.array(sa)
        for d in DIMS:
            pa, pb = ref.InputPadder(d), mine.InputPadder(d)
            assert list(pa._pad) == list(pb._pad), d
            x = torch.randn(d)
            assert torch.equal(pa.pad(x)[0], pb.pad(x)[0]) and torch.equal(pb.unpad(pb.pad(x)[0]), x)
            out["pad_%dx%d" % d[-2:]] = np.array(pa._pad)
        t = torch.rand(1, 3, 40, 56) * 1.2 - 0.1                      # values outside [0, 1]: save_img clips
        ia, ib = ref.tensor2img(t), mine.tensor2img(t)
        assert np.array_equal(ia, ib)
        fa, fb = os.path.join(tmp, "ra.png"), os.path.join(tmp, "rb.png")
        ref.save_img(ia, fa)
        mine.save_img(ib, fb)
        da, db = cv2.imread(fa, cv2.IMREAD_UNCHANGED), cv2.imread(fb, cv2.IMREAD_UNCHANGED)
        assert np.array_equal(da, db)
        out["t2i_in"], out["saved_bgr"] = t.numpy(), da
    np.savez_compressed(os.path.join(OUT, "stage2_io.npz"), **out)
    print("stage-2 I/O helpers bit-identical to the reference; fixture written")


if __name__ == "__main__":
    main()
