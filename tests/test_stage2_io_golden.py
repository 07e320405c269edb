"""Stage-2 I/O helpers (load_image, InputPadder /32, tensor2img, save_img; reference src/models/utils.py:55-60,
234-247,600-644) replayed against the fixture frozen from the reference by tests/golden/make_golden_stage2_io.py.
Bit-exact, CPU only."""
import importlib.util
import os

import cv2
import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _helpers():
    path = os.path.join(ROOT, "all-in-one-deflicker_b200", "src", "models", "utils.py")
    spec = importlib.util.spec_from_file_location("our_models_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_load_image_variants(golden_dir, tmp_path):
    z = np.load(os.path.join(golden_dir, "stage2_io.npz"))
    m = _helpers()
    p_rgb, p_grey = str(tmp_path / "a.png"), str(tmp_path / "g.png")
    Image.fromarray(z["rgb"]).save(p_rgb)
    Image.fromarray(z["grey"]).save(p_grey)
    for tag, path, size, resize in [("rgb_plain", p_rgb, None, False), ("rgb_resize32", p_rgb, None, True),
                                    ("rgb_sized", p_rgb, (96, 64), False), ("grey_sized32", p_grey, (120, 90), True)]:
        img, org = m.load_image(path, size=size, device="cpu", resize=resize)
        assert img.dtype == torch.float32 and tuple(org) == tuple(z["size_" + tag]), tag
        assert torch.equal(img, torch.from_numpy(z["img_" + tag])), tag
        if resize:
            assert img.shape[-1] % 32 == 0 and img.shape[-2] % 32 == 0


def test_padder_and_image_round_trip(golden_dir, tmp_path):
    z = np.load(os.path.join(golden_dir, "stage2_io.npz"))
    m = _helpers()
    for key in [k for k in z.files if k.startswith("pad_")]:
        h, w = (int(v) for v in key[4:].split("x"))
        p = m.InputPadder((1, 3, h, w))
        assert list(p._pad) == list(z[key]), key
        x = torch.randn(1, 3, h, w)
        y, = p.pad(x)
        assert y.shape[-1] % 32 == 0 and y.shape[-2] % 32 == 0 and torch.equal(p.unpad(y), x)
    t = torch.from_numpy(z["t2i_in"])
    out = str(tmp_path / "o.png")
    m.save_img(m.tensor2img(t), out)
    assert np.array_equal(cv2.imread(out, cv2.IMREAD_UNCHANGED), z["saved_bgr"])
