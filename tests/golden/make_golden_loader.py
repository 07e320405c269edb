"""Pins the stage-1 input producer of this repo (all-in-one-deflicker_b200/src/models/stage_1/unwrap_utils.py:
load_input_data_single, compute_consistency, resize_flow) against the reference's own functions
(src/models/stage_1/unwrap_utils.py:10-38,105-163) and freezes a small fixture.

Run ONLY in the build container (needs /root/reference):   python tests/golden/make_golden_loader.py
The fixture (tests/golden/loader.npz) holds the input frames / flow files and the reference's eight output tensors;
tests/test_loader_golden.py replays it anywhere."""
import importlib.util
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
NAMES = ("flows_mask", "frames", "flows_rev_mask", "mask_frames", "dx", "dy", "flows_rev", "flows")


def load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synth_inputs(seed=5, T=4, H=36, W=52):
    rng = np.random.RandomState(seed)
    frames = [rng.randint(0, 256, (H, W, 3)).astype(np.uint8) for _ in range(T)]
    frames[2] = frames[2][:, :, 0]                                   # one greyscale frame (tiled to 3 channels)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    flows = []
    for t in range(T - 1):
        f12 = np.stack([1.5 * np.sin(ys / 7.0 + t), 1.2 * np.cos(xs / 9.0 - t)], -1).astype(np.float32)
        f21 = (-f12 + rng.normal(0, 0.7, f12.shape)).astype(np.float32)  # partly consistent: masks hold 0s and 1s
        flows.append((f12, f21))
    return frames, flows


def write_inputs(root, vid, frames, flows):
    folder = Path(root) / vid
    folder.mkdir(parents=True)
    flow_dir = Path(root) / (vid + "_flow")
    flow_dir.mkdir()
    names = []
    for i, fr in enumerate(frames):
        name = "%05d.png" % i
        Image.fromarray(fr).save(str(folder / name))
        names.append(name)
    for i, (f12, f21) in enumerate(flows):
        np.save(flow_dir / f"{names[i]}_{names[i + 1]}.npy", f12)
        np.save(flow_dir / f"{names[i + 1]}_{names[i]}.npy", f21)
    return folder


def main():
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    ref = load_module("ref_unwrap_utils", os.path.join(REF, "src/models/stage_1/unwrap_utils.py"))
    sys.path.insert(0, os.path.join(ROOT, "all-in-one-deflicker_b200"))
    mine = load_module("our_unwrap_utils", os.path.join(ROOT, "all-in-one-deflicker_b200/src/models/stage_1/unwrap_utils.py"))
    frames, flows = synth_inputs()
    resy, resx = 30, 44                                              # != flow size: exercises resize_flow
    with tempfile.TemporaryDirectory() as tmp:
        folder = write_inputs(tmp, "vid", frames, flows)
        want = ref.load_input_data_single(resy, resx, 200, folder, True, True, folder.parent, "vid")
        got = mine.load_input_data_single(resy, resx, 200, folder, True, True, folder.parent, "vid")
        # maximum_number_of_frames below the number of files (filter_optical_flow=False is not compared: that branch of
        # the reference raises, torch.ones_like on a numpy array, unwrap_utils.py:161 — the scripts always pass True)
        want_short = ref.load_input_data_single(resy, resx, 3, folder, True, True, folder.parent, "vid")
        got_short = mine.load_input_data_single(resy, resx, 3, folder, True, True, folder.parent, "vid")
    for name, a, b in zip(NAMES, want, got):
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b), name
    for name, a, b in zip(NAMES, want_short, got_short):
        assert a.shape == b.shape and torch.equal(a, b), name + " (3 of 4 frames)"
    assert 0.05 < float(want[0].mean()) < 0.95, "the consistency masks should hold both values"
    out = {"resy": resy, "resx": resx}
    for i, fr in enumerate(frames):
        out[f"frame{i}"] = fr
    for i, (f12, f21) in enumerate(flows):
        out[f"f12_{i}"], out[f"f21_{i}"] = f12, f21
    for name, a in zip(NAMES, want):
        out["want_" + name] = a.numpy()
    np.savez_compressed(os.path.join(OUT, "loader.npz"), **out)
    print("loader fixture written; this repo's input producer is bit-identical to the reference's on it;",
          "mask mean", float(want[0].mean()))


if __name__ == "__main__":
    main()
