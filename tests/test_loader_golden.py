"""The stage-1 input producer (`load_input_data_single`, flow resize, forward/backward consistency masks;
reference src/models/stage_1/unwrap_utils.py:10-38,105-163) replayed against the fixture frozen from the
reference's own function by tests/golden/make_golden_loader.py.  Bit-exact, CPU only."""
import importlib.util
import os
from pathlib import Path

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ("flows_mask", "frames", "flows_rev_mask", "mask_frames", "dx", "dy", "flows_rev", "flows")


def _loader():
    path = os.path.join(ROOT, "all-in-one-deflicker_b200", "src", "models", "stage_1", "unwrap_utils.py")
    spec = importlib.util.spec_from_file_location("our_unwrap_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _write(tmp, z):
    folder = Path(tmp) / "vid"
    folder.mkdir()
    flow_dir = Path(tmp) / "vid_flow"
    flow_dir.mkdir()
    T = sum(1 for k in z.files if k.startswith("frame"))
    names = ["%05d.png" % i for i in range(T)]
    for i in range(T):
        Image.fromarray(z[f"frame{i}"]).save(str(folder / names[i]))
    for i in range(T - 1):
        np.save(flow_dir / f"{names[i]}_{names[i + 1]}.npy", z[f"f12_{i}"])
        np.save(flow_dir / f"{names[i + 1]}_{names[i]}.npy", z[f"f21_{i}"])
    return folder, T


def test_input_producer_is_bit_exact(golden_dir, tmp_path):
    z = np.load(os.path.join(golden_dir, "loader.npz"))
    folder, T = _write(tmp_path, z)
    got = _loader().load_input_data_single(int(z["resy"]), int(z["resx"]), 200, folder, True, True, folder.parent, "vid")
    assert len(got) == 8
    for name, t in zip(NAMES, got):
        want = torch.from_numpy(z["want_" + name])
        assert t.dtype == torch.float32 and t.shape == want.shape, name
        assert torch.equal(t, want), name
    masks = got[0]
    assert masks.shape == (int(z["resy"]), int(z["resx"]), T, 1)
    assert float(masks[:, :, T - 1].abs().max()) == 0.0 and float(got[2][:, :, 0].abs().max()) == 0.0   # no partner frame
    assert set(np.unique(masks.numpy())) == {0.0, 1.0}
    # forward differences, zero in the last column / row (unwrap_utils.py:132-133)
    frames, dx, dy = got[1], got[4], got[5]
    assert torch.equal(dx[:, :-1], frames[:, 1:] - frames[:, :-1]) and float(dx[:, -1].abs().max()) == 0.0
    assert torch.equal(dy[:-1], frames[1:] - frames[:-1]) and float(dy[-1].abs().max()) == 0.0


def test_frame_cap_and_unfiltered_masks(golden_dir, tmp_path):
    z = np.load(os.path.join(golden_dir, "loader.npz"))
    folder, T = _write(tmp_path, z)
    mod = _loader()
    short = mod.load_input_data_single(int(z["resy"]), int(z["resx"]), 3, folder, True, True, folder.parent, "vid")
    assert short[1].shape[-1] == 3
    assert torch.equal(short[1], torch.from_numpy(z["want_frames"])[..., :3])
    assert torch.equal(short[7][:, :, :, :2], torch.from_numpy(z["want_flows"])[:, :, :, :2])
    # filter_optical_flow=False: every pair with a partner frame is valid (that branch of the reference raises)
    allv = mod.load_input_data_single(int(z["resy"]), int(z["resx"]), 200, folder, True, False, folder.parent, "vid")
    assert float(allv[0][:, :, :T - 1].min()) == 1.0 and float(allv[0][:, :, T - 1].max()) == 0.0


def test_seg_loader_mattes_bit_exact(golden_dir, tmp_path):
    """`load_input_data` (reference unwrap_utils.py:40-103): the seven tensors of the single-layer loader plus the
    mattes of `<vid>_seg`, resized bilinearly (the reference's INTER_NEAREST lands in cv2.resize's `dst` slot)."""
    z = np.load(os.path.join(golden_dir, "loader.npz"))
    zs = np.load(os.path.join(golden_dir, "loader_seg.npz"))
    folder, T = _write(tmp_path, z)
    seg = Path(tmp_path) / "vid_seg"
    seg.mkdir()
    for i in range(T):
        Image.fromarray(zs[f"matte{i}"]).save(str(seg / ("%05d.png" % i)))
    got = _loader().load_input_data(int(z["resy"]), int(z["resx"]), 200, folder, True, True, folder.parent, "vid")
    for name, t in zip(NAMES, got):
        want = torch.from_numpy(zs["want_mask_frames"] if name == "mask_frames" else z["want_" + name])
        assert torch.equal(t, want), name
    assert len(np.unique(got[3].numpy())) > 8
