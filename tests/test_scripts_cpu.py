"""Host-side behaviour of the reference-named scripts that needs no GPU: loud failures where the reference
fails loudly (missing RAFT checkpoint: raft_wrapper.py:23) and for the variant this build does not ship."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "all-in-one-deflicker_b200")


def _frames(folder, n=3):
    import cv2
    os.makedirs(folder, exist_ok=True)
    for t in range(n):
        cv2.imwrite(os.path.join(folder, "%05d.png" % t), np.full((16, 24, 3), 40 * t, np.uint8))


def test_flow_prepass_refuses_random_weights(tmp_path):
    _frames(str(tmp_path / "vid"))
    env = dict(os.environ, PYTHONPATH=PKG)
    env.pop("B200_ALLOW_RANDOM_RAFT", None)
    r = subprocess.run([sys.executable, os.path.join(PKG, "src", "preprocess_optical_flow.py"), "--vid-path",
                        str(tmp_path / "vid")], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "raft-things.pth is missing" in r.stderr
    assert not any(f.endswith(".npy") for f in os.listdir(str(tmp_path / "vid_flow")))


def test_driver_moves_frames_and_runs_seg_variant(tmp_path):
    _frames(str(tmp_path / "clip"))
    env = dict(os.environ)
    env.pop("B200_ALLOW_RANDOM_RAFT", None)
    r = subprocess.run([sys.executable, os.path.join(PKG, "test.py"), "--video_frame_folder", "clip", "--class_name",
                        "person"], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    # the frame folder is moved to ./data/test/<name> as the reference does (test.py:24-31) ...
    assert os.path.isdir(str(tmp_path / "data" / "test" / "clip")) and not os.path.exists(str(tmp_path / "clip"))
    # ... and --class_name starts the segmentation variant (reference test.py:39), which stops loudly at the first
    # missing input (here the RAFT checkpoint of its flow pre-pass), never silently falling back to one layer
    assert r.returncode != 0 and "stage1_neural_atlas_seg.py" in (r.stdout + r.stderr)
    assert "raft-things.pth is missing" in r.stderr or "optical-flow pre-pass failed" in r.stderr


def test_evaluation_videos_are_written_with_opencv(tmp_path):
    """Host side of the evaluation artefacts (reference evaluate.py:714-779, unwrap_utils.py:200-231): the mp4 writers
    and the dashboard composition, no GPU involved."""
    import sys as _sys
    import cv2
    import torch
    _sys.path.insert(0, PKG)
    from src.models.stage_1.evaluate import ArtefactWriter
    from src.models.stage_1.unwrap_utils import save_mask_flow
    H, W, T = 32, 48, 4
    g = torch.Generator().manual_seed(0)
    frames = torch.full((H, W, 3, T), 0.5) + 0.02 * torch.rand(H, W, 3, T, generator=g)
    masks = torch.ones(H, W, T, 1)
    masks[:, :16] = 0                                        # a block of pixels whose flow failed the check
    masks[:, :, T - 1] = 0                                   # the last frame has no forward flow: skipped (:204-205)
    save_mask_flow(masks, frames, tmp_path)
    cap = cv2.VideoCapture(str(tmp_path / "filter_flow_0.mp4"))
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == T - 1
    ok, fr = cap.read()
    bad = masks[:, :, 0, 0].numpy() == 0
    assert ok and fr[bad][:, 2].mean() > 200 and fr[bad][:, 1].mean() < 60     # invalid pixels are red (BGR order)
    assert abs(float(fr[~bad].mean()) - 130) < 12                              # the others keep the frame's grey
    assert int(cv2.VideoCapture(str(tmp_path / "input_video.mp4")).get(cv2.CAP_PROP_FRAME_COUNT)) == T
    art = ArtefactWriter(str(tmp_path), "clip", W, H)
    for t in range(T):
        f = frames[:, :, :, t].numpy()
        art.add(f, f * 0.9, np.zeros((H, W, 2), np.float32), np.full((H, W), 3.0, np.float32), np.zeros((H, W), np.float32))
    art.close()
    for name in ("reconstruction", "residuals", "uv_1", "global_info"):
        cap = cv2.VideoCapture(str(tmp_path / f"{name}_clip.mp4"))
        assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == T, name
    ok, dash = cv2.VideoCapture(str(tmp_path / "global_info_clip.mp4")).read()
    assert ok and dash.shape == (2 * (H + 22), 3 * W, 3)


def test_fused_trainer_rejects_other_architectures():
    """A config whose network shapes differ from the ones the fused step is specialised to must not train silently
    (the reference honours these keys, src/stage1_neural_atlas.py:112-128)."""
    import sys as _sys
    import json
    import pytest
    _sys.path.insert(0, PKG)
    from b200 import _native as N, atlas as A
    cfg = json.load(open(os.path.join(PKG, "src", "config", "config_flow_100.json")))
    A.check_architecture(cfg)                                   # the shipped config is the supported one
    for key, val in (("number_of_layers_mapping1", 4), ("positional_encoding_num_atlas", 6), ("number_of_channels_atlas", 128)):
        with pytest.raises(N.B200Error):
            A.check_architecture(dict(cfg, **{key: val}))
