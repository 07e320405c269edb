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


def test_driver_moves_frames_and_rejects_unbuilt_variant(tmp_path):
    _frames(str(tmp_path / "clip"))
    seg = os.path.join(PKG, "src", "stage1_neural_atlas_seg.py")
    r = subprocess.run([sys.executable, os.path.join(PKG, "test.py"), "--video_frame_folder", "clip", "--class_name",
                        "person"], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    # the frame folder is moved to ./data/test/<name> as the reference does (test.py:24-31) ...
    assert os.path.isdir(str(tmp_path / "data" / "test" / "clip")) and not os.path.exists(str(tmp_path / "clip"))
    if not os.path.exists(seg):          # ... and the segmentation variant is refused loudly, never silently ignored
        assert r.returncode != 0 and "NotImplementedError" in r.stderr
