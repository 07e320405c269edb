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
