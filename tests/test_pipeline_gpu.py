"""End-to-end run of the reference-named scripts on a tiny synthetic video (GPU): RAFT flow pre-pass ->
stage-1 neural atlas (few hundred iterations) -> stage-2 neural filter + local refinement, driven exactly
like the reference drives them (`python src/stage1_neural_atlas.py --vid_name ...`, then
`python src/neural_filter_and_refinement.py --video_name ...`; test.py:30-42).  Checks the on-disk contract:
flow files, checkpoint keys, rendered frames, PSNR marker, stage-2 outputs.  Pretrained weights do not exist
offline, so RAFT runs with its random initialisation and the stage-2 checkpoints are random-init state_dicts
written by the test (which also proves the reference's checkpoint keys load)."""
import glob
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "all-in-one-deflicker_b200")


def _write_video(folder, T=6, H=128, W=192):
    import cv2
    os.makedirs(folder, exist_ok=True)
    rng = np.random.RandomState(0)
    base = cv2.GaussianBlur(rng.rand(H + 32, W + 32, 3).astype(np.float32), (0, 0), 3.0)
    base = (base - base.min()) / (base.max() - base.min())
    for t in range(T):
        crop = base[8 + t:8 + t + H, 4 + 2 * t:4 + 2 * t + W]             # 2 px/frame pan
        flick = 1.0 + 0.15 * np.sin(1.7 * t)                                # global brightness flicker
        cv2.imwrite(os.path.join(folder, "%05d.png" % t), np.clip(crop * flick * 255.0, 0, 255).astype(np.uint8))


def test_scripts_end_to_end(tmp_path):
    sys.path.insert(0, PKG)
    from src.models.network_filter import UNet
    from src.models.network_local import TransformNet
    work = tmp_path
    vid = "tiny"
    _write_video(str(work / "data" / "test" / vid))
    cfg = json.load(open(os.path.join(PKG, "src", "config", "config_flow_100.json")))
    cfg.update(iters_num=301, evaluate_every=300, pretrain_iter_number=3, samples_batch=2000, stop_global_rigidity=150)
    cfg_path = str(work / "cfg.json")
    json.dump(cfg, open(cfg_path, "w"))
    env = dict(os.environ, PYTHONPATH=PKG, B200_ALLOW_RANDOM_RAFT="1")    # no pretrained RAFT offline
    r = subprocess.run([sys.executable, os.path.join(PKG, "src", "stage1_neural_atlas.py"), "--vid_name", vid, "--root",
                        "data/test/", "--down", "1", "--config", cfg_path], cwd=str(work), env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    flows = glob.glob(str(work / "data" / "test" / (vid + "_flow") / "*.npy"))
    assert len(flows) == 10                                                # 5 consecutive pairs, both directions
    f = np.load(flows[0])
    assert f.shape == (128, 192, 2) and f.dtype == np.float32 and np.isfinite(f).all()
    res = work / "results" / vid / "stage_1"
    ck = torch.load(str(res / "checkpoint"), weights_only=False)
    assert set(ck) == {"F_atlas_state_dict", "iteration", "model_F_mapping1_state_dict", "optimizer_all_state_dict"}
    assert ck["iteration"] == 300 and len(ck["F_atlas_state_dict"]) == 16 and len(ck["model_F_mapping1_state_dict"]) == 12
    outs = sorted(glob.glob(str(res / "output" / "*.png")))
    assert len(outs) == 6
    marker = glob.glob(str(res / "000300" / "PSNR_*"))
    assert len(marker) == 1
    psnr = float(os.path.basename(marker[0])[len("PSNR_"):])
    assert np.isfinite(psnr) and psnr > 12.0, psnr                         # 300 small iterations already fit the pan roughly
    # evaluation artefacts of the reference (evaluate.py:714-793, unwrap_utils.py:200-231), written with OpenCV
    import cv2
    for name, frames in (("000300/reconstruction_tiny.mp4", 6), ("000300/residuals_tiny.mp4", 6), ("000300/uv_1_tiny.mp4", 6),
                         ("000300/global_info_tiny.mp4", 6), ("input_video.mp4", 6), ("filter_flow_0.mp4", None)):
        assert os.path.exists(str(res / name)), name
        cap = cv2.VideoCapture(str(res / name))
        if frames is None:        # frames without a single consistent flow are skipped (random-init RAFT: maybe all)
            assert not cap.isOpened() or int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) <= 6
        else:
            assert cap.isOpened() and int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == frames, name
        cap.release()
    cap = cv2.VideoCapture(str(res / "000300" / "global_info_tiny.mp4"))
    ok, dash = cap.read()
    assert ok and dash.shape == (2 * (128 + 22), 3 * 192, 3)
    assert glob.glob(str(res / "events.out.tfevents.*"))                   # tensorboard log with the two images
    # stage 2 with random-init checkpoints under the reference's keys
    os.makedirs(str(work / "pretrained_weights"), exist_ok=True)
    torch.manual_seed(0)
    torch.save(UNet(in_channels=6, out_channels=3, init_features=32).state_dict(), str(work / "pretrained_weights" / "neural_filter.pth"))
    tn = TransformNet(types.SimpleNamespace(nf=32, norm="IN", model="TransformNet", blocks=5), nc_in=12, nc_out=3)
    torch.save(tn.state_dict(), str(work / "pretrained_weights" / "local_refinement_net.pth"))
    r = subprocess.run([sys.executable, os.path.join(PKG, "src", "neural_filter_and_refinement.py"), "--video_name", vid],
                       cwd=str(work), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for sub in ("neural_filter/output", "final/output"):
        pngs = glob.glob(str(work / "results" / vid / sub / "*.png"))
        assert len(pngs) == 6, sub
    import cv2
    img = cv2.imread(sorted(glob.glob(str(work / "results" / vid / "final" / "output" / "*.png")))[-1])
    assert img.shape == (128, 192, 3)


def test_seg_script_end_to_end(tmp_path):
    """`python src/stage1_neural_atlas_seg.py --vid_name ... --class_name ...` (reference test.py:39) on a tiny clip with
    synthetic mattes: flow pre-pass, pre-training of both mappings, 201 iterations, checkpoint keys of
    evaluate.py:216-223, composite frames, alpha mattes, PSNR marker."""
    import cv2
    work = tmp_path
    vid = "tinyseg"
    T, H, W = 5, 96, 128
    _write_video(str(work / "data" / "test" / vid), T=T, H=H, W=W)
    seg_dir = str(work / "data" / "test" / (vid + "_seg"))
    os.makedirs(seg_dir)
    yy, xx = np.mgrid[0:H, 0:W]
    for t in range(T):          # the matte moves with the pan of _write_video (content shifts by (-2, -1) px per frame)
        m = (np.hypot(yy - (H * 0.5 - t), xx - (W * 0.5 - 2 * t)) < H * 0.3).astype(np.uint8) * 255
        cv2.imwrite(os.path.join(seg_dir, "%05d.png" % t), m)
    # exact flows of the pan, written where the pre-pass would put RAFT's (it skips existing pairs): with the random-init
    # RAFT of the offline box the alpha-weighted flow terms (coefficient 500) would drown the bootstrapping term
    flow_dir = str(work / "data" / "test" / (vid + "_flow"))
    os.makedirs(flow_dir)
    for t in range(T - 1):
        a, b = "%05d.png" % t, "%05d.png" % (t + 1)
        f12 = np.tile(np.array([-2.0, -1.0], np.float32), (H, W, 1))
        np.save(os.path.join(flow_dir, f"{a}_{b}.npy"), f12)
        np.save(os.path.join(flow_dir, f"{b}_{a}.npy"), -f12)
    cfg = json.load(open(os.path.join(PKG, "src", "config", "config_flow_100.json")))
    cfg.update(iters_num=201, evaluate_every=200, pretrain_iter_number=2, samples_batch=1500, stop_global_rigidity=100)
    cfg_path = str(work / "cfg.json")
    json.dump(cfg, open(cfg_path, "w"))
    env = dict(os.environ, PYTHONPATH=PKG, B200_ALLOW_RANDOM_RAFT="1")
    r = subprocess.run([sys.executable, os.path.join(PKG, "src", "stage1_neural_atlas_seg.py"), "--vid_name", vid, "--root",
                        "data/test/", "--down", "1", "--class_name", "person", "--config", cfg_path], cwd=str(work), env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = work / "results" / vid / "stage_1"
    ck = torch.load(str(res / "checkpoint"), weights_only=False)
    assert set(ck) == {"F_atlas_state_dict", "iteration", "model_F_mapping1_state_dict", "model_F_mapping2_state_dict",
                       "model_F_alpha_state_dict", "optimizer_all_state_dict"}
    assert ck["iteration"] == 200 and len(ck["model_F_mapping2_state_dict"]) == 8 and len(ck["model_F_alpha_state_dict"]) == 16
    assert ck["model_F_alpha_state_dict"]["hidden.0.weight"].shape == (256, 30)
    assert len(ck["optimizer_all_state_dict"]["param_groups"]) == 4
    assert len(glob.glob(str(res / "output" / "*.png"))) == T
    alphas = sorted(glob.glob(str(res / "000200" / "alpha" / "*.png")))
    assert len(alphas) == T
    a = cv2.imread(alphas[2], cv2.IMREAD_GRAYSCALE).astype(np.float64) / 255
    gt = cv2.imread(os.path.join(seg_dir, "00002.png"), cv2.IMREAD_GRAYSCALE) > 127
    print("seg e2e: alpha inside / outside the matte after 200 trips:", a[gt].mean(), a[~gt].mean())
    assert a[gt].mean() > a[~gt].mean() + 0.1          # 200 bootstrapped iterations already separate the matte
    marker = glob.glob(str(res / "000200" / "PSNR_*"))
    assert len(marker) == 1 and np.isfinite(float(os.path.basename(marker[0])[len("PSNR_"):]))
    # a missing matte folder is refused loudly
    os.rename(seg_dir, seg_dir + "_gone")
    r = subprocess.run([sys.executable, os.path.join(PKG, "src", "stage1_neural_atlas_seg.py"), "--vid_name", vid, "--root",
                        "data/test/", "--class_name", "person", "--config", cfg_path], cwd=str(work), env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "holds no mattes" in r.stderr
