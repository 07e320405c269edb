"""End-to-end quality evidence on the synthetic 'quality set' (SURVEY.md §8d): pre-train + N loop trips with the
tensor-core path and with the fp32 CUDA-core path from identical initial weights and identical index
streams, render every frame, PSNR(input, reconstruction) as evaluate.py:740-743; plus a short
side-by-side of the loss trajectory against the oracle on the CPU starting from the same state.

    python tests/perf/quality_run.py [--iters 3000] [--oracle-iters 120]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "all-in-one-deflicker_b200"))
from b200 import _native as N, atlas as A, synth          # noqa: E402
from oracle import atlas_oracle as O                         # noqa: E402


def run(prec, data, H, W, T, iters, pre_iters, seed):
    vid = A.DeviceVideo.from_reference_layout(data, "cuda")
    tr = A.AtlasTrainer(vid, {}, precision=prec, device="cuda")
    torch.manual_seed(seed)
    tr.init_like_reference()
    tr.pretrain(T, H, W, pre_iters)
    state0 = (tr.state_dict("mapping"), tr.state_dict("atlas"))
    losses = []
    t0 = time.time()
    for i in range(iters):
        inds = torch.randint(H * W * T, (10000, 1))
        l = tr.step_host(inds, i)
        if i % 50 == 0:
            losses.append(float(l[0]))
    torch.cuda.synchronize()
    dt = time.time() - t0
    ps = []
    for f in range(T):
        img = tr.render_frame(f, H, W, T)
        ps.append(A.psnr(data["frames"][:, :, :, f], img.cpu()))
    return float(np.mean(ps)), losses, iters / dt, state0, tr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3000)
    ap.add_argument("--pre-iters", type=int, default=30)
    ap.add_argument("--oracle-iters", type=int, default=100)
    args = ap.parse_args()
    H, W, T = 96, 160, 12
    data = synth.quality_set(H, W, T, seed=0)
    data = {k: v for k, v in data.items() if k != "clean"}
    out = {"video": [T, H, W], "iters": args.iters, "pretrain_sweeps": args.pre_iters}
    res = {}
    for name, prec in (("tc", N.PREC_TC), ("fp32", N.PREC_FP32)):
        psnr, losses, its, state0, tr = run(prec, data, H, W, T, args.iters, args.pre_iters, seed=0)
        res[name] = (psnr, losses)
        out[name] = {"psnr_db": psnr, "loss_first": losses[0], "loss_last": losses[-1], "host_loop_it_per_s": its}
        if name == "tc":
            keep = state0
        del tr
    out["psnr_diff_db_tc_minus_fp32"] = res["tc"][0] - res["fp32"][0]
    # ---- oracle trajectory from the same post-pretrain state and the same index stream
    if args.oracle_iters > 0:
        torch.set_num_threads(min(32, os.cpu_count() or 1))
        video = O.Video(**data)
        mp = [keep[0][f"hidden.{i}.{w}"].cpu().clone().requires_grad_(True) for i in range(6) for w in ("weight", "bias")]
        ap_ = [keep[1][f"hidden.{i}.{w}"].cpu().clone().requires_grad_(True) for i in range(8) for w in ("weight", "bias")]
        opt = O.make_optimizer(mp, ap_)
        vid = A.DeviceVideo.from_reference_layout(data, "cuda")
        tr = A.AtlasTrainer(vid, {}, precision=N.PREC_TC, device="cuda")
        tr.load_state(keep[0], keep[1])
        g = torch.Generator().manual_seed(123)
        rel = []
        for i in range(args.oracle_iters):
            inds = torch.randint(H * W * T, (10000, 1), generator=g)
            ref = O.train_iteration(video, mp, ap_, opt, inds, i)
            got = tr.step_host(inds, i)
            rel.append(abs(got[0] - ref["total"]) / abs(ref["total"]))
        out["oracle_side_by_side"] = {"iters": args.oracle_iters, "max_rel_total_loss_diff": float(np.max(rel)),
                                      "mean_rel_total_loss_diff": float(np.mean(rel)),
                                      "final_loss_oracle": ref["total"], "final_loss_b200": float(got[0])}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
