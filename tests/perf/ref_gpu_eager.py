"""R-GPU row of BASELINE.md: the reference's own PyTorch path on the GPU — the oracle restatement (same torch
ops as the reference, bit-identical on CPU) with the networks on cuda and the video tensors left on the
CPU exactly as the reference keeps them (gathers on the host, ~10 H2D copies per iteration).
Baseline measurement only; not part of the product.

    python tests/perf/ref_gpu_eager.py [--iters 100]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "all-in-one-deflicker_b200"))
from b200 import synth                                   # noqa: E402
from oracle import atlas_oracle as O                     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--threads", type=int, default=32)
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    H, W, T, B = 432, 768, 80, 10000
    data = synth.throughput_set(H, W, T, seed=0)
    video = O.Video(**data)
    torch.manual_seed(0)
    mp = [p.cuda().requires_grad_(True) for p in O.init_mlp(O.MAPPING_SPEC)]
    ap_ = [p.cuda().requires_grad_(True) for p in O.init_mlp(O.ATLAS_SPEC)]
    opt = O.make_optimizer(mp, ap_)
    g = torch.Generator().manual_seed(1)
    out = {}
    for name, it0 in (("with_global", 0), ("without_global", 6000)):
        for i in range(10):
            O.train_iteration(video, mp, ap_, opt, torch.randint(H * W * T, (B, 1), generator=g), it0, device="cuda")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.iters):
            O.train_iteration(video, mp, ap_, opt, torch.randint(H * W * T, (B, 1), generator=g), it0, device="cuda")
        torch.cuda.synchronize()
        out[name] = args.iters / (time.perf_counter() - t0)
    out["mean_it_per_s"] = 2.0 / (1.0 / out["with_global"] + 1.0 / out["without_global"])
    out["host_threads"] = args.threads
    print(json.dumps(out))


if __name__ == "__main__":
    main()
