#!/usr/bin/env python
"""Secondary benchmarks of BASELINE.json configs[3] and configs[4] (not the headline metric; bench.py is):

  config 4  RAFT at 1080p (H/8 x W/8 = 135 x 240): all-pairs correlation volume + pyramid, windowed lookup,
            one update-block iteration, and a whole frame pair (20 iterations, both kernels + encoders)
  config 5  stage 2 at 1088 x 1920: UNet neural filter and TransformNet local refinement, frames/s

Each number is CUDA-event time of OUR kernels through the C ABI.  With --with-eager the same operators are also
timed as plain torch ops (cuBLAS/cuDNN) on the same GPU for orientation: that leg runs the oracle restatement
and therefore lives under tests/ (tests/perf/eager_aux.py).  One JSON line.
    python bench_aux.py [--small] [--conv tc|fp32] [--with-eager]
"""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "all-in-one-deflicker_b200"))


def timed(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true", help="quarter resolution (quick check)")
    ap.add_argument("--conv", choices=["fp32", "tc"], default="tc", help="convolution arithmetic (b200.nn)")
    ap.add_argument("--with-eager", action="store_true", help="also time torch-eager restatements (tests/perf/eager_aux.py)")
    args = ap.parse_args()
    from b200 import nn as K
    K.set_conv_precision(args.conv)
    eager = None
    if args.with_eager:
        sys.path.insert(0, os.path.join(ROOT, "tests", "perf"))
        import eager_aux as eager
    from src.models.network_filter import UNet
    from src.models.network_local import TransformNet
    from src.models.stage_1.core.raft import RAFT
    from src.models.stage_1.core.update import BasicUpdateBlock
    dev = "cuda"
    H, W = (272, 480) if args.small else (1080, 1920)
    h8, w8 = (H + 7) // 8, (W + 7) // 8
    out = {"config": {"raft_frame": [H, W], "h8_w8": [h8, w8], "conv": args.conv}}
    g = torch.Generator(device="cpu").manual_seed(0)
    # ---------------- RAFT correlation
    f1 = torch.randn(1, 256, h8, w8, generator=g).to(dev)
    f2 = torch.randn(1, 256, h8, w8, generator=g).to(dev)
    hw = h8 * w8
    pyr = [None]
    def build():
        pyr[0] = None                      # release the previous 5.6 GB pyramid first: no allocator growth in the timed region
        pyr[0] = K.corr_build(f1, f2)
    ms = timed(build, iters=2)
    vol_bytes = 4.0 * hw * hw * (1 + 0.25 + 0.0625 + 0.015625)
    out["corr_build"] = {"ms": ms, "tflops": 2.0 * hw * hw * 256 / ms / 1e9, "gb_written": vol_bytes / 1e9,
                         "gbs": vol_bytes / ms / 1e6}
    ys, xs = torch.meshgrid(torch.arange(h8).float(), torch.arange(w8).float(), indexing="ij")
    coords = (torch.stack([xs, ys])[None] + torch.randn(1, 2, h8, w8, generator=g)).to(dev)
    ms = timed(lambda: K.corr_lookup(pyr[0], coords), iters=5)
    out["corr_lookup"] = {"ms": ms, "taps_per_s": hw * 324 * 4 / ms * 1e3}
    if eager:
        out["torch_eager_corr"] = eager.corr(timed, f1, f2, coords)
    # ---------------- update block, one iteration
    ub = BasicUpdateBlock(types.SimpleNamespace(corr_levels=4, corr_radius=4), hidden_dim=128).to(dev)
    net = torch.tanh(torch.randn(1, 128, h8, w8, generator=g)).to(dev)
    inp = torch.relu(torch.randn(1, 128, h8, w8, generator=g)).to(dev)
    flow = torch.randn(1, 2, h8, w8, generator=g).to(dev)
    corr = K.corr_lookup(pyr[0], coords)
    ms = timed(lambda: ub(net, inp, corr, flow), iters=3)
    out["update_block_iter"] = {"ms": ms, "tflops": 2 * 3.118e6 * hw / ms / 1e9}
    if eager:
        out["torch_eager_update_block_iter_ms"] = eager.update_block(timed, ub, net, inp, corr, flow)
    del pyr, corr
    torch.cuda.empty_cache()
    # ---------------- whole pair
    import argparse as ap2
    raft = RAFT(ap2.Namespace(small=False, mixed_precision=True)).to(dev).eval()
    im1 = (torch.rand(1, 3, H // 8 * 8, W // 8 * 8, generator=g) * 255).to(dev)
    im2 = (torch.rand(1, 3, H // 8 * 8, W // 8 * 8, generator=g) * 255).to(dev)
    ms = timed(lambda: raft(im1, im2, iters=20, test_mode=True), iters=1, warm=1)
    out["raft_pair_20iters"] = {"ms": ms, "pairs_per_s_both_directions": 1000.0 / (2 * ms)}
    del raft
    torch.cuda.empty_cache()
    # ---------------- stage 2
    Hp, Wp = (288, 480) if args.small else (1088, 1920)
    unet = UNet(6, 3, 32).to(dev).eval()
    tn = TransformNet(types.SimpleNamespace(nf=32, norm="IN", model="TransformNet", blocks=5), 12, 3).to(dev).eval()
    x6 = torch.rand(1, 6, Hp, Wp, generator=g).to(dev)
    x12 = torch.rand(1, 12, Hp, Wp, generator=g).to(dev)
    ms_u = timed(lambda: unet(x6), iters=2)
    ms_t = timed(lambda: tn(x12, None), iters=2)
    out["stage2"] = {"unet_ms": ms_u, "transformnet_ms": ms_t, "frames_per_s": 1000.0 / (ms_u + ms_t),
                     "unet_tflops": 2 * 524e9 * (Hp * Wp) / (1088 * 1920) / ms_u / 1e9,
                     "transformnet_tflops": 2 * 559e9 * (Hp * Wp) / (1088 * 1920) / ms_t / 1e9}
    if eager:
        out["torch_eager_stage2"] = eager.stage2(timed, unet, tn, x6, x12)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
