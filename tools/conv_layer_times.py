#!/usr/bin/env python
"""Per-convolution CUDA-event times of the update block, UNet and TransformNet at the aux-bench sizes, for
both convolution arithmetics.  Diagnostic only (synchronises around every call)."""
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "all-in-one-deflicker_b200"))


def main():
    from b200 import nn as K
    from src.models.network_filter import UNet
    from src.models.network_local import TransformNet
    from src.models.stage_1.core.update import BasicUpdateBlock
    dev = "cuda"
    rows = []
    orig = K.conv2d

    def timed_conv(x, w, *a, **kw):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = orig(x, w, *a, **kw)
        e1.record()
        torch.cuda.synchronize()
        cin = w.shape[1]
        macs = y.shape[0] * y.shape[2] * y.shape[3] * w.shape[0] * cin * w.shape[2] * w.shape[3]
        rows.append({"net": tag[0], "prec": K.conv_precision(), "cin": cin, "cout": w.shape[0], "k": list(w.shape[2:]),
                     "out_hw": list(y.shape[2:]), "pad": kw.get("pad_mode", "zeros"), "up": kw.get("upsample", 1),
                     "stride": kw.get("stride", 1), "ms": round(e0.elapsed_time(e1), 4),
                     "tflops": round(2 * macs / e0.elapsed_time(e1) / 1e9, 1)})
        return y

    K.conv2d = timed_conv
    tag = [""]
    g = torch.Generator().manual_seed(0)
    h8, w8 = 135, 240
    ub = BasicUpdateBlock(types.SimpleNamespace(corr_levels=4, corr_radius=4), hidden_dim=128).to(dev)
    net = torch.tanh(torch.randn(1, 128, h8, w8, generator=g)).to(dev)
    inp = torch.relu(torch.randn(1, 128, h8, w8, generator=g)).to(dev)
    flow = torch.randn(1, 2, h8, w8, generator=g).to(dev)
    corr = torch.randn(1, 324, h8, w8, generator=g).to(dev)
    unet = UNet(6, 3, 32).to(dev).eval()
    tn = TransformNet(types.SimpleNamespace(nf=32, norm="IN", model="TransformNet", blocks=5), 12, 3).to(dev).eval()
    x6 = torch.rand(1, 6, 1088, 1920, generator=g).to(dev)
    x12 = torch.rand(1, 12, 1088, 1920, generator=g).to(dev)
    for prec in sys.argv[1:] or ["tc"]:
        K.set_conv_precision(prec)
        for rep in range(2):
            rows.clear()
            tag[0] = "update"; ub(net, inp, corr, flow)
            tag[0] = "unet"; unet(x6)
            tag[0] = "tn"; tn(x12, None)
        for r in rows:
            print(json.dumps(r))
        for name in ("update", "unet", "tn"):
            print(name, prec, "conv total ms", round(sum(r["ms"] for r in rows if r["net"] == name), 3))


if __name__ == "__main__":
    main()
