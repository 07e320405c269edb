#!/usr/bin/env python
"""Where one RAFT frame pair at 1080p (both directions, 20 iterations, mixed precision) spends its time: CUDA-event
times of the feature encoder, the context encoder, the correlation build and the captured refinement loop."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "all-in-one-deflicker_b200"))
from b200 import nn as K  # noqa: E402
from src.models.stage_1.core.raft import RAFT  # noqa: E402

g = torch.Generator().manual_seed(0)
raft = RAFT(argparse.Namespace(small=False, mixed_precision=True)).cuda().eval()
im1 = (torch.rand(1, 3, 1080, 1920, generator=g) * 255).cuda()
im2 = (torch.rand(1, 3, 1080, 1920, generator=g) * 255).cuda()


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    a = (2 * (im1 / 255.0) - 1.0).contiguous(); b = (2 * (im2 / 255.0) - 1.0).contiguous()
    out = {}
    out["pair_ms"] = timed(lambda: raft.forward_both(im1, im2, iters=20), 3)
    with raft._autocast():
        out["fnet_both_frames_ms"] = timed(lambda: raft.fnet([a, b]))
        out["cnet_one_frame_ms"] = timed(lambda: raft.cnet(a))
        f1, f2 = raft.fnet([a, b])
    out["corr_build_ms"] = timed(lambda: K.corr_build(f1.float().contiguous(), f2.float().contiguous()))
    st = next(iter(raft._graph_state.values()))
    out["refinement_graph_20_iters_ms"] = timed(lambda: st["graph"].replay())
print(out)
