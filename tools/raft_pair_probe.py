#!/usr/bin/env python
"""One RAFT frame pair at 1080p (20 iterations, mixed precision) — for launch lists."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "all-in-one-deflicker_b200"))
from src.models.stage_1.core.raft import RAFT  # noqa: E402

g = torch.Generator().manual_seed(0)
raft = RAFT(argparse.Namespace(small=False, mixed_precision=True)).cuda().eval()
im1 = (torch.rand(1, 3, 1080, 1920, generator=g) * 255).cuda()
im2 = (torch.rand(1, 3, 1080, 1920, generator=g) * 255).cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    raft(im1, im2, iters=20, test_mode=True)
torch.cuda.synchronize()
