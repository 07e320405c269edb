"""Per-CTA busy time of the weight-gradient kernel at the benchmark configuration, grouped by work-item shape:
calibrates the cost model of the static apportionment (csrc/mlp_tc.cu: apportion_items)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "all-in-one-deflicker_b200"))
from b200 import _native as N, atlas as A, synth          # noqa: E402

H, W, T, B = 432, 768, 80, 10000
data = synth.throughput_set(H, W, T, seed=0)
vid = A.DeviceVideo.from_reference_layout(data, "cuda")
tr = A.AtlasTrainer(vid, {"samples_batch": B}, precision=N.PREC_TC, device="cuda")
torch.manual_seed(0); tr.init_like_reference()
inds = torch.randint(H * W * T, (B,), generator=torch.Generator().manual_seed(1)).cuda()
out = {}
for wg in (True, False):
    for _ in range(3):
        tr.indices.copy_(inds); tr.loss_grad(wg)
    torch.cuda.synchronize()
    cyc = (C.c_longlong * 256)(); shp = (C.c_int32 * 768)()
    n = N.lib().b200_debug_wgrad(cyc, shp, 256)
    groups = {}
    for i in range(n):
        groups.setdefault((shp[3 * i], shp[3 * i + 1], shp[3 * i + 2]), []).append(cyc[i])
    out["with_global" if wg else "without"] = {f"{k[0]}x{k[1]} split {k[2]}": [int(np.mean(v)), int(np.max(v)), len(v)] for k, v in sorted(groups.items())}
print(json.dumps(out, indent=1))
