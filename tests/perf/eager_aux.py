"""torch-eager (cuBLAS / cuDNN) timings of the oracle restatements, for orientation next to bench_aux.py's numbers.
Lives under tests/ because it executes oracle/ code (only tests, smoke() and bench.py's CPU baseline may)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import flow_oracle as FO      # noqa: E402
from oracle import stage2_oracle as SO    # noqa: E402


def corr(timed, f1, f2, coords):
    try:
        ms_build = timed(lambda: FO.corr_pyramid(f1, f2), iters=2)
        rp = FO.corr_pyramid(f1, f2)
        ms_look = timed(lambda: FO.corr_lookup(rp, coords), iters=3)
        return {"build_ms": ms_build, "lookup_ms": ms_look}
    except Exception as e:      # noqa: BLE001  (out of memory on small boxes)
        return {"error": str(e)[:100]}


def update_block(timed, ub, net, inp, corr, flow):
    sd = {k: v.detach() for k, v in ub.state_dict().items()}
    with torch.no_grad():
        return timed(lambda: FO.update_block(sd, net, inp, corr, flow), iters=3)


def stage2(timed, unet, tn, x6, x12):
    usd = {k: v.detach() for k, v in unet.state_dict().items()}
    tsd = {k: v.detach() for k, v in tn.state_dict().items()}
    with torch.no_grad():
        return {"unet_ms": timed(lambda: SO.unet_forward(usd, x6), iters=2),
                "transformnet_ms": timed(lambda: SO.transformnet_forward(tsd, x12), iters=2)}
