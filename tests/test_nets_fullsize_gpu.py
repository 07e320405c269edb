"""Parity of the RAFT / stage-2 operators AT THE SIZES bench.py --workload raft|stage2 times (BASELINE.json configs
[3]/[4]: 1080p; RAFT feature grid 135 x 240, stage 2 at 1088 x 1920), against the CPU oracles.  The small-grid tests
(tests/test_nets_gpu.py) pin the arithmetic; these pin the tiling / indexing of the full-size launches
(32 400-pixel correlation rows, 4.2 GB pyramid, 2 M-pixel convolutions).

The 4.2 GB correlation volume is not rebuilt on the CPU: a random subset of query pixels is compared (their full
135 x 240 rows at level 0 and the three pooled levels), and the windowed lookup is checked for a subset of pixels
against the oracle's lookup on those rows.  Tolerances are those of the small-grid tests: correlation 1e-4 * max,
lookup 2e-4 * max; update block 1e-4 * max (fp32 kernels) / 4e-3 * max (tensor-core convolutions, measured 3.6e-4);
UNet / TransformNet with tensor-core convolutions 5e-3 * max|oracle output| (measured 1.1e-4 / 9.6e-4).
"""
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from nets_common import seeded_weights
from oracle import flow_oracle as FO
from oracle import stage2_oracle as SO

pytestmark = pytest.mark.gpu
DEV = "cuda"
H8, W8 = 135, 240


def _level_shapes(h, w, levels=4):
    out = []
    for _ in range(levels):
        out.append((h, w))
        h, w = h // 2, w // 2
    return out


def test_correlation_pyramid_and_lookup_1080p():
    from b200 import nn as K
    g = torch.Generator().manual_seed(21)
    f1 = torch.randn(1, 256, H8, W8, generator=g)
    f2 = torch.randn(1, 256, H8, W8, generator=g)
    hw = H8 * W8
    pyr = K.corr_build(f1.to(DEV), f2.to(DEV), impl="tc")
    shapes = _level_shapes(H8, W8)
    assert pyr.numel() == sum(hw * h * w for h, w in shapes)
    q = torch.randperm(hw, generator=g)[:384].sort().values            # query pixels, incl. first / last rows
    q[0], q[-1] = 0, hw - 1
    rows0 = torch.matmul(f1.view(256, hw)[:, q].t(), f2.view(256, hw)) / torch.sqrt(torch.tensor(256).float())
    want = [rows0.view(-1, 1, H8, W8)]
    for _ in range(3):
        want.append(F.avg_pool2d(want[-1], 2, stride=2))
    off = 0
    got_rows = []
    for (h, w), ref in zip(shapes, want):
        lvl = pyr[off: off + hw * h * w].view(hw, 1, h, w)
        got = lvl[q.to(DEV)].cpu()
        got_rows.append(got)
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        assert got.shape == ref.shape and err <= 1e-4, (h, w, err)
        off += hw * h * w
    # windowed lookup of the same pixels: coordinates = pixel position + a flow of a few pixels, some outside the map
    ys, xs = torch.meshgrid(torch.arange(H8).float(), torch.arange(W8).float(), indexing="ij")
    coords = torch.stack([xs, ys])[None] + torch.randn(1, 2, H8, W8, generator=g) * 6.0
    look = K.corr_lookup(pyr, coords.to(DEV).contiguous(), 4).cpu()              # (1, 324, 135, 240)
    sub_coords = coords.view(1, 2, hw)[:, :, q].reshape(1, 2, -1, 1)
    ref = FO.corr_lookup(want, sub_coords)                                       # the oracle on the oracle's rows
    got = look.view(1, 324, hw)[:, :, q].reshape(1, 324, -1, 1)
    assert (got - ref).abs().max() <= 2e-4 * ref.abs().max()
    del pyr
    torch.cuda.empty_cache()


def test_update_block_1080p_tensor_cores(golden_dir):
    import os
    from b200 import nn as K
    from src.models.stage_1.core.update import BasicUpdateBlock
    fx = torch.load(os.path.join(golden_dir, "raft_update.pt"))
    sd = seeded_weights(fx["shapes"], fx["seed"])
    ub = BasicUpdateBlock(types.SimpleNamespace(corr_levels=4, corr_radius=4), hidden_dim=128)
    ub.load_state_dict(sd)
    ub = ub.to(DEV)
    g = torch.Generator().manual_seed(22)
    net = torch.tanh(torch.randn(1, 128, H8, W8, generator=g))
    inp = torch.relu(torch.randn(1, 128, H8, W8, generator=g))
    corr = torch.randn(1, 324, H8, W8, generator=g)
    flow = torch.randn(1, 2, H8, W8, generator=g) * 3
    o_net, o_mask, o_delta = FO.update_block(sd, net, inp, corr, flow)
    for prec, tol in (("fp32", 1e-4), ("tc", 4e-3)):      # measured 1.2e-6 / 3.6e-4
        prev = K.set_conv_precision(prec)
        try:
            n2, m2, d2 = ub(net.to(DEV), inp.to(DEV), corr.to(DEV), flow.to(DEV))
        finally:
            K.set_conv_precision(prev)
        errs = {"net": ((n2.cpu() - o_net).abs().max() / o_net.abs().max()).item(),
                "mask": ((m2.cpu() - o_mask).abs().max() / o_mask.abs().max()).item(),
                "delta": ((d2.cpu() - o_delta).abs().max() / o_delta.abs().max()).item()}
        print(f"update block 135x240 [{prec}]:", errs)
        assert max(errs.values()) <= tol, (prec, errs)
    up = K.convex_upsample(flow.to(DEV), m2)
    assert up.shape == (1, 2, 1080, 1920)
    assert (up.cpu() - FO.convex_upsample(flow, m2.cpu())).abs().max() <= 5e-4 * (8 * flow.abs().max())


def test_stage2_networks_1080p_tensor_cores(golden_dir):
    import os
    from b200 import nn as K
    from src.models.network_filter import UNet
    from src.models.network_local import TransformNet
    fx = torch.load(os.path.join(golden_dir, "stage2_nets.pt"))
    Hp, Wp = 1088, 1920
    g = torch.Generator().manual_seed(23)
    # smooth images (the networks see video frames, not white noise)
    base = F.interpolate(torch.rand(1, 12, Hp // 16, Wp // 16, generator=g), size=(Hp, Wp), mode="bilinear", align_corners=False)
    x6, x12 = base[:, :6].contiguous(), base.contiguous()
    usd = seeded_weights(fx["unet_shapes"], fx["unet_seed"])
    tsd = seeded_weights(fx["tn_shapes"], fx["tn_seed"])
    with torch.no_grad():
        oy = SO.unet_forward(usd, x6)
        ty, th, tc_ = SO.transformnet_forward(tsd, x12)
    unet = UNet(in_channels=6, out_channels=3, init_features=32)
    unet.load_state_dict(usd)
    tn = TransformNet(types.SimpleNamespace(nf=32, norm="IN", model="TransformNet", blocks=5), nc_in=12, nc_out=3)
    tn.load_state_dict(tsd, strict=False)
    prev = K.set_conv_precision("tc")
    try:
        with torch.no_grad():
            y = unet.to(DEV)(x6.to(DEV))
            yy, (hid, cell) = tn.to(DEV)(x12.to(DEV), None)
    finally:
        K.set_conv_precision(prev)
    errs = {"unet": ((y.cpu() - oy).abs().max() / oy.abs().max()).item(),
            "tn": ((yy.cpu() - ty).abs().max() / ty.abs().max()).item(),
            "tn_cell": ((cell.cpu() - tc_).abs().max() / tc_.abs().max()).item()}
    print("stage 2 at 1088x1920 [tc]:", errs)
    assert y.shape == (1, 3, Hp, Wp) and yy.shape == (1, 3, Hp, Wp)
    assert max(errs.values()) <= 5e-3, errs              # measured 1.1e-4 (UNet), 6.8e-4 / 9.6e-4 (TransformNet)


def test_raft_graph_replay_equals_eager_1080p(golden_dir):
    """The captured refinement loop (one CUDA graph per geometry) against the same loop launched eagerly, on a
    1080p pair: identical kernels on identical buffers -> bit-identical flows, also on the second replay."""
    import argparse
    import os
    from src.models.stage_1.core.raft import RAFT
    fx = torch.load(os.path.join(golden_dir, "raft_full.pt"))
    sd = seeded_weights(fx["shapes"], fx["seed"])
    g = torch.Generator().manual_seed(24)
    im1 = (torch.rand(1, 3, 1080, 1920, generator=g) * 255).to(DEV)
    im2 = (torch.rand(1, 3, 1080, 1920, generator=g) * 255).to(DEV)
    outs = {}
    for graph in (False, True):
        model = RAFT(argparse.Namespace(small=False, mixed_precision=True, cuda_graph=graph))
        model.load_state_dict(sd, strict=False)
        model = model.to(DEV).eval()
        low, up = model(im1, im2, iters=4, test_mode=True)
        if graph:
            low2, up2 = model(im1, im2, iters=4, test_mode=True)         # second call: pure replay
            assert torch.equal(up, up2) and torch.equal(low, low2)
        outs[graph] = (low.clone(), up.clone())
        del model
        torch.cuda.empty_cache()
    assert outs[True][1].shape == (1, 2, 1080, 1920) and torch.isfinite(outs[True][1]).all()
    assert torch.equal(outs[False][0], outs[True][0]) and torch.equal(outs[False][1], outs[True][1])
