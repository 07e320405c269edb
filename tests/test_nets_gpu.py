"""RAFT correlation / update block and the stage-2 networks through the C ABI against the oracles
(fp32 CUDA-core kernels; tolerance 1e-4 absolute on O(1) outputs, correlation volume 1e-4 relative)."""
import os
import types

import numpy as np
import pytest
import torch

from nets_common import seeded_weights
from oracle import flow_oracle as FO
from oracle import stage2_oracle as SO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_corr_build_and_lookup(golden_dir):
    from b200 import nn as K
    from src.models.stage_1.core.corr import CorrBlock
    z = np.load(os.path.join(golden_dir, "raft_corr.npz"))
    f1, f2, coords = (torch.from_numpy(z[k]) for k in ("f1", "f2", "coords"))
    pyr = FO.corr_pyramid(f1, f2)
    blk = CorrBlock(f1.to(DEV), f2.to(DEV), num_levels=4, radius=4)
    flat = torch.cat([p.reshape(-1) for p in pyr])
    got = blk.pyramid.cpu()
    assert got.numel() == flat.numel()
    assert (got - flat).abs().max() <= 1e-4 * flat.abs().max()
    look = blk(coords.to(DEV)).cpu()
    ref = FO.corr_lookup(pyr, coords)
    assert look.shape == ref.shape == (1, 324, 16, 24)
    assert (look - ref).abs().max() <= 2e-4 * ref.abs().max()
    # out-of-range windows read zeros (grid_sample zeros padding)
    far = coords.clone(); far[:, 0] += 500.0
    assert float(blk(far.to(DEV)).abs().max()) == 0.0


def test_update_block_and_convex_upsample(golden_dir):
    from b200 import nn as K
    from src.models.stage_1.core.update import BasicUpdateBlock
    fx = torch.load(os.path.join(golden_dir, "raft_update.pt"))
    z = np.load(os.path.join(golden_dir, "raft_corr.npz"))
    sd = seeded_weights(fx["shapes"], fx["seed"])
    ub = BasicUpdateBlock(types.SimpleNamespace(corr_levels=4, corr_radius=4), hidden_dim=128)
    ub.load_state_dict(sd)
    ub = ub.to(DEV)
    corr = torch.from_numpy(z["lookup"]).to(DEV)
    net, mask, delta = ub(fx["net"].to(DEV), fx["inp"].to(DEV), corr, fx["flow"].to(DEV))
    o_net, o_mask, o_delta = FO.update_block(sd, fx["net"], fx["inp"], torch.from_numpy(z["lookup"]), fx["flow"])
    assert (net.cpu() - o_net).abs().max() <= 1e-4
    assert (delta.cpu() - o_delta).abs().max() <= 1e-4
    assert (mask.cpu() - o_mask).abs().max() <= 1e-4
    up = K.convex_upsample(fx["flow"].to(DEV), mask)
    assert (up.cpu() - FO.convex_upsample(fx["flow"], o_mask)).abs().max() <= 5e-4


def test_unet_and_transformnet(golden_dir):
    from src.models.network_filter import UNet
    from src.models.network_local import TransformNet
    fx = torch.load(os.path.join(golden_dir, "stage2_nets.pt"))
    unet = UNet(in_channels=6, out_channels=3, init_features=32)
    usd = seeded_weights(fx["unet_shapes"], fx["unet_seed"])
    unet.load_state_dict(usd)
    y = unet.to(DEV)(fx["unet_x"].to(DEV))
    assert (y.cpu() - SO.unet_forward(usd, fx["unet_x"])).abs().max() <= 1e-4
    opts = types.SimpleNamespace(nf=32, norm="IN", model="TransformNet", blocks=5)
    tn = TransformNet(opts, nc_in=12, nc_out=3)
    assert len(tn.state_dict()) == 89
    tsd = seeded_weights(fx["tn_shapes"], fx["tn_seed"])
    tn.load_state_dict(tsd, strict=False)
    yy, (hid, cell) = tn.to(DEV)(fx["tn_x"].to(DEV), None)
    oy, oh, oc = SO.transformnet_forward(tsd, fx["tn_x"])
    assert (yy.cpu() - oy).abs().max() <= 1e-4
    assert (hid.cpu() - oh).abs().max() <= 1e-4
    assert (cell.cpu() - oc).abs().max() <= 1e-4


def test_full_raft_against_reference_fixture(golden_dir):
    """Whole RAFT (encoders with instance / folded batch norm, correlation, 3 update iterations, convex
    upsampling) against outputs of the reference model frozen by make_golden_nets.py."""
    import argparse
    from src.models.stage_1.core.raft import RAFT
    fx = torch.load(os.path.join(golden_dir, "raft_full.pt"))
    model = RAFT(argparse.Namespace(small=False, mixed_precision=True))
    assert len(model.state_dict()) == 179
    model.load_state_dict(seeded_weights(fx["shapes"], fx["seed"]), strict=False)
    model = model.to(DEV).eval()
    low, up = model(fx["im1"].to(DEV), fx["im2"].to(DEV), iters=3, test_mode=True)
    assert low.shape == fx["flow_low"].shape and up.shape == fx["flow_up"].shape == (1, 2, 128, 192)
    scale = fx["flow_up"].abs().max().item()
    assert (low.cpu() - fx["flow_low"]).abs().max() <= 2e-3 * max(fx["flow_low"].abs().max().item(), 1.0)
    assert (up.cpu() - fx["flow_up"]).abs().max() <= 2e-3 * max(scale, 1.0)
