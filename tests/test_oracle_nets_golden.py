"""RAFT / stage-2 oracles replayed against fixtures frozen from the reference modules
(tests/golden/make_golden_nets.py).  CPU only."""
import os

import numpy as np
import torch

from nets_common import seeded_weights
from oracle import flow_oracle as FO
from oracle import stage2_oracle as SO


def test_corr_pyramid_and_lookup(golden_dir):
    torch.set_num_threads(1)
    z = np.load(os.path.join(golden_dir, "raft_corr.npz"))
    f1, f2, coords = (torch.from_numpy(z[k]) for k in ("f1", "f2", "coords"))
    pyr = FO.corr_pyramid(f1, f2)
    np.testing.assert_allclose(pyr[0][:8].numpy(), z["lvl0_head"], atol=2e-5)
    np.testing.assert_allclose(pyr[3].numpy(), z["lvl3"], atol=2e-5)
    np.testing.assert_allclose(FO.corr_lookup(pyr, coords).numpy(), z["lookup"], atol=5e-5)
    assert FO.corr_lookup(pyr, coords).shape == (1, 324, 16, 24)


def test_update_block_and_upsample(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "raft_update.pt"))
    z = np.load(os.path.join(golden_dir, "raft_corr.npz"))
    sd = seeded_weights(fx["shapes"], fx["seed"])
    net, mask, delta = FO.update_block(sd, fx["net"], fx["inp"], torch.from_numpy(z["lookup"]), fx["flow"])
    torch.testing.assert_close(net, fx["out_net"], atol=2e-5, rtol=0)
    torch.testing.assert_close(delta, fx["out_delta"], atol=2e-5, rtol=0)
    torch.testing.assert_close(mask[0, :8], fx["out_mask_head"], atol=2e-5, rtol=0)
    torch.testing.assert_close(FO.convex_upsample(fx["flow"], mask), fx["up"], atol=5e-5, rtol=0)


def test_stage2_networks(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "stage2_nets.pt"))
    usd = seeded_weights(fx["unet_shapes"], fx["unet_seed"])
    torch.testing.assert_close(SO.unet_forward(usd, fx["unet_x"]), fx["unet_y"], atol=2e-5, rtol=0)
    tsd = seeded_weights(fx["tn_shapes"], fx["tn_seed"])
    y, h, _ = SO.transformnet_forward(tsd, fx["tn_x"])
    torch.testing.assert_close(y, fx["tn_y"], atol=2e-5, rtol=0)
    torch.testing.assert_close(h, fx["tn_h"], atol=2e-5, rtol=0)
    assert fx["tn_keys"] == 89
