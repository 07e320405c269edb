"""Host build of csrc/loss_math.h (the arithmetic the CUDA loss/sampling kernels run per sample)
against the golden coordinate table and against autograd of the oracle's loss functions. CPU only."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from b200 import _native as N
from csrc_build import ensure_built
from oracle import atlas_oracle as O


@pytest.fixture(scope="module")
def host():
    ensure_built()
    return N.hostcheck()


def test_norm_coord_bit_exact(host, golden_dir):
    z = np.load(os.path.join(golden_dir, "coords.npz"))
    ints = z["ints"].astype(np.float32).reshape(-1)
    out = np.empty_like(ints)
    for key, half in [("L768", 384.0), ("L432", 216.0), ("L256", 128.0), ("L160", 80.0), ("L90", 45.0),
                      ("L25", 12.5), ("T80", 40.0), ("T16", 8.0), ("T7", 3.5)]:
        host.b200_host_norm_coords(ints.ctypes.data, ints.size, np.float32(half), out.ctypes.data)
        assert np.array_equal(out, z[key].reshape(-1)), key
    v = (z["ints"].astype(np.float32) + z["fl"]).reshape(-1).astype(np.float32)   # int + flow, fp32 add
    host.b200_host_norm_coords(v.ctypes.data, v.size, np.float32(384.0), out.ctypes.data)
    assert np.array_equal(out, z["flt768"].reshape(-1))


def test_pe_frequencies(host):
    b = O.pe_frequencies(O.ATLAS_SPEC)
    assert b.dtype == torch.float32
    for k in range(10):
        assert host.b200_host_pe_freq(k) == float(b[k])


class _Queue:
    """Stands in for a network: returns pre-made leaf tensors in call order."""
    def __init__(self, outs): self.outs, self.i = list(outs), 0
    def __call__(self, x):
        o = self.outs[self.i]; self.i += 1
        assert o.shape[0] == x.shape[0]
        return o


@pytest.mark.parametrize("with_global", [True, False])
def test_loss_head_matches_oracle_autograd(host, golden_dir, with_global):
    z = np.load(os.path.join(golden_dir, "iteration.npz"))
    video = O.Video(**{k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("video_")})
    H, W, T = video.H, video.W, video.T
    inds = torch.from_numpy(z["inds"])
    B = inds.shape[0]
    jif = O.pixel_table(T, H, W)[:, inds]
    larger = max(W, H)
    g = torch.Generator().manual_seed(3)
    wf = video.mask_fwd[jif[1].squeeze(), jif[0].squeeze(), jif[2].squeeze(), 0] != 0
    wb = video.mask_bwd[jif[1].squeeze(), jif[0].squeeze(), jif[2].squeeze(), 0] != 0
    nf, nb = int(wf.sum()), int(wb.sum())
    # leaves: small neighbour differences so the Jacobians are O(1) like in training
    base = (torch.rand(B, 2, generator=g) - 0.5)
    def near(scale): return (base + scale * (torch.rand(B, 2, generator=g) - 0.5)).requires_grad_(True)
    uv = {"base": base.clone().requires_grad_(True), "xp": near(0.05), "yp": near(0.05),
          "ymd": near(0.06), "xmd": near(0.06), "ymg": near(2.0), "xmg": near(2.0),
          "f": near(0.05), "b": near(0.05)}
    y = {k: (torch.rand(B, 3, generator=g) * 1.6 - 0.8).requires_grad_(True) for k in ("base", "xp", "yp")}
    cfgd = O.DEFAULT_CONFIG
    # --- oracle composition (src/stage1_neural_atlas.py:174-227) with queued network outputs
    rgb = video.frames[jif[1], jif[0], :, jif[2]].squeeze(1)
    rgb_out = (y["base"] + 1.0) * 0.5
    gl = O.gradient_loss(video, jif, _Queue([uv["yp"], uv["xp"]]), _Queue([y["yp"], y["xp"]]), rgb_out, W)
    rl = (torch.norm(rgb_out - rgb, dim=1) ** 2).mean()
    rig = O.rigidity_loss(jif, 1, larger, T, _Queue([torch.cat((uv["ymd"], uv["xmd"]))]), uv["base"], uv_scale=0.8)
    total = rig * 1.0 + rl * 5000 + gl * 1000
    rigg = torch.zeros(())
    if with_global:
        rigg = O.rigidity_loss(jif, 100, larger, T, _Queue([torch.cat((uv["ymg"], uv["xmg"]))]), uv["base"],
                               uv_scale=0.8)
        total = total + 5.0 * rigg
    fl = O.flow_loss(video, jif, uv["base"], larger, _Queue([uv["f"][wf], uv["b"][wb]]), 0.8)
    total = total + 500.0 * fl
    total.backward()
    # --- host build of the CUDA loss head
    dxgt = video.frames_dx[jif[1], jif[0], :, jif[2]].squeeze(1)
    dygt = video.frames_dy[jif[1], jif[0], :, jif[2]].squeeze(1)
    order = ["base", "xp", "yp", "ymd", "xmd", "f", "b", "ymg", "xmg"]
    cfg = np.array([larger, 0.8, 1, 100, 5000, 1000, 1.0, 5.0 if with_global else 0.0, 500.0, 1.0 / B,
                    1.0 / nf, 1.0 / nb], np.float32)
    sums = np.zeros(6)
    worst = 0.0
    for s in range(B):
        vin = np.concatenate([np.stack([uv[k][s].detach().numpy() for k in order]).reshape(-1),
                              np.stack([y[k][s].detach().numpy() for k in ("base", "xp", "yp")]).reshape(-1),
                              rgb[s].numpy(), dxgt[s].numpy(), dygt[s].numpy(),
                              [float(wf[s]), float(wb[s])]]).astype(np.float32)
        out = np.zeros(33, np.float32)
        host.b200_host_sample_loss(vin.ctypes.data, cfg.ctypes.data, out.ctypes.data)
        sums += out[27:33]
        for gi, k in enumerate(order):
            ref = uv[k].grad[s].numpy() if uv[k].grad is not None else np.zeros(2)
            if k in ("ymg", "xmg") and not with_global:
                ref = np.zeros(2)
            got = out[2 * gi: 2 * gi + 2]
            if k == "f" and not wf[s]: assert np.all(got == 0)
            if k == "b" and not wb[s]: assert np.all(got == 0)
            tol = 2e-4 * max(np.abs(ref).max(), 1e-3)
            worst = max(worst, np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-3))
            assert np.allclose(got, ref, rtol=0, atol=tol), (s, k, got, ref)
        for gi, k in enumerate(("base", "xp", "yp")):
            ref = y[k].grad[s].numpy()
            got = out[18 + 3 * gi: 21 + 3 * gi]
            assert np.allclose(got, ref, rtol=1e-4, atol=1e-7), (s, k, got, ref)
    np.testing.assert_allclose(sums[0] / B, float(rl), rtol=1e-5)
    np.testing.assert_allclose(sums[1] / B, float(gl), rtol=1e-5)
    np.testing.assert_allclose(sums[2] / B, float(rig), rtol=1e-5)
    if with_global:
        np.testing.assert_allclose(sums[3] / B, float(rigg), rtol=1e-5)
    np.testing.assert_allclose(0.5 * (sums[4] / nf + sums[5] / nb), float(fl), rtol=1e-5)


def test_pretrain_term(host):
    g = torch.Generator().manual_seed(0)
    xy = torch.rand(32, 2, generator=g) * 2 - 1
    uv = (torch.rand(32, 2, generator=g) * 2 - 1).requires_grad_(True)
    loss = (xy * 0.8 - uv).norm(dim=1).mean()
    loss.backward()
    tot = 0.0
    for s in range(32):
        u = uv[s].detach().numpy().astype(np.float32)
        gg = np.zeros(2, np.float32)
        tot += host.b200_host_pretrain(float(xy[s, 0]), float(xy[s, 1]), u.ctypes.data, np.float32(0.8),
                                       np.float32(1 / 32), gg.ctypes.data)
        np.testing.assert_allclose(gg, uv.grad[s].numpy(), rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(tot / 32, float(loss), rtol=1e-6)
