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


@pytest.mark.parametrize("with_global", [True, False])
def test_seg_loss_head_matches_oracle_autograd(host, golden_dir, with_global):
    """csrc/seg_loss_math.h (the per-sample code of seg_loss_kernel) against autograd of oracle/seg_oracle.py."""
    from oracle import seg_oracle as S
    from seg_common import load_fixture
    z, video, masks, _ = load_fixture(golden_dir)
    H, W, T = video.H, video.W, video.T
    inds = torch.from_numpy(z["inds"])
    B = inds.shape[0]
    jif = O.pixel_table(T, H, W)[:, inds]
    larger = max(W, H)
    g = torch.Generator().manual_seed(8)
    wf = video.mask_fwd[jif[1].squeeze(), jif[0].squeeze(), jif[2].squeeze(), 0] != 0
    wb = video.mask_bwd[jif[1].squeeze(), jif[0].squeeze(), jif[2].squeeze(), 0] != 0
    nf, nb = int(wf.sum()), int(wb.sum())
    order = ["base", "xp", "yp", "ymd", "xmd", "f", "b", "ymg", "xmg"]
    spread = dict(base=0, xp=0.05, yp=0.05, ymd=0.06, xmd=0.06, f=0.05, b=0.05, ymg=2.0, xmg=2.0)

    def mapping_leaves():
        base = torch.rand(B, 2, generator=g) - 0.5
        return {k: (base + spread[k] * (torch.rand(B, 2, generator=g) - 0.5)).requires_grad_(True) for k in order}
    uv1, uv2 = mapping_leaves(), mapping_leaves()
    a_order = ["base", "xp", "yp", "f", "b"]
    ar = {k: (torch.rand(B, 1, generator=g) * 1.8 - 0.9).requires_grad_(True) for k in a_order}
    y_order = ["1base", "1xp", "1yp", "2base", "2xp", "2yp"]
    y = {k: (torch.rand(B, 3, generator=g) * 1.6 - 0.8).requires_grad_(True) for k in y_order}
    c = S.SEG_CONFIG
    # --- oracle composition (stage1_neural_atlas_seg.py:225-315) with queued network outputs
    rgb = video.frames[jif[1], jif[0], :, jif[2]].squeeze(1)
    a_gt = masks[jif[1], jif[0], jif[2]].squeeze(1).unsqueeze(-1)
    alpha = S.alpha_of(ar["base"])
    rgb1, rgb2 = (y["1base"] + 1.0) * 0.5, (y["2base"] + 1.0) * 0.5
    out = rgb1 * alpha + rgb2 * (1.0 - alpha)
    gl = S.gradient_loss_seg(video, jif, _Queue([uv1["yp"], uv1["xp"]]), _Queue([uv2["yp"], uv2["xp"]]),
                             _Queue([y["1yp"], y["1xp"], y["2yp"], y["2xp"]]), _Queue([ar["xp"], ar["yp"]]), out, W)
    rl = (torch.norm(out - rgb, dim=1) ** 2).mean()
    sp = (torch.norm(rgb1 * (1.0 - alpha), dim=1) ** 2).mean()
    r1 = O.rigidity_loss(jif, 1, larger, T, _Queue([torch.cat((uv1["ymd"], uv1["xmd"]))]), uv1["base"], uv_scale=0.8)
    r2 = O.rigidity_loss(jif, 1, larger, T, _Queue([torch.cat((uv2["ymd"], uv2["xmd"]))]), uv2["base"], uv_scale=0.8)
    total = r1 + r2
    g1 = g2 = torch.zeros(())
    if with_global:
        g1 = O.rigidity_loss(jif, 100, larger, T, _Queue([torch.cat((uv1["ymg"], uv1["xmg"]))]), uv1["base"], uv_scale=0.8)
        g2 = O.rigidity_loss(jif, 100, larger, T, _Queue([torch.cat((uv2["ymg"], uv2["xmg"]))]), uv2["base"], uv_scale=0.8)
        total = total + 5.0 * g1 + 50.0 * g2
    f1 = S.flow_loss_alpha(video, jif, uv1["base"], larger, _Queue([uv1["f"][wf], uv1["b"][wb]]), 0.8, alpha)
    f2 = S.flow_loss_alpha(video, jif, uv2["base"], larger, _Queue([uv2["f"][wf], uv2["b"][wb]]), 0.8, 1 - alpha)
    fa = S.flow_alpha_loss(video, jif, alpha, larger, _Queue([ar["f"][wf], ar["b"][wb]]))
    bce = torch.mean(-a_gt * torch.log(alpha) - (1 - a_gt) * torch.log(1 - alpha))
    total = total + rl * c["rgb_coeff"] + c["optical_flow_coeff"] * (f1 + f2) + bce * c["alpha_bootstrapping_factor"] \
        + fa * c["alpha_flow_factor"] + sp * c["sparsity_coeff"] + gl * c["gradient_loss_coeff"]
    total.backward()
    # --- host build of the CUDA loss head
    dxgt = video.frames_dx[jif[1], jif[0], :, jif[2]].squeeze(1)
    dygt = video.frames_dy[jif[1], jif[0], :, jif[2]].squeeze(1)
    cfg = np.array([larger, 0.8, 1, 100, c["rgb_coeff"], c["gradient_loss_coeff"], 1.0, 5.0, 50.0, c["optical_flow_coeff"],
                    c["alpha_flow_factor"], c["sparsity_coeff"], c["alpha_bootstrapping_factor"], float(with_global),
                    1.0 / B, 1.0 / nf, 1.0 / nb], np.float32)
    sums = np.zeros(14)

    def grad_of(t, s, n):
        return t.grad[s].numpy() if t.grad is not None else np.zeros(n)
    for s in range(B):
        vin = np.concatenate([np.stack([uv1[k][s].detach().numpy() for k in order]).reshape(-1),
                              np.stack([uv2[k][s].detach().numpy() for k in order]).reshape(-1),
                              [float(ar[k][s].detach()) for k in a_order],
                              np.stack([y[k][s].detach().numpy() for k in y_order]).reshape(-1),
                              rgb[s].numpy(), dxgt[s].numpy(), dygt[s].numpy(),
                              [float(a_gt[s]), float(wf[s]), float(wb[s])]]).astype(np.float32)
        assert vin.size == 71
        res = np.zeros(73, np.float32)
        host.b200_host_seg_sample_loss(vin.ctypes.data, cfg.ctypes.data, res.ctypes.data)
        sums += res[59:73]
        for net, leaves in ((0, uv1), (1, uv2)):
            for gi, k in enumerate(order):
                ref = grad_of(leaves[k], s, 2)
                if k in ("ymg", "xmg") and not with_global:
                    ref = np.zeros(2)
                got = res[18 * net + 2 * gi: 18 * net + 2 * gi + 2]
                if (k == "f" and not wf[s]) or (k == "b" and not wb[s]):
                    assert np.all(got == 0)
                    continue
                assert np.allclose(got, ref, rtol=0, atol=2e-4 * max(np.abs(ref).max(), 1e-3)), (s, net, k, got, ref)
        for ai, k in enumerate(a_order):
            ref = grad_of(ar[k], s, 1)[0]
            got = res[36 + ai]
            if (k == "f" and not wf[s]) or (k == "b" and not wb[s]):
                assert got == 0
                continue
            assert abs(got - ref) <= 2e-4 * max(abs(ref), 1e-3), (s, k, got, ref)
        for yi, k in enumerate(y_order):
            ref = grad_of(y[k], s, 3)
            got = res[41 + 3 * yi: 44 + 3 * yi]
            assert np.allclose(got, ref, rtol=1e-4, atol=1e-7), (s, k, got, ref)
    want = [rl, gl, sp, r1, r2, g1, g2]
    for i, w in enumerate(want):
        np.testing.assert_allclose(sums[i] / B, float(w), rtol=2e-5, atol=1e-12)
    np.testing.assert_allclose(0.5 * (sums[7] / nf + sums[8] / nb), float(f1), rtol=2e-5)
    np.testing.assert_allclose(0.5 * (sums[9] / nf + sums[10] / nb), float(f2), rtol=2e-5)
    np.testing.assert_allclose(0.5 * (sums[11] / nf + sums[12] / nb), float(fa), rtol=2e-5)
    np.testing.assert_allclose(sums[13] / B, float(bce), rtol=2e-5)
