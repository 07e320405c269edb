"""Pins oracle/seg_oracle.py against the reference's own segmentation-variant functions and freezes
``tests/golden/seg_iteration.npz``.

Run ONLY in the build container (needs /root/reference):

    python tests/golden/make_golden_seg.py

The reference's ``IMLP`` and ``loss_utils`` are loaded by file path, unchanged; the loop body of
src/stage1_neural_atlas_seg.py:207-315 is driven with them on a small seeded clip, the oracle
restatement must return BIT-IDENTICAL losses and parameter gradients, and three optimiser steps of
both must leave bit-identical parameters.  ``tests/test_seg_oracle_golden.py`` replays the fixture
without the reference.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "all-in-one-deflicker_b200"))
REF = "/root/reference"
sys.modules.setdefault("imageio", types.ModuleType("imageio"))


def _load_reference(name, rel):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert os.path.realpath(mod.__file__).startswith(REF + os.sep), mod.__file__
    return mod


_ref_inn = _load_reference("ref_implicit_neural_networks", "src/models/stage_1/implicit_neural_networks.py")
ref_loss = _load_reference("ref_loss_utils", "src/models/stage_1/loss_utils.py")
ref_unwrap = _load_reference("ref_unwrap_utils", "src/models/stage_1/unwrap_utils.py")
IMLP = _ref_inn.IMLP

from oracle import atlas_oracle as O  # noqa: E402
from oracle import seg_oracle as S  # noqa: E402
from b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ORDER = ("mapping1", "mapping2", "atlas", "alpha")          # construction order of the script
GROUPS = ("mapping1", "mapping2", "alpha", "atlas")         # optimiser group order


def build_reference_nets(seed):
    """stage1_neural_atlas_seg.py:127-161 with config_flow_100.json."""
    torch.manual_seed(seed)
    m1 = IMLP(input_dim=3, output_dim=2, hidden_dim=256, use_positional=False, positional_dim=4, num_layers=6,
              skip_layers=[], verbose=False)
    m2 = IMLP(input_dim=3, output_dim=2, hidden_dim=256, use_positional=False, positional_dim=2, num_layers=4,
              skip_layers=[], verbose=False)
    at = IMLP(input_dim=2, output_dim=3, hidden_dim=256, use_positional=True, positional_dim=10, num_layers=8,
              skip_layers=[4, 7], verbose=False)
    al = IMLP(input_dim=3, output_dim=1, hidden_dim=256, use_positional=True, positional_dim=5, num_layers=8,
              skip_layers=[], verbose=False)
    return dict(mapping1=m1, mapping2=m2, atlas=at, alpha=al)


def same(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.equal(a, b), f"{what}: oracle differs from reference, max abs {float((a - b).abs().max())}"


def reference_iteration(ref, video, mask_frames, table, inds, it, cfg):
    """The loop body, reference functions only."""
    H, W, T = video.H, video.W, video.T
    larger = np.maximum(W, H)
    m1, m2, at, al = ref["mapping1"], ref["mapping2"], ref["atlas"], ref["alpha"]
    boot = cfg["alpha_bootstrapping_factor"] if it <= cfg["stop_bootstrapping_iteration"] else 0
    jif = table[:, inds]
    rgb_cur = video.frames[jif[1, :], jif[0, :], :, jif[2, :]].squeeze(1)
    a_gt = mask_frames[jif[1, :], jif[0, :], jif[2, :]].squeeze(1).unsqueeze(-1)
    xyt = torch.cat((jif[0, :] / (larger / 2) - 1, jif[1, :] / (larger / 2) - 1, jif[2, :] / (T / 2.0) - 1), dim=1)
    uv1, uv2 = m1(xyt), m2(xyt)
    alpha = 0.5 * (al(xyt) + 1.0)
    alpha = alpha * 0.99
    alpha = alpha + 0.001
    rgb1 = (at(uv1 * 0.5 + 0.5) + 1.0) * 0.5
    rgb2 = (at(uv2 * 0.5 - 0.5) + 1.0) * 0.5
    out = rgb1 * alpha + rgb2 * (1.0 - alpha)
    gl = ref_loss.get_gradient_loss(video.frames_dx, video.frames_dy, jif, m1, m2, at, out, "cpu", W, T, al)
    out_not = rgb1 * (1.0 - alpha)
    rl = (torch.norm(out - rgb_cur, dim=1) ** 2).mean()
    sp = (torch.norm(out_not, dim=1) ** 2).mean()
    s = cfg["uv_mapping_scale"]
    r1 = ref_loss.get_rigidity_loss(jif, cfg["derivative_amount"], larger, T, m1, uv1, "cpu", uv_mapping_scale=s)
    r2 = ref_loss.get_rigidity_loss(jif, cfg["derivative_amount"], larger, T, m2, uv2, "cpu", uv_mapping_scale=s)
    terms = dict(gradient=gl, rgb=rl, sparsity=sp, rigidity1=r1, rigidity2=r2)
    total = cfg["rigidity_coeff"] * (r1 + r2)
    if it <= cfg["stop_global_rigidity"]:
        g1 = ref_loss.get_rigidity_loss(jif, cfg["global_rigidity_derivative_amount_fg"], larger, T, m1, uv1, "cpu",
                                        uv_mapping_scale=s)
        g2 = ref_loss.get_rigidity_loss(jif, cfg["global_rigidity_derivative_amount_bg"], larger, T, m2, uv2, "cpu",
                                        uv_mapping_scale=s)
        terms.update(rigidity_global1=g1, rigidity_global2=g2)
        total = total + cfg["global_rigidity_coeff_fg"] * g1 + cfg["global_rigidity_coeff_bg"] * g2
    f1 = ref_loss.get_optical_flow_loss(jif, uv1, video.flow_bwd, video.mask_bwd, larger, T, m1, video.flow_fwd,
                                        video.mask_fwd, s, "cpu", use_alpha=True, alpha=alpha)
    f2 = ref_loss.get_optical_flow_loss(jif, uv2, video.flow_bwd, video.mask_bwd, larger, T, m2, video.flow_fwd,
                                        video.mask_fwd, s, "cpu", use_alpha=True, alpha=1 - alpha)
    fa = ref_loss.get_optical_flow_alpha_loss(al, jif, alpha, video.flow_bwd, video.mask_bwd, larger, T,
                                              video.flow_fwd, video.mask_fwd, "cpu")
    bce = torch.mean(-a_gt * torch.log(alpha) - (1 - a_gt) * torch.log(1 - alpha))
    terms.update(flow1=f1, flow2=f2, flow_alpha=fa, bootstrapping=bce)
    total = total + rl * cfg["rgb_coeff"] + cfg["optical_flow_coeff"] * (f1 + f2) + bce * boot \
        + fa * cfg["alpha_flow_factor"] + sp * cfg["sparsity_coeff"] + gl * cfg["gradient_loss_coeff"]
    terms["total"] = total
    return terms, dict(uv1=uv1, uv2=uv2, alpha=alpha, rgb_out=out)


def seg_masks(H, W, T, seed):
    """A soft-edged blob moving over the clip: values in [0, 1] like a resized 8-bit matte."""
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    m = np.zeros((H, W, T), np.float64)
    for t in range(T):
        cy, cx = H * (0.4 + 0.03 * t), W * (0.35 + 0.04 * t)
        r = np.hypot((yy - cy) / (0.28 * H), (xx - cx) / (0.22 * W))
        m[:, :, t] = np.round(np.clip(1.5 - r * 1.5 + 0.02 * g.standard_normal((H, W)), 0, 1) * 255) / 255
    return torch.from_numpy(m).float()


def main():
    torch.set_num_threads(1)
    cfg = S.SEG_CONFIG
    ref = build_reference_nets(4321)
    torch.manual_seed(4321)
    nets = S.init_nets()
    for k in ORDER:
        for p, q in zip(nets[k], ref[k].parameters()):
            same(p, q.detach(), f"{k} init")
    H, W, T, B = 24, 40, 6, 64
    data = synth.throughput_set(H, W, T, seed=5)
    video = O.Video(**data)
    masks = seg_masks(H, W, T, 9)
    table = ref_unwrap.get_tuples(T, video.frames)
    inds = torch.randint(table.shape[1], (B, 1), generator=torch.Generator().manual_seed(13))
    fixture = dict(inds=inds.numpy(), H=H, W=W, T=T, masks=masks.numpy())
    for k, v in data.items():
        fixture["video_" + k] = v.numpy()
    fixture["init_seed"] = 4321          # parameters = S.init_nets() after torch.manual_seed(4321) (checked above)
    for k in ORDER:
        fixture[f"init_{k}_sum"] = np.float64(sum(p.double().sum() for p in nets[k]))
    for it in (0, 6000, 10001):
        for k in ORDER:
            for p in ref[k].parameters():
                p.grad = None
        rt, rmid = reference_iteration(ref, video, masks, table, inds, it, cfg)
        rt["total"].backward()
        mine = {k: [p.detach().clone().requires_grad_(True) for p in nets[k]] for k in ORDER}
        ot = S.seg_iteration_losses(video, masks, mine, inds, it, cfg)
        ot["total"].backward()
        assert set(ot) == set(rt)
        for k in rt:
            same(ot[k].detach(), rt[k].detach(), f"it {it} loss {k}")
        tag = f"it{it}_"
        for k in ORDER:
            for i, (p, q) in enumerate(zip(mine[k], ref[k].parameters())):
                same(p.grad, q.grad, f"it {it} grad {k}[{i}]")
                gflat = p.grad.flatten()
                fixture[tag + f"grad_{k}_{i}_sum"] = np.float64(gflat.double().sum())
                fixture[tag + f"grad_{k}_{i}_abs"] = np.float64(gflat.double().abs().sum())
                fixture[tag + f"grad_{k}_{i}_head"] = gflat[:32].numpy()
        for k, v in rt.items():
            fixture[tag + "loss_" + k] = np.float32(v.detach())
        for k, v in rmid.items():
            fixture[tag + k] = v.detach().numpy()

    # three optimiser steps: reference modules + torch Adam (group order of the script) vs the oracle
    ref_opt = torch.optim.Adam([{"params": list(ref[k].parameters())} for k in GROUPS], lr=1e-4)
    mine = {k: [p.detach().clone().requires_grad_(True) for p in nets[k]] for k in ORDER}
    opt = S.make_optimizer(mine)
    gi = torch.Generator().manual_seed(17)
    traj, all_inds = [], []
    for it in range(3):
        ii = torch.randint(table.shape[1], (B, 1), generator=gi)
        all_inds.append(ii.numpy())
        rt, _ = reference_iteration(ref, video, masks, table, ii, it, cfg)
        ref_opt.zero_grad(); rt["total"].backward(); ref_opt.step()
        ot = S.seg_iteration_losses(video, masks, mine, ii, it, cfg)
        opt.zero_grad(); ot["total"].backward(); opt.step()
        traj.append([float(ot[k].detach()) for k in sorted(ot)])
    for k in ORDER:
        for p, q in zip(mine[k], ref[k].parameters()):
            same(p.detach(), q.detach(), f"trajectory params {k}")
    fixture["traj_inds"] = np.stack(all_inds)
    fixture["traj_keys"] = np.array(sorted(ot))
    fixture["traj_losses"] = np.array(traj, np.float64)
    for k in ORDER:
        fixture[f"traj_{k}_sum"] = np.float64(sum(p.double().sum() for p in mine[k]).detach())
        fixture[f"traj_{k}_head"] = mine[k][0].detach().flatten()[:64].numpy()

    # render of one frame (composite + alpha), reference modules vs oracle
    f = 3
    with torch.no_grad():
        ys, xs = torch.where(torch.ones(H, W) > 0)
        larger = np.maximum(np.int64(W), np.int64(H))
        xyt = torch.cat((xs.unsqueeze(1) / (larger / 2) - 1, ys.unsqueeze(1) / (larger / 2) - 1,
                         (f / (T / 2.0) - 1) * torch.ones(ys.shape[0], 1)), dim=1)
        a = 0.5 * (ref["alpha"](xyt) + 1.0)
        a = a * 0.99
        a = a + 0.001
        c1 = (ref["atlas"](ref["mapping1"](xyt) * 0.5 + 0.5) + 1) * 0.5
        c2 = (ref["atlas"](ref["mapping2"](xyt) * 0.5 - 0.5) + 1) * 0.5
        ref_img = (c1 * a + c2 * (1.0 - a)).view(H, W, 3)
    img, alpha_img = S.render_frame_seg({k: [p.detach() for p in mine[k]] for k in ORDER}, f, H, W, T)
    same(img, ref_img, "render")
    same(alpha_img, a.view(H, W), "render alpha")
    fixture.update(render_frame=f, render_img=img.numpy(), render_alpha=alpha_img.numpy())
    np.savez_compressed(os.path.join(OUT, "seg_iteration.npz"), **fixture)
    print("seg fixture written")


if __name__ == "__main__":
    main()
