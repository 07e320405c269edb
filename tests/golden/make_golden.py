"""Pins oracle/atlas_oracle.py against the reference's own modules and freezes fixtures.

Run ONLY in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's stage-1 modules unchanged (with an ``imageio`` stub — only the
visualisation helper uses it), drives them on small seeded inputs, asserts that the oracle
restatement returns BIT-IDENTICAL tensors, and writes the inputs/outputs to
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` replays the fixtures without the
reference.
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
    """Reference modules are loaded BY FILE PATH: the repo's own mirror package is also called `src`, and a
    regular package on sys.path would shadow the reference's namespace package.  The three modules have no
    intra-package imports, so this is exactly the code the reference runs."""
    import importlib.util
    path = os.path.join(REF, rel)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert os.path.realpath(mod.__file__).startswith(REF + os.sep), mod.__file__
    return mod


_ref_inn = _load_reference("ref_implicit_neural_networks", "src/models/stage_1/implicit_neural_networks.py")
ref_loss = _load_reference("ref_loss_utils", "src/models/stage_1/loss_utils.py")
ref_unwrap = _load_reference("ref_unwrap_utils", "src/models/stage_1/unwrap_utils.py")
IMLP, positionalEncoding_vec = _ref_inn.IMLP, _ref_inn.positionalEncoding_vec

from oracle import atlas_oracle as O  # noqa: E402
from b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def build_reference_nets(seed):
    torch.manual_seed(seed)
    m = IMLP(input_dim=3, output_dim=2, hidden_dim=256, use_positional=False, positional_dim=4,
             num_layers=6, skip_layers=[], verbose=False)
    a = IMLP(input_dim=2, output_dim=3, hidden_dim=256, use_positional=True, positional_dim=10,
             num_layers=8, skip_layers=[4, 7], verbose=False)
    return m, a


def same(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.equal(a, b) or (torch.isnan(a).all() and torch.isnan(b).all()), \
        f"{what}: oracle differs from reference, max abs {float((a - b).abs().max())}"


def main():
    torch.set_num_threads(1)   # deterministic summation order inside addmm for the fixtures
    # ---------------- 1. init parity: oracle init == nn.Linear init on the same RNG stream
    ref_m, ref_a = build_reference_nets(1234)
    torch.manual_seed(1234)
    map_p = O.init_mlp(O.MAPPING_SPEC)
    atl_p = O.init_mlp(O.ATLAS_SPEC)
    for p, q in zip(map_p, ref_m.parameters()):
        same(p, q.detach(), "mapping init")
    for p, q in zip(atl_p, ref_a.parameters()):
        same(p, q.detach(), "atlas init")
    assert O.MAPPING_SPEC.num_params() == 264706 and O.ATLAS_SPEC.num_params() == 416379

    # ---------------- 2. PE + forward parity
    g = torch.Generator().manual_seed(7)
    x3 = torch.rand(96, 3, generator=g) * 2 - 1
    x2 = torch.rand(96, 2, generator=g)
    same(O.positional_encoding(x2, O.pe_frequencies(O.ATLAS_SPEC)), positionalEncoding_vec(x2, ref_a.b), "PE")
    y_map = O.mlp_forward(O.MAPPING_SPEC, map_p, x3)
    y_atl = O.mlp_forward(O.ATLAS_SPEC, atl_p, x2)
    same(y_map, ref_m(x3).detach(), "mapping fwd")
    same(y_atl, ref_a(x2).detach(), "atlas fwd")
    np.savez(os.path.join(OUT, "params_seed1234.npz"),
             **{f"map{i}": p.numpy() for i, p in enumerate(map_p)},
             **{f"atl{i}": p.numpy() for i, p in enumerate(atl_p)})
    np.savez(os.path.join(OUT, "mlp_forward.npz"), x3=x3.numpy(), x2=x2.numpy(),
             pe=O.positional_encoding(x2, O.pe_frequencies(O.ATLAS_SPEC)).numpy(),
             y_map=y_map.detach().numpy(), y_atl=y_atl.detach().numpy())

    # ---------------- 3. coordinate normalisation (int64 / numpy scalar -> fp32)
    ints = torch.arange(-120, 1200, dtype=torch.int64).unsqueeze(1)
    coords = {}
    for L in (768, 432, 256, 160, 90, 25):
        coords[f"L{L}"] = (ints / (np.int64(L) / 2) - 1).numpy()
    for T in (80, 16, 7):
        coords[f"T{T}"] = (ints / (T / 2.0) - 1).numpy()
    fl = torch.linspace(-3, 3, 1320).unsqueeze(1)
    coords["flt768"] = ((ints + fl) / (np.int64(768) / 2) - 1).numpy()
    np.savez(os.path.join(OUT, "coords.npz"), ints=ints.numpy(), fl=fl.numpy(), **coords)

    # ---------------- 4. one full iteration: every loss + gradients, both regimes
    H, W, T, B = 24, 40, 6, 64
    data = synth.throughput_set(H, W, T, seed=3)
    video = O.Video(**data)
    table = ref_unwrap.get_tuples(T, video.frames)
    same(O.pixel_table(T, H, W), table, "pixel table")
    inds = torch.randint(table.shape[1], (B, 1), generator=torch.Generator().manual_seed(11))
    larger = np.maximum(W, H)
    fixture = dict(inds=inds.numpy(), H=H, W=W, T=T)
    for k, v in data.items():
        fixture["video_" + k] = v.numpy()
    for it in (0, 6000):
        for p in list(ref_m.parameters()) + list(ref_a.parameters()):
            p.grad = None
        jif = table[:, inds]
        rgb_cur = video.frames[jif[1, :], jif[0, :], :, jif[2, :]].squeeze(1)
        xyt = torch.cat((jif[0, :] / (larger / 2) - 1, jif[1, :] / (larger / 2) - 1,
                         jif[2, :] / (T / 2.0) - 1), dim=1)
        uv = ref_m(xyt)
        rgb_out = (ref_a(uv * 0.5 + 0.5) + 1.0) * 0.5
        gl = ref_loss.get_gradient_loss_single(video.frames_dx, video.frames_dy, jif, ref_m, ref_a,
                                               rgb_out, "cpu", W, T)
        rl = (torch.norm(rgb_out - rgb_cur, dim=1) ** 2).mean()
        rig = ref_loss.get_rigidity_loss(jif, 1, larger, T, ref_m, uv, "cpu", uv_mapping_scale=0.8)
        total = 1.0 * rig
        if it <= 5000:
            rig_g = ref_loss.get_rigidity_loss(jif, 100, larger, T, ref_m, uv, "cpu", uv_mapping_scale=0.8)
            total = total + 5.0 * rig_g
        alpha = torch.ones(B, 1)
        fll = ref_loss.get_optical_flow_loss(jif, uv, video.flow_bwd, video.mask_bwd, larger, T, ref_m,
                                             video.flow_fwd, video.mask_fwd, 0.8, "cpu",
                                             use_alpha=True, alpha=alpha)
        total = total + rl * 5000 + 500.0 * fll + gl * 1000
        total.backward()

        mp = [p.detach().clone().requires_grad_(True) for p in map_p]
        ap = [p.detach().clone().requires_grad_(True) for p in atl_p]
        terms = O.iteration_losses(video, mp, ap, inds, it)
        terms["total"].backward()
        same(terms["gradient"].detach(), gl.detach(), "gradient loss")
        same(terms["rgb"].detach(), rl.detach(), "rgb loss")
        same(terms["rigidity"].detach(), rig.detach(), "rigidity loss")
        same(terms["flow"].detach(), fll.detach(), "flow loss")
        same(terms["total"].detach(), total.detach(), "total loss")
        if it <= 5000:
            same(terms["rigidity_global"].detach(), rig_g.detach(), "global rigidity")
        for i, (p, q) in enumerate(zip(mp + ap, list(ref_m.parameters()) + list(ref_a.parameters()))):
            same(p.grad, q.grad, f"grad {i}")
        tag = f"it{it}_"
        fixture[tag + "uv"] = uv.detach().numpy()
        fixture[tag + "rgb_out"] = rgb_out.detach().numpy()
        for k, v in terms.items():
            fixture[tag + "loss_" + k] = np.float32(v.detach())
        for i, p in enumerate(mp + ap):
            gflat = p.grad.flatten()
            fixture[tag + f"grad{i}_sum"] = np.float64(gflat.double().sum())
            fixture[tag + f"grad{i}_abs"] = np.float64(gflat.double().abs().sum())
            fixture[tag + f"grad{i}_head"] = gflat[:32].numpy()
    np.savez_compressed(os.path.join(OUT, "iteration.npz"), **fixture)

    # ---------------- 5. three Adam steps (trajectory) : reference optimiser vs oracle optimiser
    ref_opt = torch.optim.Adam([{"params": list(ref_m.parameters())}, {"params": list(ref_a.parameters())}], lr=1e-4)
    mp = [p.detach().clone().requires_grad_(True) for p in map_p]
    ap = [p.detach().clone().requires_grad_(True) for p in atl_p]
    opt = O.make_optimizer(mp, ap)
    gi = torch.Generator().manual_seed(5)
    traj = []
    all_inds = []
    for it in range(3):
        inds = torch.randint(table.shape[1], (B, 1), generator=gi)
        all_inds.append(inds.numpy())
        out = O.train_iteration(video, mp, ap, opt, inds, it)
        traj.append([out[k] for k in ("total", "rgb", "gradient", "rigidity", "rigidity_global", "flow")])
    np.savez(os.path.join(OUT, "trajectory.npz"), inds=np.stack(all_inds), losses=np.array(traj, np.float64),
             map0_head=mp[0].detach().flatten()[:64].numpy(), atl14_head=ap[14].detach().flatten()[:64].numpy(),
             map_sum=np.float64(sum(p.double().sum() for p in mp).detach()),
             atl_sum=np.float64(sum(p.double().sum() for p in ap).detach()))
    del ref_opt

    # ---------------- 6. pre_train_mapping: 1 sweep over 2 frames, reference vs oracle loop
    ref_m2, _ = build_reference_nets(99)
    torch.manual_seed(99)
    mp = [p.requires_grad_(True) for p in O.init_mlp(O.MAPPING_SPEC)]
    Hp, Wp, Tp = 20, 36, 2
    torch.manual_seed(5)
    ref_unwrap.pre_train_mapping(ref_m2, Tp, 0.8, resx=Wp, resy=Hp, larger_dim=np.maximum(Wp, Hp),
                                 device="cpu", pretrain_iters=1)
    torch.manual_seed(5)
    opt = torch.optim.Adam(mp, lr=1e-4)
    pl = []
    for f in range(Tp):
        ys = torch.randint(Hp, (10000, 1))
        xs = torch.randint(Wp, (10000, 1))
        loss = O.pretrain_losses(mp, f, ys, xs, Tp, max(Wp, Hp), 0.8)
        opt.zero_grad()
        loss.backward()
        opt.step()
        pl.append(float(loss))
    for p, q in zip(mp, ref_m2.parameters()):
        same(p.detach(), q.detach(), "pretrain params")
    np.savez(os.path.join(OUT, "pretrain.npz"), losses=np.array(pl), H=Hp, W=Wp, T=Tp,
             w0_head=mp[0].detach().flatten()[:64].numpy(), w10_head=mp[10].detach().flatten()[:64].numpy())

    # ---------------- 7. render + uint8 + PSNR
    img = O.render_frame(map_p, atl_p, 2, H, W, T)
    np.savez(os.path.join(OUT, "render.npz"), frame=2, img=img.numpy(), u8=O.to_uint8(img),
             psnr=O.psnr(video.frames[:, :, :, 2], img))
    # ---------------- 8. per-pixel evaluation maps (evaluate.py:640-700): uv, rigidity of every pixel, flow error
    ev = {}
    for f in (2, T - 1):
        ys, xs = torch.where(torch.ones(H, W) > 0)
        with torch.no_grad():
            xyt = torch.cat((xs.unsqueeze(1) / (larger / 2) - 1, ys.unsqueeze(1) / (larger / 2) - 1,
                             (f / (T / 2.0) - 1) * torch.ones(ys.shape[0], 1)), dim=1)
            uv_r = ref_m(xyt)
            jf = torch.cat((xs.unsqueeze(-1), ys.unsqueeze(-1), torch.ones_like(ys.unsqueeze(-1)) * f), dim=1).T.unsqueeze(-1)
            rig_r = ref_loss.get_rigidity_loss(jf, 1, larger, T, ref_m, uv_r, "cpu", uv_mapping_scale=0.8, return_all=True)
            if f < T - 1:
                fl_r = ref_loss.get_optical_flow_loss_all(jf, uv_r, larger, T, ref_m, video.flow_fwd, video.mask_fwd, 0.8,
                                                          "cpu", alpha=torch.ones(ys.shape[0], 1))
            else:
                fl_r = torch.zeros(ys.shape[0])
        uv_o, rig_o, fl_o = O.eval_maps(video, [p.detach() for p in ref_m.parameters()], f)
        same(uv_o.reshape(-1, 2), uv_r, "eval uv")
        same(rig_o.reshape(-1), rig_r, "eval rigidity")
        same(fl_o.reshape(-1), fl_r, "eval flow error")
        ev[f"f{f}_uv"], ev[f"f{f}_rig"], ev[f"f{f}_flow"] = uv_o.numpy(), rig_o.numpy(), fl_o.numpy()
    np.savez_compressed(os.path.join(OUT, "eval_maps.npz"), frames=np.array([2, T - 1]),
                        **{f"map{i}": p.detach().numpy() for i, p in enumerate(ref_m.parameters())}, **ev)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
