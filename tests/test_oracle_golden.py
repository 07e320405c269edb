"""The oracle (oracle/atlas_oracle.py) replayed against the fixtures frozen from the
reference's own modules by tests/golden/make_golden.py.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import atlas_oracle as O


def _params(golden_dir):
    z = np.load(os.path.join(golden_dir, "params_seed1234.npz"))
    mp = [torch.from_numpy(z[f"map{i}"]) for i in range(12)]
    ap = [torch.from_numpy(z[f"atl{i}"]) for i in range(16)]
    return mp, ap


@pytest.fixture(autouse=True)
def _single_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)      # fixtures were frozen with 1 thread (addmm summation order)
    yield
    torch.set_num_threads(n)


def test_init_matches_nn_linear_stream(golden_dir):
    mp, ap = _params(golden_dir)
    torch.manual_seed(1234)
    for p, q in zip(O.init_mlp(O.MAPPING_SPEC) + O.init_mlp(O.ATLAS_SPEC), mp + ap):
        assert torch.equal(p, q)
    assert O.MAPPING_SPEC.num_params() == 264706
    assert O.ATLAS_SPEC.num_params() == 416379
    assert [tuple(p.shape) for p in ap[8:10]] == [(256, 296), (256,)]


def test_forward_and_pe(golden_dir):
    mp, ap = _params(golden_dir)
    z = np.load(os.path.join(golden_dir, "mlp_forward.npz"))
    x3, x2 = torch.from_numpy(z["x3"]), torch.from_numpy(z["x2"])
    assert torch.equal(O.positional_encoding(x2, O.pe_frequencies(O.ATLAS_SPEC)), torch.from_numpy(z["pe"]))
    torch.testing.assert_close(O.mlp_forward(O.MAPPING_SPEC, mp, x3), torch.from_numpy(z["y_map"]), rtol=0, atol=2e-6)
    torch.testing.assert_close(O.mlp_forward(O.ATLAS_SPEC, ap, x2), torch.from_numpy(z["y_atl"]), rtol=0, atol=2e-6)


def test_coordinate_normalisation(golden_dir):
    z = np.load(os.path.join(golden_dir, "coords.npz"))
    ints = torch.from_numpy(z["ints"])
    for L in (768, 432, 256, 160, 90, 25):
        got = ints / O._half(L) - 1
        assert got.dtype == torch.float32 and np.array_equal(got.numpy(), z[f"L{L}"])
        # the rule the CUDA kernels implement: fp32 divide, then fp32 subtract
        manual = ints.float().numpy() / np.float32(L / 2) - np.float32(1)
        assert np.array_equal(manual.astype(np.float32), z[f"L{L}"])
    for T in (80, 16, 7):
        assert np.array_equal((ints / (T / 2.0) - 1).numpy(), z[f"T{T}"])


@pytest.mark.parametrize("it", [0, 6000])
def test_iteration_losses_and_grads(golden_dir, it):
    mp, ap = _params(golden_dir)
    z = np.load(os.path.join(golden_dir, "iteration.npz"))
    video = O.Video(**{k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("video_")})
    assert torch.equal(O.pixel_table(int(z["T"]), int(z["H"]), int(z["W"]))[:, 1000], torch.tensor([0, 1, 1]))
    mp = [p.clone().requires_grad_(True) for p in mp]
    ap = [p.clone().requires_grad_(True) for p in ap]
    terms = O.iteration_losses(video, mp, ap, torch.from_numpy(z["inds"]), it)
    terms["total"].backward()
    tag = f"it{it}_"
    for k, v in terms.items():
        np.testing.assert_allclose(float(v.detach()), float(z[tag + "loss_" + k]), rtol=2e-5, err_msg=k)
    assert ("rigidity_global" in terms) == (it <= 5000)
    for i, p in enumerate(mp + ap):
        g = p.grad.flatten()
        scale = float(z[tag + f"grad{i}_abs"]) + 1e-30
        assert abs(float(g.double().sum()) - float(z[tag + f"grad{i}_sum"])) <= 2e-4 * scale
        np.testing.assert_allclose(g[:32].numpy(), z[tag + f"grad{i}_head"], rtol=2e-3,
                                   atol=2e-5 * scale / g.numel() + 1e-9)


def test_three_adam_steps(golden_dir):
    mp, ap = _params(golden_dir)
    z = np.load(os.path.join(golden_dir, "iteration.npz"))
    video = O.Video(**{k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("video_")})
    t = np.load(os.path.join(golden_dir, "trajectory.npz"))
    mp = [p.clone().requires_grad_(True) for p in mp]
    ap = [p.clone().requires_grad_(True) for p in ap]
    opt = O.make_optimizer(mp, ap)
    for it in range(3):
        out = O.train_iteration(video, mp, ap, opt, torch.from_numpy(t["inds"][it]), it)
        got = [out[k] for k in ("total", "rgb", "gradient", "rigidity", "rigidity_global", "flow")]
        np.testing.assert_allclose(got, t["losses"][it], rtol=5e-4)
    np.testing.assert_allclose(mp[0].detach().flatten()[:64].numpy(), t["map0_head"], atol=5e-6)


def test_pretrain_step(golden_dir):
    z = np.load(os.path.join(golden_dir, "pretrain.npz"))
    torch.manual_seed(99)
    mp = [p.requires_grad_(True) for p in O.init_mlp(O.MAPPING_SPEC)]
    torch.manual_seed(5)
    opt = torch.optim.Adam(mp, lr=1e-4)
    H, W, T = int(z["H"]), int(z["W"]), int(z["T"])
    for f in range(T):
        ys = torch.randint(H, (10000, 1))
        xs = torch.randint(W, (10000, 1))
        loss = O.pretrain_losses(mp, f, ys, xs, T, max(W, H), 0.8)
        opt.zero_grad(); loss.backward(); opt.step()
        np.testing.assert_allclose(float(loss.detach()), z["losses"][f], rtol=1e-5)
    np.testing.assert_allclose(mp[0].detach().flatten()[:64].numpy(), z["w0_head"], atol=2e-6)


def test_render_and_psnr(golden_dir):
    mp, ap = _params(golden_dir)
    z = np.load(os.path.join(golden_dir, "render.npz"))
    it = np.load(os.path.join(golden_dir, "iteration.npz"))
    img = O.render_frame(mp, ap, int(z["frame"]), int(it["H"]), int(it["W"]), int(it["T"]))
    np.testing.assert_allclose(img.numpy(), z["img"], atol=2e-6)
    assert np.abs(O.to_uint8(img).astype(int) - z["u8"].astype(int)).max() <= 1
    frames = torch.from_numpy(it["video_frames"])
    assert abs(O.psnr(frames[:, :, :, int(z["frame"])], img) - float(z["psnr"])) < 1e-3
