"""oracle/seg_oracle.py replayed against the fixture frozen from the reference's segmentation-variant
functions by tests/golden/make_golden_seg.py.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import seg_oracle as S
from seg_common import ORDER, load_fixture


@pytest.fixture(autouse=True)
def _single_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def test_specs():
    assert S.MAPPING2_SPEC.num_params() == 3 * 256 + 256 + 2 * (256 * 256 + 256) + 256 * 2 + 2
    assert S.ALPHA_SPEC.enc_dim == 30 and S.ALPHA_SPEC.layer_dims()[-1] == (256, 1)


@pytest.mark.parametrize("it", [0, 6000, 10001])
def test_seg_iteration_losses_and_grads(golden_dir, it):
    z, video, masks, nets = load_fixture(golden_dir)
    mine = {k: [p.clone().requires_grad_(True) for p in nets[k]] for k in ORDER}
    terms = S.seg_iteration_losses(video, masks, mine, torch.from_numpy(z["inds"]), it)
    terms["total"].backward()
    tag = f"it{it}_"
    for k, v in terms.items():
        np.testing.assert_allclose(float(v.detach()), float(z[tag + "loss_" + k]), rtol=2e-5, err_msg=k)
    assert ("rigidity_global1" in terms) == (it <= 5000)
    for k in ORDER:
        for i, p in enumerate(mine[k]):
            g = p.grad.flatten()
            scale = float(z[tag + f"grad_{k}_{i}_abs"]) + 1e-30
            assert abs(float(g.double().sum()) - float(z[tag + f"grad_{k}_{i}_sum"])) <= 2e-4 * scale, (k, i)
            np.testing.assert_allclose(g[:32].numpy(), z[tag + f"grad_{k}_{i}_head"], rtol=2e-3,
                                       atol=2e-5 * scale / g.numel() + 1e-9, err_msg=f"{k}[{i}]")


def test_seg_trajectory_and_render(golden_dir):
    z, video, masks, nets = load_fixture(golden_dir)
    mine = {k: [p.clone().requires_grad_(True) for p in nets[k]] for k in ORDER}
    opt = S.make_optimizer(mine)
    keys = [str(k) for k in z["traj_keys"]]
    for it in range(3):
        terms = S.seg_iteration_losses(video, masks, mine, torch.from_numpy(z["traj_inds"][it]), it)
        opt.zero_grad()
        terms["total"].backward()
        opt.step()
        np.testing.assert_allclose([float(terms[k].detach()) for k in keys], z["traj_losses"][it], rtol=5e-5)
    for k in ORDER:
        np.testing.assert_allclose(mine[k][0].detach().flatten()[:64].numpy(), z[f"traj_{k}_head"], rtol=0, atol=2e-6)
    img, alpha = S.render_frame_seg({k: [p.detach() for p in mine[k]] for k in ORDER}, int(z["render_frame"]),
                                    video.H, video.W, video.T)
    np.testing.assert_allclose(img.numpy(), z["render_img"], atol=1e-5)
    np.testing.assert_allclose(alpha.numpy(), z["render_alpha"], atol=1e-5)
