"""The three loss functions of the drop-in boundary (all-in-one-deflicker_b200/src/models/stage_1/loss_utils.py,
signatures of the reference's loss_utils.py:134,227,299) against the oracle: same CPU video tensors, same `jif`,
the repo's IMLP objects on the GPU versus the oracle networks with the same parameters.

Tolerances: loss value rtol 2e-5 (fp32 sums in a different order); gradient of the returned scalar with respect
to every network parameter |err| <= 2e-4 * max|g| per tensor + 5e-4 * max|g| over the network + 2e-7 (fp32 path of the IMLP kernels; the absolute
floor covers tensors whose gradient is a near-cancelling sum: the three per-sample terms of the image-gradient loss
add up to exactly zero before tanh', so the last atlas bias gradient is ~1e-5 made of ~1e-3 addends)."""
import os

import numpy as np
import pytest
import torch

from b200 import synth
from oracle import atlas_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _fp32_imlp(monkeypatch):
    """These tests pin the loss-head arithmetic with tight bounds: run the IMLP objects on the fp32 CUDA-core kernels
    (the tensor-core IMLP has its own test, tests/test_tc_gpu.py::test_imlp_class_on_tensor_cores)."""
    monkeypatch.setenv("B200_IMLP_PRECISION", "fp32")


def _nets(golden_dir):
    from src.models.stage_1.implicit_neural_networks import IMLP
    z = np.load(os.path.join(golden_dir, "params_seed1234.npz"))
    mp = [torch.from_numpy(z[f"map{i}"]).clone().requires_grad_(True) for i in range(12)]
    ap = [torch.from_numpy(z[f"atl{i}"]).clone().requires_grad_(True) for i in range(16)]
    m = IMLP(input_dim=3, output_dim=2, hidden_dim=256, use_positional=False, positional_dim=4, num_layers=6,
             skip_layers=[], verbose=False)
    a = IMLP(input_dim=2, output_dim=3, hidden_dim=256, use_positional=True, positional_dim=10, num_layers=8,
             skip_layers=[4, 7], verbose=False)
    m.load_state_dict(O.state_dict_of([p.detach() for p in mp]))
    a.load_state_dict(O.state_dict_of([p.detach() for p in ap]))
    return mp, ap, m.to(DEV), a.to(DEV)


def _setup(golden_dir, B=700, mask_density=0.7):
    H, W, T = 40, 56, 7
    data = synth.throughput_set(H, W, T, seed=11, mask_density=mask_density)
    video = O.Video(**data)
    inds = torch.randint(H * W * T, (B, 1), generator=torch.Generator().manual_seed(5))
    jif = O.pixel_table(T, H, W)[:, inds]
    mp, ap, m, a = _nets(golden_dir)
    o_map = lambda x: O.mlp_forward(O.MAPPING_SPEC, mp, x)
    o_atl = lambda x: O.mlp_forward(O.ATLAS_SPEC, ap, x)
    xyt = O.normalise_xyt(jif, max(H, W), T)
    return data, video, jif, xyt, (mp, ap, o_map, o_atl), (m, a), (H, W, T)


def _check_grads(ref_params, module, what):
    views = module._views(module.flat.grad)
    # network-level gradient scale: the floor for tensors whose gradient is a near-cancelling sum (the output-layer
    # bias cancels exactly between the two evaluations that every one of these losses subtracts)
    scale = max(float(p.grad.abs().max()) for p in ref_params if p.grad is not None)
    for i in range(len(ref_params) // 2):
        for kind, p in (("weight", ref_params[2 * i]), ("bias", ref_params[2 * i + 1])):
            got = views[f"hidden.{i}.{kind}"].cpu()
            want = p.grad if p.grad is not None else torch.zeros_like(p)
            tol = 2e-4 * float(want.abs().max()) + 5e-4 * scale + 2e-7
            assert float((got - want).abs().max()) <= tol, (what, i, kind, float((got - want).abs().max()), tol)


def test_gradient_loss_single(golden_dir):
    from src.models.stage_1 import loss_utils as LU
    data, video, jif, xyt, (mp, ap, o_map, o_atl), (m, a), (H, W, T) = _setup(golden_dir)
    rgb_o = (o_atl(o_map(xyt) * 0.5 + 0.5) + 1.0) * 0.5
    want = O.gradient_loss(video, jif, o_map, o_atl, rgb_o, W)
    want.backward()
    rgb = (a(m(xyt.to(DEV)) * 0.5 + 0.5) + 1.0) * 0.5
    got = LU.get_gradient_loss_single(data["frames_dx"], data["frames_dy"], jif, m, a, rgb, DEV, W, T)
    assert got.dim() == 0 and got.dtype == torch.float32 and got.requires_grad
    np.testing.assert_allclose(float(got), float(want), rtol=2e-5)
    got.backward()
    _check_grads(mp, m, "gradient/mapping")
    _check_grads(ap, a, "gradient/atlas")


@pytest.mark.parametrize("d", [1, 100])
def test_rigidity_loss(golden_dir, d):
    from src.models.stage_1 import loss_utils as LU
    data, video, jif, xyt, (mp, ap, o_map, o_atl), (m, a), (H, W, T) = _setup(golden_dir)
    L = max(H, W)
    want = O.rigidity_loss(jif, d, L, T, o_map, o_map(xyt), uv_scale=0.8)
    want.backward()
    got = LU.get_rigidity_loss(jif, d, L, T, m, m(xyt.to(DEV)), DEV, uv_mapping_scale=0.8)
    np.testing.assert_allclose(float(got), float(want), rtol=2e-5)
    got.backward()
    _check_grads(mp, m, f"rigidity d={d}")
    with torch.no_grad():
        every = LU.get_rigidity_loss(jif, d, L, T, m, m(xyt.to(DEV)), DEV, uv_mapping_scale=0.8, return_all=True)
        ref = O.rigidity_loss(jif, d, L, T, o_map, o_map(xyt), uv_scale=0.8, per_sample=True)
    np.testing.assert_allclose(every.cpu().numpy(), ref.detach().numpy(), rtol=2e-4)


def test_optical_flow_loss(golden_dir):
    from src.models.stage_1 import loss_utils as LU
    data, video, jif, xyt, (mp, ap, o_map, o_atl), (m, a), (H, W, T) = _setup(golden_dir)
    L = max(H, W)
    want = O.flow_loss(video, jif, o_map(xyt), L, o_map, 0.8)
    want.backward()
    got = LU.get_optical_flow_loss(jif, m(xyt.to(DEV)), data["flow_bwd"], data["mask_bwd"], L, T, m, data["flow_fwd"],
                                   data["mask_fwd"], 0.8, DEV)
    np.testing.assert_allclose(float(got), float(want), rtol=2e-5)
    got.backward()
    _check_grads(mp, m, "flow")


def test_optical_flow_loss_empty_set_is_nan(golden_dir):
    from src.models.stage_1 import loss_utils as LU
    data, video, jif, xyt, _, (m, a), (H, W, T) = _setup(golden_dir, B=64, mask_density=0.0)
    got = LU.get_optical_flow_loss(jif, m(xyt.to(DEV)), data["flow_bwd"], data["mask_bwd"], max(H, W), T, m,
                                   data["flow_fwd"], data["mask_fwd"], 0.8, DEV)
    assert torch.isnan(got)                      # mean over an empty set, as the reference (loss_utils.py:320-322)


# ---------------------------------------------------------------------------------------------------------------
# segmentation-variant functions (reference loss_utils.py:173-224, :299-322 with use_alpha=True, :385-408)
# ---------------------------------------------------------------------------------------------------------------
def _seg_setup(golden_dir, B=64):
    from oracle import seg_oracle as S
    from seg_common import ORDER, load_fixture
    from src.models.stage_1.implicit_neural_networks import IMLP
    z, video, masks, nets = load_fixture(golden_dir)
    ref = {k: [p.clone().requires_grad_(True) for p in nets[k]] for k in ORDER}
    mods = dict(
        mapping1=IMLP(input_dim=3, output_dim=2, hidden_dim=256, use_positional=False, positional_dim=4, num_layers=6,
                      skip_layers=[], verbose=False),
        mapping2=IMLP(input_dim=3, output_dim=2, hidden_dim=256, use_positional=False, positional_dim=2, num_layers=4,
                      skip_layers=[], verbose=False),
        atlas=IMLP(input_dim=2, output_dim=3, hidden_dim=256, use_positional=True, positional_dim=10, num_layers=8,
                   skip_layers=[4, 7], verbose=False),
        alpha=IMLP(input_dim=3, output_dim=1, hidden_dim=256, use_positional=True, positional_dim=5, num_layers=8,
                   skip_layers=[], verbose=False))
    for k in ORDER:
        mods[k].load_state_dict(O.state_dict_of(nets[k]))
        mods[k] = mods[k].to(DEV)
    specs = dict(mapping1=S.MAPPING1_SPEC, mapping2=S.MAPPING2_SPEC, alpha=S.ALPHA_SPEC, atlas=S.ATLAS_SPEC)
    onet = {k: (lambda x, k=k: O.mlp_forward(specs[k], ref[k], x)) for k in ORDER}
    inds = torch.from_numpy(z["inds"])
    jif = O.pixel_table(video.T, video.H, video.W)[:, inds]
    xyt = O.normalise_xyt(jif, max(video.H, video.W), video.T)
    return S, video, jif, xyt, ref, onet, mods


def test_seg_gradient_loss(golden_dir):
    from src.models.stage_1 import loss_utils as LU
    S, video, jif, xyt, ref, onet, mods = _seg_setup(golden_dir)
    a_o = S.alpha_of(onet["alpha"](xyt))
    out_o = (onet["atlas"](onet["mapping1"](xyt) * 0.5 + 0.5) + 1.0) * 0.5 * a_o \
        + (onet["atlas"](onet["mapping2"](xyt) * 0.5 - 0.5) + 1.0) * 0.5 * (1.0 - a_o)
    want = S.gradient_loss_seg(video, jif, onet["mapping1"], onet["mapping2"], onet["atlas"], onet["alpha"], out_o, video.W)
    want.backward()
    x = xyt.to(DEV)
    a = LU._alpha_of(mods["alpha"](x))
    out = (mods["atlas"](mods["mapping1"](x) * 0.5 + 0.5) + 1.0) * 0.5 * a \
        + (mods["atlas"](mods["mapping2"](x) * 0.5 - 0.5) + 1.0) * 0.5 * (1.0 - a)
    got = LU.get_gradient_loss(video.frames_dx, video.frames_dy, jif, mods["mapping1"], mods["mapping2"], mods["atlas"],
                               out, DEV, video.W, video.T, mods["alpha"])
    np.testing.assert_allclose(float(got), float(want), rtol=2e-5)
    got.backward()
    for k in ("mapping1", "mapping2", "atlas", "alpha"):
        _check_grads(ref[k], mods[k], f"seg gradient/{k}")


def test_seg_alpha_weighted_flow_and_alpha_flow(golden_dir):
    from src.models.stage_1 import loss_utils as LU
    S, video, jif, xyt, ref, onet, mods = _seg_setup(golden_dir)
    L = max(video.H, video.W)
    a_o = S.alpha_of(onet["alpha"](xyt))
    want = S.flow_loss_alpha(video, jif, onet["mapping1"](xyt), L, onet["mapping1"], 0.8, a_o) \
        + S.flow_loss_alpha(video, jif, onet["mapping2"](xyt), L, onet["mapping2"], 0.8, 1 - a_o) \
        + 10.0 * S.flow_alpha_loss(video, jif, a_o, L, onet["alpha"])
    want.backward()
    x = xyt.to(DEV)
    a = LU._alpha_of(mods["alpha"](x))
    f1 = LU.get_optical_flow_loss(jif, mods["mapping1"](x), video.flow_bwd, video.mask_bwd, L, video.T, mods["mapping1"],
                                  video.flow_fwd, video.mask_fwd, 0.8, DEV, use_alpha=True, alpha=a)
    f2 = LU.get_optical_flow_loss(jif, mods["mapping2"](x), video.flow_bwd, video.mask_bwd, L, video.T, mods["mapping2"],
                                  video.flow_fwd, video.mask_fwd, 0.8, DEV, use_alpha=True, alpha=1 - a)
    fa = LU.get_optical_flow_alpha_loss(mods["alpha"], jif, a, video.flow_bwd, video.mask_bwd, L, video.T,
                                        video.flow_fwd, video.mask_fwd, DEV)
    got = f1 + f2 + 10.0 * fa
    np.testing.assert_allclose(float(got), float(want), rtol=2e-5)
    got.backward()
    for k in ("mapping1", "mapping2", "alpha"):
        _check_grads(ref[k], mods[k], f"seg flow/{k}")
