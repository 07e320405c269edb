"""N>1 host logic on CPU (gloo, world_size 2 and 4): the frame-sharded formulation — every rank sees the
same index batch, keeps the samples whose frame it owns, normalises by the GLOBAL batch / GLOBAL
flow-row counts, and one SUM all-reduce of (gradients ‖ loss vector) reproduces the unsharded
iteration.  The per-rank arithmetic here is the oracle; the CUDA path implements the same
formulation (tests/test_atlas_gpu.py::test_frame_sharding_is_linear checks it on the GPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200 import atlas as A
from oracle import atlas_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load():
    z = np.load(os.path.join(GOLDEN, "iteration.npz"))
    video = O.Video(**{k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("video_")})
    p = np.load(os.path.join(GOLDEN, "params_seed1234.npz"))
    mp_ = [torch.from_numpy(p[f"map{i}"]).clone().requires_grad_(True) for i in range(12)]
    ap_ = [torch.from_numpy(p[f"atl{i}"]).clone().requires_grad_(True) for i in range(16)]
    return video, mp_, ap_, torch.from_numpy(z["inds"])


def _sharded_terms(video, mp_, ap_, inds, t0, t1):
    """This rank's share of every loss term, normalised by global counts."""
    H, W, T = video.H, video.W, video.T
    table = O.pixel_table(T, H, W)
    t_all = table[2, inds.reshape(-1)]
    local = (t_all >= t0) & (t_all < t1)
    B = inds.shape[0]
    jif_all = table[:, inds]
    wf_all = video.mask_fwd[jif_all[1].squeeze(), jif_all[0].squeeze(), jif_all[2].squeeze(), 0] != 0
    wb_all = video.mask_bwd[jif_all[1].squeeze(), jif_all[0].squeeze(), jif_all[2].squeeze(), 0] != 0
    n_f, n_b = int(wf_all.sum()), int(wb_all.sum())
    sub = inds[local]
    n_loc = sub.shape[0]
    zero = sum(p.sum() * 0 for p in mp_ + ap_)
    if n_loc == 0:
        return {k: zero for k in ("rgb", "gradient", "rigidity", "rigidity_global", "flow")}
    terms = O.iteration_losses(video, mp_, ap_, sub, 0)
    scale = n_loc / B
    out = {k: terms[k] * scale for k in ("rgb", "gradient", "rigidity", "rigidity_global")}
    # flow: the oracle returns 0.5*mean_b + 0.5*mean_f over the local valid rows; rebuild the two means
    jif = table[:, sub]
    larger = max(W, H)
    mapping = lambda x: O.mlp_forward(O.MAPPING_SPEC, mp_, x)
    uv = mapping(O.normalise_xyt(jif, larger, T))
    uvf, xf, _ = O.flow_matches(jif, video.mask_fwd, video.flow_fwd, larger, T, True, uv)
    uvb, xb, _ = O.flow_matches(jif, video.mask_bwd, video.flow_bwd, larger, T, False, uv)
    lf = ((mapping(xf) - uvf).norm(dim=1) * larger / 1.6).sum() / n_f if xf.shape[0] else zero
    lb = ((mapping(xb) - uvb).norm(dim=1) * larger / 1.6).sum() / n_b if xb.shape[0] else zero
    out["flow"] = 0.5 * lf + 0.5 * lb
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    video, mp_, ap_, inds = _load()
    t0, t1 = A.frame_range(rank, world, video.T)
    terms = _sharded_terms(video, mp_, ap_, inds, t0, t1)
    total = terms["rigidity"] + 5.0 * terms["rigidity_global"] + 5000 * terms["rgb"] + 500.0 * terms["flow"] \
        + 1000 * terms["gradient"]
    total.backward()
    flat = torch.cat([p.grad.flatten() if p.grad is not None else torch.zeros(p.numel()) for p in mp_ + ap_]
                     + [torch.stack([total.detach()] + [terms[k].detach() for k in
                                                        ("rgb", "gradient", "rigidity", "rigidity_global", "flow")])])
    dist.all_reduce(flat)                     # ONE collective: gradients ‖ losses
    if rank == 0:
        q.put(flat.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 4])
def test_frame_sharding_reproduces_single_rank(world):
    """world 4 on the 6-frame golden video gives ragged frame blocks (2, 2, 1, 1 frames)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 7 * world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    got = q.get(timeout=240)
    for p in procs: p.join(60)
    assert all(p.exitcode == 0 for p in procs)
    video, mp_, ap_, inds = _load()
    terms = O.iteration_losses(video, mp_, ap_, inds, 0)
    terms["total"].backward()
    ref = torch.cat([p.grad.flatten() for p in mp_ + ap_]).numpy()
    n = ref.size
    scale = np.abs(ref).max()
    assert np.abs(got[:n] - ref).max() <= 2e-4 * scale
    ref_l = [float(terms[k].detach()) for k in ("total", "rgb", "gradient", "rigidity", "rigidity_global", "flow")]
    np.testing.assert_allclose(got[n:], ref_l, rtol=2e-4)
