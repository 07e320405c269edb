"""Seeded synthetic inputs for the stage-1 atlas loop.

Two sets, both produced in the *reference* CPU layouts (T innermost,
src/models/stage_1/unwrap_utils.py:112-122) so that the same tensors feed the oracle and,
after `VideoDev.from_reference_layout`, the CUDA path:

* throughput set — i.i.d. noise; what `bench.py` times (per-iteration cost does not depend on
  image content);
* quality set — a smooth texture translating at constant velocity with per-frame flicker gain /
  offset, exact flow, consistency masks: used for PSNR / trajectory parity.
"""
from __future__ import annotations

import numpy as np
import torch


def _differences(frames: torch.Tensor):
    """Forward differences exactly as the reference loader builds them
    (unwrap_utils.py:132-133): zero in the last column / row."""
    dx = torch.zeros_like(frames)
    dy = torch.zeros_like(frames)
    dy[:-1] = frames[1:] - frames[:-1]
    dx[:, :-1] = frames[:, 1:] - frames[:, :-1]
    return dx, dy


def throughput_set(H: int, W: int, T: int, seed: int = 0, mask_density: float = 0.7) -> dict:
    g = torch.Generator().manual_seed(seed)
    frames = torch.rand(H, W, 3, T, generator=g)
    dx, dy = _differences(frames)
    flow_fwd = torch.randn(H, W, 2, T, 1, generator=g)
    flow_bwd = torch.randn(H, W, 2, T, 1, generator=g)
    mask_fwd = (torch.rand(H, W, T, 1, generator=g) < mask_density).float()
    mask_bwd = (torch.rand(H, W, T, 1, generator=g) < mask_density).float()
    # the loader never fills the forward slot of the last frame nor the backward slot of the
    # first one (unwrap_utils.py:135,152-159)
    flow_fwd[:, :, :, T - 1] = 0
    mask_fwd[:, :, T - 1] = 0
    flow_bwd[:, :, :, 0] = 0
    mask_bwd[:, :, 0] = 0
    return dict(frames=frames, frames_dx=dx, frames_dy=dy, flow_fwd=flow_fwd, flow_bwd=flow_bwd,
                mask_fwd=mask_fwd, mask_bwd=mask_bwd)


def _smooth_texture(h: int, w: int, seed: int) -> np.ndarray:
    """Low-pass seeded noise, 3 channels, values in [0.1, 0.9], periodic in both axes."""
    rng = np.random.default_rng(seed)
    spec = rng.standard_normal((h, w, 3)) + 1j * rng.standard_normal((h, w, 3))
    fy = np.fft.fftfreq(h)[:, None, None]
    fx = np.fft.fftfreq(w)[None, :, None]
    spec *= np.exp(-((fy ** 2 + fx ** 2) * (24.0 ** 2)) * 40.0)
    tex = np.real(np.fft.ifft2(spec, axes=(0, 1)))
    tex -= tex.min(axis=(0, 1), keepdims=True)
    tex /= tex.max(axis=(0, 1), keepdims=True)
    return (0.1 + 0.8 * tex).astype(np.float64)


def _bilinear_periodic(tex: np.ndarray, ys: np.ndarray, xs: np.ndarray) -> np.ndarray:
    h, w, _ = tex.shape
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    fy = (ys - y0)[..., None]
    fx = (xs - x0)[..., None]
    y0 %= h; x0 %= w
    y1 = (y0 + 1) % h
    x1 = (x0 + 1) % w
    return (tex[y0, x0] * (1 - fy) * (1 - fx) + tex[y0, x1] * (1 - fy) * fx
            + tex[y1, x0] * fy * (1 - fx) + tex[y1, x1] * fy * fx)


def quality_set(H: int, W: int, T: int, seed: int = 0, velocity=(1.5, -0.75)) -> dict:
    """Flickering translation.  Pixel (x, y) of frame t shows texture(x - vx t, y - vy t), so the
    content at (x, y, t) reappears at (x + vx, y + vy, t + 1): flow_fwd = (+vx, +vy),
    flow_bwd = (-vx, -vy).  Masks: forward/backward consistency of these exact flows is 0, so all
    pixels whose match stays inside the frame are valid (what `compute_consistency` <1 px yields,
    unwrap_utils.py:10-14,148-149, up to the remap border)."""
    vx, vy = velocity
    rng = np.random.default_rng(seed + 1)
    tex = _smooth_texture(2 * H, 2 * W, seed)
    gain = rng.uniform(0.8, 1.2, T)
    offs = rng.uniform(-0.05, 0.05, T)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    frames = np.zeros((H, W, 3, T), np.float32)
    clean = np.zeros((H, W, 3, T), np.float32)
    for t in range(T):
        img = _bilinear_periodic(tex, yy - vy * t + H / 2, xx - vx * t + W / 2)
        clean[..., t] = img
        frames[..., t] = np.clip(img * gain[t] + offs[t], 0.0, 1.0)
    frames = torch.from_numpy(frames)
    dx, dy = _differences(frames)
    flow_fwd = torch.zeros(H, W, 2, T, 1)
    flow_bwd = torch.zeros(H, W, 2, T, 1)
    mask_fwd = torch.zeros(H, W, T, 1)
    mask_bwd = torch.zeros(H, W, T, 1)
    inside_f = torch.from_numpy(((xx + vx >= 0) & (xx + vx <= W - 1) & (yy + vy >= 0) & (yy + vy <= H - 1)))
    inside_b = torch.from_numpy(((xx - vx >= 0) & (xx - vx <= W - 1) & (yy - vy >= 0) & (yy - vy <= H - 1)))
    for t in range(T - 1):
        flow_fwd[:, :, 0, t, 0] = vx
        flow_fwd[:, :, 1, t, 0] = vy
        mask_fwd[:, :, t, 0] = inside_f.float()
        flow_bwd[:, :, 0, t + 1, 0] = -vx
        flow_bwd[:, :, 1, t + 1, 0] = -vy
        mask_bwd[:, :, t + 1, 0] = inside_b.float()
    return dict(frames=frames, frames_dx=dx, frames_dy=dy, flow_fwd=flow_fwd, flow_bwd=flow_bwd,
                mask_fwd=mask_fwd, mask_bwd=mask_bwd, clean=torch.from_numpy(clean))
