"""Inference helpers with the names the reference's RAFT code imports from core/utils/utils.py:
`InputPadder` (replicate-pad H and W up to multiples of 8) and `coords_grid`."""
import torch
import torch.nn.functional as F


class InputPadder:
    """Same contract as core/utils/utils.py:7-24: 'sintel' splits the padding evenly between both sides of each
    axis, any other mode puts all of the vertical padding at the bottom."""

    def __init__(self, dims, mode='sintel'):
        self.ht, self.wd = int(dims[-2]), int(dims[-1])
        extra_h, extra_w = -self.ht % 8, -self.wd % 8
        left, right = extra_w // 2, extra_w - extra_w // 2
        top, bottom = (extra_h // 2, extra_h - extra_h // 2) if mode == 'sintel' else (0, extra_h)
        self._pad = [left, right, top, bottom]              # F.pad order: last dimension first

    def pad(self, *inputs):
        return [F.pad(t, self._pad, mode='replicate') for t in inputs]

    def unpad(self, x):
        left, right, top, bottom = self._pad
        return x[..., top:x.shape[-2] - bottom, left:x.shape[-1] - right]


def coords_grid(batch, ht, wd):
    """(batch, 2, ht, wd) float grid, channel 0 = x, channel 1 = y."""
    xs = torch.arange(wd, dtype=torch.float32).view(1, wd).expand(ht, wd)
    ys = torch.arange(ht, dtype=torch.float32).view(ht, 1).expand(ht, wd)
    return torch.stack((xs, ys), dim=0).unsqueeze(0).repeat(batch, 1, 1, 1)
