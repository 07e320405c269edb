"""Shared by the RAFT / stage-2 tests: deterministic weights regenerated from (name, shape) lists."""
import torch


def seeded_weights(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in shapes:
        fan = 1
        for d in shape[1:]:
            fan *= d
        scale = (1.0 / max(fan, 1)) ** 0.5
        if name.endswith("running_var"):
            out[name] = torch.rand(*shape, generator=g) + 0.5
        else:
            out[name] = (torch.rand(*shape, generator=g) * 2 - 1) * scale if len(shape) else torch.zeros(())
    return out
