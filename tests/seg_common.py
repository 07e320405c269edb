"""Shared loading of the segmentation-variant fixture (tests/golden/seg_iteration.npz)."""
import os

import numpy as np
import torch

from oracle import atlas_oracle as O
from oracle import seg_oracle as S

ORDER = ("mapping1", "mapping2", "atlas", "alpha")


def load_fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "seg_iteration.npz"))
    video = O.Video(**{k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("video_")})
    masks = torch.from_numpy(z["masks"])
    torch.manual_seed(int(z["init_seed"]))
    nets = S.init_nets()
    for k in ORDER:
        assert abs(float(sum(p.double().sum() for p in nets[k])) - float(z[f"init_{k}_sum"])) < 1e-9
    return z, video, masks, nets
