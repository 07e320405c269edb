"""`RAFTWrapper` with the reference's interface (src/models/stage_1/raft_wrapper.py:16-73): loads a
`DataParallel`-style checkpoint (`module.` prefixed keys), /8 padding, 20 refinement iterations."""
import argparse
import sys

import cv2
import numpy as np
import torch
from PIL import Image

from src.models.stage_1.core.raft import RAFT
from src.models.stage_1.core.utils.utils import InputPadder

device = torch.device("cuda:0")


class RAFTWrapper():
    def __init__(self, model_path, max_long_edge=900):
        args = argparse.Namespace()
        args.small, args.mixed_precision = False, True
        args.model, args.max_long_edge = model_path, max_long_edge
        self.model = RAFT(args)
        if model_path is not None:
            sd = torch.load(model_path, map_location="cpu")
            self.model.load_state_dict({k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()})
        self.model.to(device).eval()
        self.args = args

    def load_image(self, fn):
        img = np.array(Image.open(fn)).astype(np.uint8)
        im_h, im_w = img.shape[0], img.shape[1]
        factor = max(im_w, im_h) / self.args.max_long_edge
        if factor > 1:
            img = cv2.resize(img, (int(im_w // factor), int(im_h // factor)), interpolation=cv2.INTER_AREA)
        return torch.from_numpy(img).permute(2, 0, 1).float()

    def load_images(self, fn1, fn2):
        images = torch.stack([self.load_image(f) for f in sorted([fn1, fn2])], dim=0).to(device)
        images = InputPadder(images.shape).pad(images)[0]
        return images[0, None], images[1, None]

    def compute_flow(self, im1, im2):
        padder = InputPadder(im1.shape)
        im1, im2 = padder.pad(im1, im2)
        _, flow12 = self.model(im1, im2, iters=20, test_mode=True)
        return flow12[0].permute(1, 2, 0).detach().cpu().numpy()

    def compute_flow_both(self, im1, im2):
        """(flow 1->2, flow 2->1), identical to two compute_flow calls but with the feature encoder run once."""
        padder = InputPadder(im1.shape)
        im1, im2 = padder.pad(im1, im2)
        (_, f12), (_, f21) = self.model.forward_both(im1, im2, iters=20)
        return tuple(f[0].permute(1, 2, 0).detach().cpu().numpy() for f in (f12, f21))
