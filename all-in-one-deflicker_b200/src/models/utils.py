"""The four helpers of the reference's src/models/utils.py that stage-2 inference uses (:55-60,234-247,
600-644): load_image, InputPadder (/32, replicate), tensor2img, save_img."""
import cv2
import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image


def load_image(path, size=None, device=None, resize=True):
    img = np.array(Image.open(path)) / 255.
    img = np.repeat(img[:, :, None], 3, axis=2) if img.ndim == 2 else img[..., :3]
    if size is not None:
        img = cv2.resize(img, size, cv2.INTER_LINEAR)      # the flag lands in `dst`: always bilinear
    orig_h, orig_w = img.shape[:2]
    if resize:
        img = cv2.resize(img, (orig_w // 32 * 32, orig_h // 32 * 32), cv2.INTER_LINEAR)
    img = torch.from_numpy(img).permute(2, 0, 1).unsqueeze(0).to(device, dtype=torch.float)
    return img, (orig_w, orig_h)


class InputPadder:
    """Pads images such that dimensions are divisible by 32 (utils.py:626-644)."""
    def __init__(self, dims, mode='other'):
        self.ht, self.wd = dims[-2:]
        pad_ht = (((self.ht // 32) + 1) * 32 - self.ht) % 32
        pad_wd = (((self.wd // 32) + 1) * 32 - self.wd) % 32
        self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, 0, pad_ht]

    def pad(self, *inputs):
        return [F.pad(x, self._pad, mode='replicate') for x in inputs]

    def unpad(self, x):
        ht, wd = x.shape[-2:]
        return x[..., self._pad[2]:ht - self._pad[3], self._pad[0]:wd - self._pad[1]]


def tensor2img(img_t):
    img = img_t[0].detach().to("cpu").numpy()
    return np.transpose(img, (1, 2, 0))


def save_img(img, filename):
    if img.ndim == 3:
        img = img[:, :, ::-1]                      # RGB -> BGR
    cv2.imwrite(filename, np.clip(img * 255.0, 0, 255).astype(np.uint8), [cv2.IMWRITE_PNG_COMPRESSION, 0])
