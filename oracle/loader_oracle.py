"""CPU restatement (oracle) of the arithmetic inside the reference's stage-1 input producer that a device-side
producer has to reproduce: the forward/backward flow-consistency mask of
src/models/stage_1/unwrap_utils.py:10-23, i.e. `cv2.remap(flow21, grid + flow12, INTER_LINEAR)` with the
default constant-zero border.

TEST INFRASTRUCTURE ONLY (see oracle/atlas_oracle.py): nothing in the product imports this module.

Third-party arithmetic: OpenCV `remap` (pinned 4.7.0.72 in the reference's environment.yml:141, 4.13 here).  Its
bilinear path quantises the sampling coordinates to 1/32 pixel (INTER_BITS = 5): s = round_half_even(32 * x),
integer part s >> 5, fraction (s & 31) / 32; the four weights are the fp32 products (1-fy)(1-fx), (1-fy)fx,
fy(1-fx), fy*fx, each tap outside the image contributes 0, and the taps are accumulated in fp32 in the order
nw, ne, sw, se.  tests/test_loader_oracle.py pins this restatement bit-exactly against cv2.remap itself."""
import numpy as np

INTER_BITS = 5
TAB = 1 << INTER_BITS


def remap_bilinear_zero(img, map_x, map_y):
    """cv2.remap(img, (map_x, map_y), INTER_LINEAR, BORDER_CONSTANT 0) for fp32 images (H, W, C)."""
    img = np.asarray(img, dtype=np.float32)
    H, W = img.shape[:2]
    sx = np.rint(np.asarray(map_x, dtype=np.float64) * TAB).astype(np.int64)     # cvRound: half to even
    sy = np.rint(np.asarray(map_y, dtype=np.float64) * TAB).astype(np.int64)
    ix, iy = sx >> INTER_BITS, sy >> INTER_BITS
    fx = (sx & (TAB - 1)).astype(np.float32) / np.float32(TAB)
    fy = (sy & (TAB - 1)).astype(np.float32) / np.float32(TAB)
    one = np.float32(1)

    def tap(y, x):
        inside = (x >= 0) & (x < W) & (y >= 0) & (y < H)
        v = img[np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)]
        return np.where(inside[..., None], v, np.float32(0))

    w_nw, w_ne = ((one - fy) * (one - fx))[..., None], ((one - fy) * fx)[..., None]
    w_sw, w_se = (fy * (one - fx))[..., None], (fy * fx)[..., None]
    out = tap(iy, ix) * w_nw + tap(iy, ix + 1) * w_ne + tap(iy + 1, ix) * w_sw + tap(iy + 1, ix + 1) * w_se
    return out.astype(np.float32)


def consistency_error(flow12, flow21):
    """|flow12 + flow21 sampled at (p + flow12)| per pixel (unwrap_utils.py:10-14); the mask is `< 1.0`."""
    flow12 = np.asarray(flow12, dtype=np.float32)
    H, W = flow12.shape[:2]
    gx = flow12[:, :, 0] + np.arange(W, dtype=np.float32)[None, :]
    gy = flow12[:, :, 1] + np.arange(H, dtype=np.float32)[:, None]
    d = flow12 + remap_bilinear_zero(flow21, gx, gy)
    return (d[:, :, 0] ** 2 + d[:, :, 1] ** 2) ** np.float32(.5)
