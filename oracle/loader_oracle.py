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


# ----------------------------------------------------------------------------------------------------------------
# cv2.resize(INTER_LINEAR) for fp32 images when the output is not larger than the input in either dimension — the
# case `resize_flow` (unwrap_utils.py:33-38) meets: RAFT flows are computed at the frame files' resolution and the
# atlas runs at that resolution divided by `--down`.  Pinned bit-exactly against cv2.resize by
# tests/test_loader_oracle.py (random sizes, integer and fractional factors).
#   * sampling position of destination index d: fx = fp32((d + 0.5) * (src / dst) - 0.5) (the product in float64),
#     s = floor(fx), weight = fx - s; positions left of the first / right of the last source sample clamp to it with
#     weight 0
#   * two passes, each result rounded to fp32: rows first  S[s]*(1-a) + S[s+1]*a, then columns  R[s]*(1-b) + R[s+1]*b,
#     products and sums as separate fp32 operations (no fused multiply-add)
#   * exactly half size in BOTH dimensions is OpenCV's "area fast" path: ((a + b) + c) + d) * 0.25 over the 2x2 block
# ----------------------------------------------------------------------------------------------------------------
def _resize_taps(dst: int, src: int):
    scale = np.float64(src) / np.float64(dst)
    f = ((np.arange(dst) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    f[s < 0] = 0; s[s < 0] = 0
    f[s >= src - 1] = 0; s[s >= src - 1] = src - 1
    return s, np.minimum(s + 1, src - 1), f


def resize_bilinear_down(img, neww: int, newh: int):
    """cv2.resize(img, (neww, newh), interpolation=cv2.INTER_LINEAR) for fp32 (H, W, C), newh <= H and neww <= W."""
    img = np.asarray(img, dtype=np.float32)
    h, w = img.shape[:2]
    assert newh <= h and neww <= w, "the restatement covers down-scaling (and identity) only"
    f32 = lambda a: a.astype(np.float32)
    if h == 2 * newh and w == 2 * neww:
        a, b, c, d = img[0::2, 0::2], img[0::2, 1::2], img[1::2, 0::2], img[1::2, 1::2]
        return f32(f32(f32(f32(a + b) + c) + d) * np.float32(0.25))
    sx, sx1, fx = _resize_taps(neww, w)
    sy, sy1, fy = _resize_taps(newh, h)
    a0, a1 = f32(np.float32(1) - fx)[None, :, None], fx[None, :, None]
    b0, b1 = f32(np.float32(1) - fy)[:, None, None], fy[:, None, None]
    rows = f32(f32(img[:, sx] * a0) + f32(img[:, sx1] * a1))
    return f32(f32(rows[sy] * b0) + f32(rows[sy1] * b1))


def resize_flow(flow, newh: int, neww: int):
    """`resize_flow` of unwrap_utils.py:33-38, including its swapped scale factors (x by newh/oldh, y by neww/oldw)."""
    oldh, oldw = flow.shape[:2]
    out = resize_bilinear_down(flow, neww, newh).copy()
    out[:, :, 0] *= np.float32(newh / oldh)
    out[:, :, 1] *= np.float32(neww / oldw)
    return out
