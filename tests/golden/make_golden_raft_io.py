"""Pins RAFTWrapper.load_image / load_image_list / load_images of this repo against the reference's
src/models/stage_1/raft_wrapper.py:29-63 (image decode, long-edge area downsampling, /8 replicate padding) and
freezes tests/golden/raft_io.npz.  Run ONLY in the build container:   python tests/golden/make_golden_raft_io.py
Each implementation is imported in its own interpreter (both trees are packages called `src`)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.dirname(os.path.abspath(__file__))

WORKER = r"""
import sys, types, numpy as np, torch
tree, tmp, out = sys.argv[1:4]
sys.path.insert(0, tree)
import src.models.stage_1.raft_wrapper as W
W.device = torch.device("cpu")
res = {}
for edge in (100, 2000):
    w = W.RAFTWrapper.__new__(W.RAFTWrapper)
    w.args = types.SimpleNamespace(max_long_edge=edge)
    a, b = w.load_images(tmp + "/b.png", tmp + "/a.png")          # unsorted on purpose: the loader sorts
    res["im1_%d" % edge], res["im2_%d" % edge] = a.numpy(), b.numpy()
    res["single_%d" % edge] = w.load_image(tmp + "/a.png").numpy()
np.savez(out, **res)
"""


def main():
    rng = np.random.RandomState(21)
    a = rng.randint(0, 256, (130, 250, 3)).astype(np.uint8)
    b = rng.randint(0, 256, (130, 250, 3)).astype(np.uint8)
    with tempfile.TemporaryDirectory() as tmp:
        Image.fromarray(a).save(os.path.join(tmp, "a.png"))
        Image.fromarray(b).save(os.path.join(tmp, "b.png"))
        outs = {}
        for tag, tree in (("ref", "/root/reference"), ("ours", os.path.join(ROOT, "all-in-one-deflicker_b200"))):
            dst = os.path.join(tmp, tag + ".npz")
            subprocess.run([sys.executable, "-c", WORKER, tree, tmp, dst], check=True)
            outs[tag] = dict(np.load(dst))
    for k, v in outs["ref"].items():
        assert v.dtype == outs["ours"][k].dtype and np.array_equal(v, outs["ours"][k]), k
    assert outs["ref"]["im1_100"].shape == (1, 3, 56, 104)          # 130x250 -> 52x100 -> padded to 56x104
    np.savez_compressed(os.path.join(OUT, "raft_io.npz"), a=a, b=b, **{"want_" + k: v for k, v in outs["ref"].items()})
    print("RAFT image loading bit-identical to the reference; fixture written")


if __name__ == "__main__":
    main()
