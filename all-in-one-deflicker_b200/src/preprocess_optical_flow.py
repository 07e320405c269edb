"""RAFT flow pre-pass with the reference's CLI and output naming (src/preprocess_optical_flow.py):
`<vid>_flow/{fn1}_{fn2}.npy` and `{fn2}_{fn1}.npy`, (H, W, 2) fp32, for consecutive frames."""
import argparse
import os
import sys
from pathlib import Path

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
from tqdm import tqdm  # noqa: E402


def preprocess(args):
    from src.models.stage_1.raft_wrapper import RAFTWrapper
    files = sorted(args.vid_path.glob('*.*g'))
    out_dir = args.vid_path.parent / f'{args.vid_path.name}_flow'
    out_dir.mkdir(exist_ok=True)
    ckpt = 'pretrained_weights/raft-things.pth'
    wrapper = RAFTWrapper(model_path=ckpt if os.path.exists(ckpt) else None, max_long_edge=args.max_long_edge)
    for f1, f2 in tqdm(list(zip(files[:-1], files[1:])), desc='computing flow'):
        o12, o21 = out_dir / f'{f1.name}_{f2.name}.npy', out_dir / f'{f2.name}_{f1.name}.npy'
        if not o12.exists() and not o21.exists():
            im1, im2 = wrapper.load_images(str(f1), str(f2))
            flow12, flow21 = wrapper.compute_flow_both(im1, im2)      # one feature-encoder pass for both directions
            np.save(o12, flow12)
            np.save(o21, flow21)


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description='Preprocess image sequence')
    parser.add_argument('--vid-path', type=Path, default=Path('./data/'), help='folder to process')
    parser.add_argument('--max_long_edge', type=int, default=2000)
    parser.add_argument('--gpu', type=int, default=0)
    args = parser.parse_args()
    os.environ["CUDA_VISIBLE_DEVICES"] = "%d" % args.gpu
    preprocess(args)
