"""Checkpoint + full-video render + PSNR of the reference's `evaluate_model_single`
(src/models/stage_1/evaluate.py:605-793).  The dashboards / mp4 dumps / tensorboard images of the
reference are visualisation and out of scope."""
import os

import cv2
import numpy as np
import torch

from b200 import atlas as A


def evaluate_model_single(trainer, resx, resy, number_of_frames, video_frames, results_folder, iteration,
                          vid_name=None, save_checkpoint=True):
    os.makedirs(os.path.join(results_folder, '%06d' % iteration), exist_ok=True)
    os.makedirs(os.path.join(results_folder, "output"), exist_ok=True)
    if save_checkpoint:      # evaluate.py:616-622 — same file name and keys
        torch.save({'F_atlas_state_dict': {k: v.cpu() for k, v in trainer.state_dict("atlas").items()},
                    'iteration': iteration,
                    'model_F_mapping1_state_dict': {k: v.cpu() for k, v in trainer.state_dict("mapping").items()},
                    'optimizer_all_state_dict': trainer.optimizer_state_dict()},
                   '%s/checkpoint' % results_folder)
    psnrs = np.zeros((number_of_frames, 1))
    for f in range(number_of_frames):
        img, u8 = trainer.render_frame(f, int(resy), int(resx), number_of_frames, want_u8=True)
        cv2.imwrite(os.path.join(results_folder, 'output', '%05d.png' % f),
                    cv2.cvtColor(u8.cpu().numpy(), cv2.COLOR_RGB2BGR))       # evaluate.py:732-733
        psnrs[f] = A.psnr(video_frames[:, :, :, f], img.cpu())               # :740-743
    open(os.path.join(results_folder, '%06d' % iteration, "PSNR_%f" % psnrs.mean()), "w").close()   # :782
    print("PSNR: %f" % psnrs.mean())
    return float(psnrs.mean())
